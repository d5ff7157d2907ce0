// eo_contract.h -- even-odd decomposition of the 1-D contractions (build option -DB200FE_EVEN_ODD; OFF by default).
//
// The 1-D matrices of a real basis have a point symmetry: B(nq-1-q, nm-1-i) = B(q, i) (GLL nodes and Gauss / GLL points are
// symmetric about the cell centre) and D(nq-1-p, nq-1-n) = -D(p, n).  Splitting the input into its even and odd parts
//   e_n = in[n] + in[N-1-n],  o_n = in[n] - in[N-1-n]
// halves the multiply-adds of every contraction (the trick deal.II's CPU kernels use, evaluate_evenodd):
//   B  (q <- i):  E_q = sum_i Be[q][i] e_i,  O_q = sum_i Bo[q][i] o_i,   out[q] = E + O,  out[nq-1-q] = E - O
//   B^T(i <- q):  E_i = sum_q Be[q][i] e_q,  O_i = sum_q Bo[q][i] o_q,   out[i] = E + O,  out[nm-1-i] = E - O
//   D  (p <- n):  E_p = sum_n De[p][n] e_n,  O_p = sum_n Do[p][n] o_n,   out[p] = E + O,  out[nq-1-p] = O - E
//   D^T(n <- p):  X_n = sum_p Do[p][n] e_p,  Y_n = sum_p De[p][n] o_p,   out[n] = X + Y,  out[nq-1-n] = Y - X
// with Me[r][c] = (M(r,c) + M(r,N-1-c))/2 (the middle column, N odd, is M(r,mid) itself), Mo[r][c] = (M(r,c) - M(r,N-1-c))/2,
// rows r < ceil(rows/2).  Why: BK3 / BP3 at p = 7-8 are bound by FP64 issue (44 % pipe utilisation at 34 % DRAM,
// profiles/r01b_bk_v2_captures_summary.txt); the contractions are 80-90 % of their flops.
// Only valid for symmetric matrices: the launcher checks B and D and refuses others (the reference's cos() test matrices).
// Written host/device so that tests can check the algebra on the CPU (csrc/eo_selftest.cc, tests/test_eo_contract.py).
#pragma once

#if defined(__CUDACC__)
#define B200FE_HD __host__ __device__ __forceinline__
#define B200FE_CHD constexpr __host__ __device__
#else
#define B200FE_HD inline
#define B200FE_CHD constexpr
#endif
#if defined(__CUDA_ARCH__)
#define B200FE_UNROLL _Pragma("unroll")
#else
#define B200FE_UNROLL
#endif

namespace b200fe {
namespace eo {

B200FE_CHD int half_up(int n) { return (n + 1) / 2; }
B200FE_CHD int half_dn(int n) { return n / 2; }

template <int NM, int NQ>
struct EoMats {
    static constexpr int QH = half_up(NQ), IH = half_up(NM), IL = half_dn(NM), NH = half_up(NQ), NL = half_dn(NQ);
    double Be[QH * IH];                 // [q][i]
    double Bo[QH * (IL > 0 ? IL : 1)];  // [q][i]
    double De[NH * NH];                 // [p][n]
    double Do[NH * (NL > 0 ? NL : 1)];  // [p][n]
};

// B[q*NM+i], D[p*NQ+n] (BK layout).  Returns the largest violation of the symmetries, relative to the largest entry.
template <int NM, int NQ>
double fill(const double *B, const double *D, EoMats<NM, NQ> &m)
{
    using E = EoMats<NM, NQ>;
    double viol = 0.0, big = 0.0;
    auto absd = [](double x) { return x < 0 ? -x : x; };
    for (int q = 0; q < NQ; ++q)
        for (int i = 0; i < NM; ++i) {
            const double a = B[q * NM + i], b = B[(NQ - 1 - q) * NM + (NM - 1 - i)];
            if (absd(a) > big) big = absd(a);
            if (absd(a - b) > viol) viol = absd(a - b);
        }
    for (int p = 0; p < NQ; ++p)
        for (int n = 0; n < NQ; ++n) {
            const double a = D[p * NQ + n], b = D[(NQ - 1 - p) * NQ + (NQ - 1 - n)];
            if (absd(a) > big) big = absd(a);
            if (absd(a + b) > viol) viol = absd(a + b);
        }
    for (int q = 0; q < E::QH; ++q) {
        for (int i = 0; i < E::IL; ++i) {
            m.Be[q * E::IH + i] = 0.5 * (B[q * NM + i] + B[q * NM + NM - 1 - i]);
            m.Bo[q * E::IL + i] = 0.5 * (B[q * NM + i] - B[q * NM + NM - 1 - i]);
        }
        if (NM % 2) m.Be[q * E::IH + E::IL] = B[q * NM + E::IL];
    }
    for (int p = 0; p < E::NH; ++p) {
        for (int n = 0; n < E::NL; ++n) {
            m.De[p * E::NH + n] = 0.5 * (D[p * NQ + n] + D[p * NQ + NQ - 1 - n]);
            m.Do[p * E::NL + n] = 0.5 * (D[p * NQ + n] - D[p * NQ + NQ - 1 - n]);
        }
        if (NQ % 2) m.De[p * E::NH + E::NL] = D[p * NQ + E::NL];
    }
    return big > 0 ? viol / big : 0.0;
}

// even / odd parts of a register column
template <int N>
B200FE_HD void split(const double (&in)[N], double (&e)[half_up(N)], double (&o)[half_up(N)])
{
B200FE_UNROLL
    for (int n = 0; n < half_dn(N); ++n) {
        e[n] = in[n] + in[N - 1 - n];
        o[n] = in[n] - in[N - 1 - n];
    }
    if (N % 2) {
        e[half_dn(N)] = in[half_dn(N)];
        o[half_dn(N)] = 0.0;
    }
}

// out[q] = sum_i B(q,i) in[i]
template <int NM, int NQ>
B200FE_HD void interp(const EoMats<NM, NQ> &m, const double (&in)[NM], double (&out)[NQ])
{
    using E = EoMats<NM, NQ>;
    double e[E::IH], o[E::IH];
    split<NM>(in, e, o);
B200FE_UNROLL
    for (int q = 0; q < E::QH; ++q) {
        double se = 0.0, so = 0.0;
B200FE_UNROLL
        for (int i = 0; i < E::IH; ++i) se = fma(m.Be[q * E::IH + i], e[i], se);
B200FE_UNROLL
        for (int i = 0; i < E::IL; ++i) so = fma(m.Bo[q * E::IL + i], o[i], so);
        out[q] = se + so;
        if (q != NQ - 1 - q) out[NQ - 1 - q] = se - so;
    }
}

// out[i] = sum_q B(q,i) in[q]
template <int NM, int NQ>
B200FE_HD void interp_t(const EoMats<NM, NQ> &m, const double (&in)[NQ], double (&out)[NM])
{
    using E = EoMats<NM, NQ>;
    double e[E::QH], o[E::QH];
    split<NQ>(in, e, o);
B200FE_UNROLL
    for (int i = 0; i < E::IH; ++i) {
        double se = 0.0, so = 0.0;
B200FE_UNROLL
        for (int q = 0; q < E::QH; ++q) se = fma(m.Be[q * E::IH + i], e[q], se);
        if (i < E::IL) {
B200FE_UNROLL
            for (int q = 0; q < half_dn(NQ); ++q) so = fma(m.Bo[q * E::IL + i], o[q], so);
        }
        out[i] = se + so;
        if (i != NM - 1 - i) out[NM - 1 - i] = se - so;
    }
}

// out[p] = sum_n D(p,n) in[n]
template <int NM, int NQ>
B200FE_HD void deriv(const EoMats<NM, NQ> &m, const double (&in)[NQ], double (&out)[NQ])
{
    using E = EoMats<NM, NQ>;
    double e[E::NH], o[E::NH];
    split<NQ>(in, e, o);
B200FE_UNROLL
    for (int p = 0; p < E::NH; ++p) {
        double se = 0.0, so = 0.0;
B200FE_UNROLL
        for (int n = 0; n < E::NH; ++n) se = fma(m.De[p * E::NH + n], e[n], se);
B200FE_UNROLL
        for (int n = 0; n < E::NL; ++n) so = fma(m.Do[p * E::NL + n], o[n], so);
        out[p] = se + so;
        if (p != NQ - 1 - p) out[NQ - 1 - p] = so - se;
    }
}

// out[n] = sum_p D(p,n) in[p]
template <int NM, int NQ>
B200FE_HD void deriv_t(const EoMats<NM, NQ> &m, const double (&in)[NQ], double (&out)[NQ])
{
    using E = EoMats<NM, NQ>;
    double e[E::NH], o[E::NH];
    split<NQ>(in, e, o);
B200FE_UNROLL
    for (int n = 0; n < E::NH; ++n) {
        double sx = 0.0, sy = 0.0;
        if (n < E::NL) {
B200FE_UNROLL
            for (int p = 0; p < E::NH; ++p) sx = fma(m.Do[p * E::NL + n], e[p], sx);
        }
B200FE_UNROLL
        for (int p = 0; p < E::NL; ++p) sy = fma(m.De[p * E::NH + n], o[p], sy);
        out[n] = sx + sy;
        if (n != NQ - 1 - n) out[NQ - 1 - n] = sy - sx;
    }
}

// A square matrix that is symmetric AND point-symmetric (A = A^T, A(n-1-i, n-1-j) = A(i,j): the 1-D stiffness and mass
// matrices of a real basis on symmetric points): its even and odd halves are symmetric again and are stored packed (upper
// triangle, by rows).  sumfact_cart.cuh contracts with these.
template <int N>
struct SymEo {
    static constexpr int H = (N + 1) / 2, L = N / 2;
    double e[H * (H + 1) / 2];                      // even half  Ae[q][i], q <= i < H, packed by rows
    double o[L * (L + 1) / 2 > 0 ? L * (L + 1) / 2 : 1];  // odd half Ao[q][i], q <= i < L
    static B200FE_CHD int at(int n, int r, int c) { return r <= c ? r * n - r * (r - 1) / 2 + (c - r) : c * n - c * (c - 1) / 2 + (r - c); }
    // from the full row-major matrix; returns the largest violation of A = A^T and A(n-1-i, n-1-j) = A(i,j), relative
    double fill(const double *A)
    {
        double viol = 0.0, big = 0.0;
        auto ab = [](double x) { return x < 0 ? -x : x; };
        for (int i = 0; i < N; ++i)
            for (int j = 0; j < N; ++j) {
                const double a = A[i * N + j];
                if (ab(a) > big) big = ab(a);
                if (ab(a - A[j * N + i]) > viol) viol = ab(a - A[j * N + i]);
                if (ab(a - A[(N - 1 - i) * N + (N - 1 - j)]) > viol) viol = ab(a - A[(N - 1 - i) * N + (N - 1 - j)]);
            }
        for (int q = 0; q < H; ++q)
            for (int i = q; i < H; ++i) e[at(H, q, i)] = (N % 2 && i == L) ? A[q * N + i] : 0.5 * (A[q * N + i] + A[q * N + N - 1 - i]);
        for (int q = 0; q < L; ++q)
            for (int i = q; i < L; ++i) o[at(L, q, i)] = 0.5 * (A[q * N + i] - A[q * N + N - 1 - i]);
        if (L == 0) o[0] = 0.0;
        return big > 0 ? viol / big : 0.0;
    }
};


// out = A in through the even-odd split, A given by its packed halves
template <int N>
B200FE_HD void sym_apply(const SymEo<N> &A, const double (&in)[N], double (&out)[N])
{
    constexpr int H = SymEo<N>::H, L = SymEo<N>::L;
    double ev[H], od[H];
    split<N>(in, ev, od);
B200FE_UNROLL
    for (int q = 0; q < H; ++q) {
        double se = 0.0, so = 0.0;
B200FE_UNROLL
        for (int i = 0; i < H; ++i) se = fma(A.e[SymEo<N>::at(H, q, i)], ev[i], se);
        if (q < L) {
B200FE_UNROLL
            for (int i = 0; i < L; ++i) so = fma(A.o[SymEo<N>::at(L, q, i)], od[i], so);
        }
        out[q] = se + so;
        if (q != N - 1 - q) out[N - 1 - q] = se - so;
    }
}

}  // namespace eo
}  // namespace b200fe
