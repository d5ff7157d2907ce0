// eo_selftest.cc -- host-only check of eo_contract.h: the four even-odd contractions against the plain sums, for every
// (nm, nq) the library instantiates, with the real 1-D matrices (Gauss and Gauss-Lobatto points).  Prints the largest
// relative deviation; exit code 0 when below 1e-13.   g++ -std=c++17 -I. eo_selftest.cc basis.cc <error stubs>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <vector>

#include "common.h"
#include "eo_contract.h"

namespace b200fe {
char *error_buffer() { static thread_local char buf[512]; return buf; }
int fail(int code, const char *fmt, ...) { va_list ap; va_start(ap, fmt); vsnprintf(error_buffer(), 512, fmt, ap); va_end(ap); return code; }
int fail_cuda(cudaError_t, const char *) { return 3; }
}  // namespace b200fe

using namespace b200fe;

static double worst = 0.0, worst_sym = 0.0;

template <int NM, int NQ>
void check(int quad)
{
    const int p = NM - 1;
    std::vector<double> sv(NM * NQ), cg(NQ * NQ), B(NQ * NM), D(NQ * NQ);
    if (b200fe_basis_1d(p, NQ, quad, sv.data(), cg.data(), nullptr, nullptr, nullptr)) { std::printf("basis failed\n"); worst = 1; return; }
    for (int q = 0; q < NQ; ++q) {  // BK layouts as in b200fe_op_create
        for (int i = 0; i < NM; ++i) B[q * NM + i] = sv[i * NQ + q];
        for (int n = 0; n < NQ; ++n) D[q * NQ + n] = cg[n * NQ + q];
    }
    eo::EoMats<NM, NQ> m;
    const double sym = eo::fill<NM, NQ>(B.data(), D.data(), m);
    if (sym > worst_sym) worst_sym = sym;
    double in_m[NM], in_q[NQ], out_q[NQ], out_m[NM];
    for (int i = 0; i < NM; ++i) in_m[i] = std::sin(1.0 + 3.7 * i) + 0.3 * i;
    for (int q = 0; q < NQ; ++q) in_q[q] = std::cos(0.5 + 2.3 * q) - 0.1 * q * q;
    auto dev = [&](const double *got, const double *want, int n) {
        double big = 0, d = 0;
        for (int k = 0; k < n; ++k) { big = std::fmax(big, std::fabs(want[k])); d = std::fmax(d, std::fabs(got[k] - want[k])); }
        if (d / big > worst) worst = d / big;
    };
    double ref_q[NQ], ref_m[NM];
    eo::interp<NM, NQ>(m, in_m, out_q);
    for (int q = 0; q < NQ; ++q) { double s = 0; for (int i = 0; i < NM; ++i) s += B[q * NM + i] * in_m[i]; ref_q[q] = s; }
    dev(out_q, ref_q, NQ);
    eo::interp_t<NM, NQ>(m, in_q, out_m);
    for (int i = 0; i < NM; ++i) { double s = 0; for (int q = 0; q < NQ; ++q) s += B[q * NM + i] * in_q[q]; ref_m[i] = s; }
    dev(out_m, ref_m, NM);
    eo::deriv<NM, NQ>(m, in_q, out_q);
    for (int pp = 0; pp < NQ; ++pp) { double s = 0; for (int n = 0; n < NQ; ++n) s += D[pp * NQ + n] * in_q[n]; ref_q[pp] = s; }
    dev(out_q, ref_q, NQ);
    eo::deriv_t<NM, NQ>(m, in_q, out_q);
    for (int n = 0; n < NQ; ++n) { double s = 0; for (int pp = 0; pp < NQ; ++pp) s += D[pp * NQ + n] * in_q[pp]; ref_q[n] = s; }
    dev(out_q, ref_q, NQ);
}

// the packed symmetric even-odd halves of the nodal 1-D stiffness and mass matrices (sumfact_cart.cuh) against plain products;
// K = (D B)^T W (D B), M = B^T W B as b200fe_op_create builds them
template <int NM, int NQ>
void check_sym(int quad)
{
    const int p = NM - 1;
    std::vector<double> sv(NM * NQ), cg(NQ * NQ), w(NQ), B(NQ * NM), D(NQ * NQ), DB(NQ * NM), K(NM * NM), M(NM * NM);
    if (b200fe_basis_1d(p, NQ, quad, sv.data(), cg.data(), nullptr, nullptr, w.data())) { std::printf("basis failed\n"); worst = 1; return; }
    for (int q = 0; q < NQ; ++q) {
        for (int i = 0; i < NM; ++i) B[q * NM + i] = sv[i * NQ + q];
        for (int n = 0; n < NQ; ++n) D[q * NQ + n] = cg[n * NQ + q];
    }
    for (int q = 0; q < NQ; ++q)
        for (int i = 0; i < NM; ++i) { double s = 0; for (int n = 0; n < NQ; ++n) s += D[q * NQ + n] * B[n * NM + i]; DB[q * NM + i] = s; }
    for (int i = 0; i < NM; ++i)
        for (int j = 0; j < NM; ++j) {
            double k = 0, m = 0;
            for (int q = 0; q < NQ; ++q) { k += DB[q * NM + i] * w[q] * DB[q * NM + j]; m += B[q * NM + i] * w[q] * B[q * NM + j]; }
            K[i * NM + j] = k; M[i * NM + j] = m;
        }
    double in[NM], out[NM];
    for (int i = 0; i < NM; ++i) in[i] = std::sin(0.7 + 2.9 * i) - 0.2 * i;
    for (const std::vector<double> *A : {&K, &M}) {
        eo::SymEo<NM> h;
        const double sym = h.fill(A->data());
        if (sym > worst_sym) worst_sym = sym;
        eo::sym_apply<NM>(h, in, out);
        double big = 0, d = 0;
        for (int i = 0; i < NM; ++i) {
            double s = 0;
            for (int j = 0; j < NM; ++j) s += (*A)[i * NM + j] * in[j];
            big = std::fmax(big, std::fabs(s)); d = std::fmax(d, std::fabs(out[i] - s));
        }
        if (d / big > worst) worst = d / big;
    }
}

template <int P>
void degree()
{
    check_sym<P + 1, P + 2>(B200FE_QUAD_GAUSS);
    check_sym<P + 1, P + 1>(B200FE_QUAD_GAUSS);
    check_sym<P + 1, P + 1>(B200FE_QUAD_GLL);
    check<P + 1, P + 2>(B200FE_QUAD_GAUSS);
    check<P + 1, P + 1>(B200FE_QUAD_GAUSS);
    check<P + 1, P + 1>(B200FE_QUAD_GLL);
}

int main()
{
    degree<1>(); degree<2>(); degree<3>(); degree<4>(); degree<5>(); degree<6>(); degree<7>(); degree<8>();
    std::printf("eo_selftest: worst relative deviation %.3e, worst symmetry violation of the real matrices %.3e\n", worst, worst_sym);
    return worst < 1e-13 && worst_sym < 1e-12 ? 0 : 1;
}
