// basis.cc -- 1-D quadrature rules and Lagrange bases (host only).
//
// deal.II supplies these to the reference (QGauss<1>(nq), FE_Q(p) on Gauss-Lobatto support points,
// MatrixFree shape_values / co_shape_gradients).  Array layouts are the
// ones the reference operator hands to its cell kernel:
//   shape_values[i*nq+q]        = phi_i(x_q)          (CEED_bp/include/bk3_kokkos_kernel.h:161)
//   co_shape_gradients[n*nq+q]  = l_n'(x_q)           (CEED_bp/include/bk3_kokkos_kernel.h:243)
//   shape_gradients[i*nq+q]     = phi_i'(x_q)
// All on the unit interval [0,1], points ascending, weights summing to 1.
#include <cmath>
#include <vector>

#include "common.h"

namespace b200fe {

static void legendre(int n, double x, double &Pn, double &Pnm1)
{
    double p0 = 1.0, p1 = x;
    if (n == 0) { Pn = 1.0; Pnm1 = 0.0; return; }
    for (int k = 2; k <= n; ++k) {
        const double p2 = ((2 * k - 1) * x * p1 - (k - 1) * p0) / k;
        p0 = p1;
        p1 = p2;
    }
    Pn = p1;
    Pnm1 = p0;
}

// n-point Gauss-Legendre on [0,1]
void gauss_legendre_01(int n, double *x, double *w)
{
    const double pi = 3.14159265358979323846;
    for (int i = 0; i < (n + 1) / 2; ++i) {
        double z = std::cos(pi * (i + 0.75) / (n + 0.5));
        double Pn, Pnm1, dP = 1.0;
        for (int it = 0; it < 100; ++it) {
            legendre(n, z, Pn, Pnm1);
            dP = n * (z * Pn - Pnm1) / (z * z - 1.0);
            const double dz = Pn / dP;
            z -= dz;
            if (std::fabs(dz) < 1e-16) break;
        }
        legendre(n, z, Pn, Pnm1);
        dP = n * (z * Pn - Pnm1) / (z * z - 1.0);
        const double wi = 2.0 / ((1.0 - z * z) * dP * dP);
        x[i] = 0.5 * (1.0 - z);          // ascending
        x[n - 1 - i] = 0.5 * (1.0 + z);
        w[i] = w[n - 1 - i] = 0.5 * wi;
    }
}

// n-point Gauss-Lobatto-Legendre on [0,1] (n >= 2)
void gauss_lobatto_01(int n, double *x, double *w)
{
    const double pi = 3.14159265358979323846;
    const int N = n - 1;
    for (int i = 0; i <= N / 2; ++i) {
        double z = std::cos(pi * i / N);
        double Pn, Pnm1;
        if (i > 0) {
            for (int it = 0; it < 100; ++it) {
                legendre(N, z, Pn, Pnm1);
                // Newton on (1-z^2) P_N'(z) = N (P_{N-1} - z P_N)
                const double dz = (z * Pn - Pnm1) / ((N + 1) * Pn);
                z -= dz;
                if (std::fabs(dz) < 1e-16) break;
            }
        }
        legendre(N, z, Pn, Pnm1);
        const double wi = 2.0 / (N * (N + 1) * Pn * Pn);
        x[i] = 0.5 * (1.0 - z);
        x[N - i] = 0.5 * (1.0 + z);
        w[i] = w[N - i] = 0.5 * wi;
    }
    if (N % 2 == 0) x[N / 2] = 0.5;
}

// V[q*n+i] = l_i(x_q), Dv[q*n+i] = l_i'(x_q) for the Lagrange basis through nodes[0..n)
void lagrange_eval(int n, const double *nodes, int nx, const double *x, double *V, double *Dv)
{
    for (int q = 0; q < nx; ++q)
        for (int i = 0; i < n; ++i) {
            double val = 1.0, der = 0.0;
            for (int m = 0; m < n; ++m)
                if (m != i) val *= (x[q] - nodes[m]) / (nodes[i] - nodes[m]);
            for (int k = 0; k < n; ++k) {
                if (k == i) continue;
                double t = 1.0 / (nodes[i] - nodes[k]);
                for (int m = 0; m < n; ++m)
                    if (m != i && m != k) t *= (x[q] - nodes[m]) / (nodes[i] - nodes[m]);
                der += t;
            }
            if (V) V[q * n + i] = val;
            if (Dv) Dv[q * n + i] = der;
        }
}

}  // namespace b200fe

using namespace b200fe;

extern "C" int b200fe_basis_1d(int p, int nq, int quad_kind, double *h_shape_values,
                               double *h_co_shape_gradients, double *h_shape_gradients,
                               double *h_points, double *h_weights)
{
    if (p < 1 || p > 16 || nq < 1 || nq > 20)
        return fail(B200FE_ERR_UNSUPPORTED, "b200fe_basis_1d: p=%d nq=%d out of range", p, nq);
    if (quad_kind != B200FE_QUAD_GAUSS && quad_kind != B200FE_QUAD_GLL)
        return fail(B200FE_ERR_INVALID_ARG, "b200fe_basis_1d: quad_kind must be GAUSS or GLL");
    if (quad_kind == B200FE_QUAD_GLL && nq < 2)
        return fail(B200FE_ERR_INVALID_ARG, "b200fe_basis_1d: GLL needs nq >= 2");
    const int nm = p + 1;
    std::vector<double> nodes(nm), wn(nm), xq(nq), wq(nq), V(nq * nm), Dv(nq * nm), Dc(nq * nq);
    gauss_lobatto_01(nm, nodes.data(), wn.data());
    if (quad_kind == B200FE_QUAD_GAUSS)
        gauss_legendre_01(nq, xq.data(), wq.data());
    else
        gauss_lobatto_01(nq, xq.data(), wq.data());
    lagrange_eval(nm, nodes.data(), nq, xq.data(), V.data(), Dv.data());
    lagrange_eval(nq, xq.data(), nq, xq.data(), nullptr, Dc.data());
    const bool collocated = quad_kind == B200FE_QUAD_GLL && nq == nm;
    for (int i = 0; i < nm; ++i)
        for (int q = 0; q < nq; ++q) {
            if (h_shape_values) h_shape_values[i * nq + q] = collocated ? (i == q ? 1.0 : 0.0) : V[q * nm + i];
            if (h_shape_gradients) h_shape_gradients[i * nq + q] = Dv[q * nm + i];
        }
    for (int n = 0; n < nq; ++n)
        for (int q = 0; q < nq; ++q)
            if (h_co_shape_gradients) h_co_shape_gradients[n * nq + q] = Dc[q * nq + n];
    for (int q = 0; q < nq; ++q) {
        if (h_points) h_points[q] = xq[q];
        if (h_weights) h_weights[q] = wq[q];
    }
    return B200FE_OK;
}
