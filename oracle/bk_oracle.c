/*
 * oracle/bk_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * CPU restatement, in plain C, of the reference's serial sum-factorisation
 * kernels for the CEED bake-off kernels on element-local ("E") vectors:
 *
 *   BK1 mass        out_e = B^T ( JxW .* (B u_e) )
 *       follows CEED_BK/include/kernels/BK1/serial_kernels.hpp:9-135
 *   BK3 Laplacian   out_e = B^T D^T G D B u_e          (nq Gauss points, nm modes)
 *       follows CEED_BK/include/kernels/BK3/serial_kernels.hpp:10-197
 *   BK5 collocated  out_e = D^T G D u_e                (nm == nq)
 *       follows sum_factorization/include/kernels/BK5/serial_kernels.hpp:10-85
 *       (the CEED_BK copy reads geometric-factor component 0 six times,
 *        CEED_BK/include/kernels/BK5/serial_kernels.hpp:28-33)
 *
 * Layout conventions (all taken from the reference):
 *   in/out   [e][i][j][k]            k fastest     (BK1 serial :29)
 *   basis    [q][i]  = basis[q*nm+i]               (BK1 serial :39)
 *   dbasis   [p][n]  = dbasis[p*nq+n]  derivative of collocation function n at
 *                                      point p      (BK3 serial :98)
 *   JxW      [e][p][q][r]                          (BK1 serial :77)
 *   G        g_layout == 0 : [e][p][q][6][r]       (BK3 serial :87-92)
 *            g_layout == 1 : [e][6][p][q][r]       (BK3 CUDA templated_cuda_kernels.cuh:156-161;
 *                                                   this is the product layout)
 *   symmetric components 0..5 = rr, rs, rt, ss, st, tt with r the direction of the
 *   SLOWEST local index (p / i), t the fastest (BK3 serial :110-112).
 *
 * Summation order inside every 1-D contraction is ascending in the contracted
 * index, like the reference's "+=" loops, so results agree with the compiled
 * reference (oracle/_ref) to the last bit when both are built without FMA
 * contraction.  Parity is pinned by tests/test_oracle_pins.py against
 * (a) oracle/_ref (the reference headers compiled in place) and
 * (b) the golden norm table in tests/golden/bk_norms.json.
 *
 * The general (nm, nq) signature is a superset of the reference (which fixes
 * nm = nq-1 for BK1/BK3): the L-vector operators also use nq = nm ("bp35").
 */
#include <stdlib.h>
#include <string.h>

#ifdef __cplusplus
extern "C" {
#endif

/* out[a][m][c] = sum_n M[m*ldm_row + n*ldm_col] * in[a][n][c]
 * a in [0,na), c in [0,nc): one tensor contraction along the middle axis. */
static void contract_mid(int na, int nin, int nout, int nc,
                         const double *M, int row_stride, int col_stride,
                         const double *in, double *out)
{
    for (int a = 0; a < na; ++a)
        for (int m = 0; m < nout; ++m)
            for (int c = 0; c < nc; ++c) {
                double s = 0.0;
                for (int n = 0; n < nin; ++n)
                    s += in[((size_t)a * nin + n) * nc + c] * M[m * row_stride + n * col_stride];
                out[((size_t)a * nout + m) * nc + c] = s;
            }
}

/* forward interpolation nm^3 -> nq^3, direction 0 (slowest) first like the
 * reference (BK1 serial steps 2-4), result indexed [p][q][r]. */
static void interp_forward(int nm, int nq, const double *B, const double *u,
                           double *t0, double *t1, double *v)
{
    /* dir 0: [i][j k] -> [p][j k] */
    contract_mid(1, nm, nq, nm * nm, B, nm, 1, u, t0);
    /* dir 1: [p][j][k] -> [p][q][k] */
    contract_mid(nq, nm, nq, nm, B, nm, 1, t0, t1);
    /* dir 2: [p q][k][1] -> [p q][r][1] */
    contract_mid(nq * nq, nm, nq, 1, B, nm, 1, t1, v);
}

/* transpose of interp_forward: nq^3 -> nm^3 (BK1 serial steps 6-8: direction 2 first). */
static void interp_backward(int nm, int nq, const double *B, const double *w,
                            double *t0, double *t1, double *out)
{
    /* dir 2: [p q][r] -> [p q][k], matrix B^T: M[k][r] = B[r*nm+k] */
    contract_mid(nq * nq, nq, nm, 1, B, 1, nm, w, t0);
    /* dir 1: [p][q][k] -> [p][j][k] */
    contract_mid(nq, nq, nm, nm, B, 1, nm, t0, t1);
    /* dir 0: [p][j k] -> [i][j k] */
    contract_mid(1, nq, nm, nm * nm, B, 1, nm, t1, out);
}

static inline size_t g_index(int g_layout, int nq, int c, int p, int q, int r)
{
    if (g_layout == 0) /* [p][q][6][r] */
        return (((size_t)p * nq + q) * 6 + c) * nq + r;
    return (((size_t)c * nq + p) * nq + q) * nq + r; /* [6][p][q][r] */
}

/* w = D^T G D v on one element's nq^3 quadrature values (BK3 serial steps 5-8). */
static void laplace_at_quad(int nq, const double *D, const double *Ge, int g_layout,
                            const double *v, double *rqr, double *rqs, double *rqt, double *w)
{
    const int n2 = nq * nq;
    for (int p = 0; p < nq; ++p)
        for (int q = 0; q < nq; ++q)
            for (int r = 0; r < nq; ++r) {
                double qr = 0.0, qs = 0.0, qt = 0.0;
                for (int n = 0; n < nq; ++n) qr += v[n * n2 + q * nq + r] * D[p * nq + n];
                for (int n = 0; n < nq; ++n) qs += v[p * n2 + n * nq + r] * D[q * nq + n];
                for (int n = 0; n < nq; ++n) qt += v[p * n2 + q * nq + n] * D[r * nq + n];
                const double Grr = Ge[g_index(g_layout, nq, 0, p, q, r)];
                const double Grs = Ge[g_index(g_layout, nq, 1, p, q, r)];
                const double Grt = Ge[g_index(g_layout, nq, 2, p, q, r)];
                const double Gss = Ge[g_index(g_layout, nq, 3, p, q, r)];
                const double Gst = Ge[g_index(g_layout, nq, 4, p, q, r)];
                const double Gtt = Ge[g_index(g_layout, nq, 5, p, q, r)];
                const int idx = p * n2 + q * nq + r;
                rqr[idx] = Grr * qr + Grs * qs + Grt * qt;
                rqs[idx] = Grs * qr + Gss * qs + Gst * qt;
                rqt[idx] = Grt * qr + Gst * qs + Gtt * qt;
            }
    for (int p = 0; p < nq; ++p)
        for (int q = 0; q < nq; ++q)
            for (int r = 0; r < nq; ++r) {
                double t = 0.0;
                for (int n = 0; n < nq; ++n) t += rqr[n * n2 + q * nq + r] * D[n * nq + p];
                for (int n = 0; n < nq; ++n) t += rqs[p * n2 + n * nq + r] * D[n * nq + q];
                for (int n = 0; n < nq; ++n) t += rqt[p * n2 + q * nq + n] * D[n * nq + r];
                w[p * n2 + q * nq + r] = t;
            }
}

static double sum_sq(const double *x, size_t n)
{
    double s = 0.0;
    for (size_t i = 0; i < n; ++i) s += x[i] * x[i];
    return s;
}

/* BK1: CEED_BK/include/kernels/BK1/serial_kernels.hpp:9-135. Returns sum(out^2). */
double oracle_bk1(int nm, int nq, unsigned nelmt, const double *basis, const double *JxW,
                  const double *in, double *out)
{
    const size_t nm3 = (size_t)nm * nm * nm, nq3 = (size_t)nq * nq * nq;
    double *t0 = (double *)malloc(nq3 * sizeof(double));
    double *t1 = (double *)malloc(nq3 * sizeof(double));
    double *v = (double *)malloc(nq3 * sizeof(double));
    for (unsigned e = 0; e < nelmt; ++e) {
        interp_forward(nm, nq, basis, in + e * nm3, t0, t1, v);
        for (size_t q = 0; q < nq3; ++q) v[q] *= JxW[e * nq3 + q];
        interp_backward(nm, nq, basis, v, t0, t1, out + e * nm3);
    }
    free(t0); free(t1); free(v);
    return sum_sq(out, nelmt * nm3);
}

/* BK3: CEED_BK/include/kernels/BK3/serial_kernels.hpp:10-197. */
double oracle_bk3(int nm, int nq, unsigned nelmt, const double *basis, const double *dbasis,
                  const double *G, int g_layout, const double *in, double *out)
{
    const size_t nm3 = (size_t)nm * nm * nm, nq3 = (size_t)nq * nq * nq;
    double *buf = (double *)malloc(7 * nq3 * sizeof(double));
    double *t0 = buf, *t1 = buf + nq3, *v = buf + 2 * nq3, *rqr = buf + 3 * nq3,
           *rqs = buf + 4 * nq3, *rqt = buf + 5 * nq3, *w = buf + 6 * nq3;
    for (unsigned e = 0; e < nelmt; ++e) {
        interp_forward(nm, nq, basis, in + e * nm3, t0, t1, v);
        laplace_at_quad(nq, dbasis, G + (size_t)e * 6 * nq3, g_layout, v, rqr, rqs, rqt, w);
        interp_backward(nm, nq, basis, w, t0, t1, out + e * nm3);
    }
    free(buf);
    return sum_sq(out, nelmt * nm3);
}

/* BK5: sum_factorization/include/kernels/BK5/serial_kernels.hpp:10-85. */
double oracle_bk5(int nq, unsigned nelmt, const double *dbasis, const double *G, int g_layout,
                  const double *in, double *out)
{
    const size_t nq3 = (size_t)nq * nq * nq;
    double *buf = (double *)malloc(3 * nq3 * sizeof(double));
    for (unsigned e = 0; e < nelmt; ++e)
        laplace_at_quad(nq, dbasis, G + (size_t)e * 6 * nq3, g_layout, in + e * nq3,
                        buf, buf + nq3, buf + 2 * nq3, out + e * nq3);
    free(buf);
    return sum_sq(out, nelmt * nq3);
}

/* Brute-force O(nm^3 nq^3) mass operator, the reference's only cross-algorithm check:
 * sum_factorization/include/kernels/BK1/serial_kernels.hpp:9-61. */
double oracle_bk1_direct(int nm, int nq, unsigned nelmt, const double *basis, const double *JxW,
                         const double *in, double *out)
{
    const size_t nm3 = (size_t)nm * nm * nm, nq3 = (size_t)nq * nq * nq;
    double *v = (double *)malloc(nq3 * sizeof(double));
    for (unsigned e = 0; e < nelmt; ++e) {
        for (int p = 0; p < nq; ++p)
            for (int q = 0; q < nq; ++q)
                for (int r = 0; r < nq; ++r) {
                    double s = 0.0;
                    for (int i = 0; i < nm; ++i)
                        for (int j = 0; j < nm; ++j)
                            for (int k = 0; k < nm; ++k)
                                s += in[e * nm3 + ((size_t)i * nm + j) * nm + k] * basis[p * nm + i] *
                                     basis[q * nm + j] * basis[r * nm + k];
                    const size_t qi = ((size_t)p * nq + q) * nq + r;
                    v[qi] = s * JxW[e * nq3 + qi];
                }
        for (int i = 0; i < nm; ++i)
            for (int j = 0; j < nm; ++j)
                for (int k = 0; k < nm; ++k) {
                    double s = 0.0;
                    for (int p = 0; p < nq; ++p)
                        for (int q = 0; q < nq; ++q)
                            for (int r = 0; r < nq; ++r)
                                s += v[((size_t)p * nq + q) * nq + r] * basis[p * nm + i] *
                                     basis[q * nm + j] * basis[r * nm + k];
                    out[e * nm3 + ((size_t)i * nm + j) * nm + k] = s;
                }
    }
    free(v);
    return sum_sq(out, nelmt * nm3);
}

#ifdef __cplusplus
}
#endif
