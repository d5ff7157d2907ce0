/*
 * oracle/fe_oracle.c -- TEST INFRASTRUCTURE / CPU BASELINE ONLY (never linked into the product).
 *
 * CPU restatement (C + OpenMP) of the reference's BP operator path on process-local
 * ("L") vectors and of the CG loop that drives it:
 *
 *   gather with invalid-index mask        CEED_bp/include/bk3_kokkos_kernel.h:115-141
 *   forward interpolation (3 sweeps)      CEED_bp/include/bk3_kokkos_kernel.h:147-208
 *   collocation gradient, G, transpose    CEED_bp/include/bk3_kokkos_kernel.h:212-284
 *   backward interpolation                CEED_bp/include/bk3_kokkos_kernel.h:292-353
 *   additive scatter                      CEED_bp/include/bk3_kokkos_kernel.h:357-381
 *   dst=0 ... constrained rows = identity CEED_bp/include/portable_laplace_operator.h:124-172
 *   mass + Laplace (Helmholtz) quad op    bp5_kokkos/benchmark.cc:62-137
 *   SolverCG / ReductionControl           deal.II 9.7 (un-vendored); call site CEED_bp/src/bp3.cc:268-285,
 *                                         restated from SURVEY.md Appendix A9
 *
 * Array conventions are the reference operator's (deal.II) ones:
 *   shape_values[i*nq+q]        value of shape function i at point q     (bk3_kokkos_kernel.h:161)
 *   co_shape_gradients[n*nq+q]  derivative of collocation function n at q (bk3_kokkos_kernel.h:243)
 *   dof_indices[cell*nm^3 + l]  l = i*nm^2+j*nm+k (k = x fastest), 0xFFFFFFFF = masked
 *   G[cell][6][nq^3], point index p*nq^2+q*nq+r (p <-> i <-> slowest), components rr,rs,rt,ss,st,tt
 *   JxW[cell][nq^3]
 *
 * Scatter races are avoided with a caller-supplied cell colouring (cells of one colour share
 * no DoF); the colour loop is outside the OpenMP loop.  Pinned by tests/test_oracle_pins.py
 * against oracle/fe_oracle.py (which reproduces the CG goldens of
 * CEED_bp/results/1xGH200_P4.txt:636-640).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define INVALID_INDEX 0xFFFFFFFFu
#define MAXN 12

enum { OP_LAPLACE = 1, OP_MASS = 2 };

#ifdef __cplusplus
extern "C" {
#endif

int oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* out[a][m][c] = sum_n M(m,n) in[a][n][c]; M(m,n) = M[m*rs + n*cs] */
static void contract(int na, int nin, int nout, int nc, const double *M, int rs, int cs,
                     const double *in, double *out)
{
    for (int a = 0; a < na; ++a)
        for (int m = 0; m < nout; ++m) {
            double *o = out + ((size_t)a * nout + m) * nc;
            for (int c = 0; c < nc; ++c) o[c] = 0.0;
            for (int n = 0; n < nin; ++n) {
                const double w = M[m * rs + n * cs];
                const double *x = in + ((size_t)a * nin + n) * nc;
                for (int c = 0; c < nc; ++c) o[c] += w * x[c];
            }
        }
}

typedef struct {
    int nm, nq, collocated, flags;
    const double *S;  /* shape_values[i*nq+q] */
    const double *Dg; /* co_shape_gradients[n*nq+q] */
} cell_ctx;

/* u[nm^3] -> out[nm^3]; ws needs 7*nq^3 doubles */
static void cell_apply(const cell_ctx *c, const double *G, const double *JxW, const double *u,
                       double *out, double *ws)
{
    const int nm = c->nm, nq = c->nq, n2 = nq * nq, n3 = n2 * nq;
    double *t0 = ws, *t1 = ws + n3, *v = ws + 2 * n3, *gr = ws + 3 * n3, *gs = ws + 4 * n3,
           *gt = ws + 5 * n3, *w = ws + 6 * n3;
    if (c->collocated)
        memcpy(v, u, sizeof(double) * n3);
    else {
        /* B(q,i) = S[i*nq+q]: row stride 1, col stride nq */
        contract(1, nm, nq, nm * nm, c->S, 1, nq, u, t0);
        contract(nq, nm, nq, nm, c->S, 1, nq, t0, t1);
        contract(nq * nq, nm, nq, 1, c->S, 1, nq, t1, v);
    }
    for (int i = 0; i < n3; ++i) w[i] = 0.0;
    if (c->flags & OP_LAPLACE) {
        /* D(p,n) = Dg[n*nq+p] */
        contract(1, nq, nq, n2, c->Dg, 1, nq, v, gr);
        contract(nq, nq, nq, nq, c->Dg, 1, nq, v, gs);
        contract(n2, nq, nq, 1, c->Dg, 1, nq, v, gt);
        for (int i = 0; i < n3; ++i) {
            const double qr = gr[i], qs = gs[i], qt = gt[i];
            gr[i] = G[i] * qr + G[n3 + i] * qs + G[2 * n3 + i] * qt;
            gs[i] = G[n3 + i] * qr + G[3 * n3 + i] * qs + G[4 * n3 + i] * qt;
            gt[i] = G[2 * n3 + i] * qr + G[4 * n3 + i] * qs + G[5 * n3 + i] * qt;
        }
        /* D^T(p,n) = D(n,p) = Dg[p*nq+n] */
        contract(1, nq, nq, n2, c->Dg, nq, 1, gr, t0);
        for (int i = 0; i < n3; ++i) w[i] += t0[i];
        contract(nq, nq, nq, nq, c->Dg, nq, 1, gs, t0);
        for (int i = 0; i < n3; ++i) w[i] += t0[i];
        contract(n2, nq, nq, 1, c->Dg, nq, 1, gt, t0);
        for (int i = 0; i < n3; ++i) w[i] += t0[i];
    }
    if (c->flags & OP_MASS)
        for (int i = 0; i < n3; ++i) w[i] += JxW[i] * v[i];
    if (c->collocated)
        memcpy(out, w, sizeof(double) * n3);
    else {
        /* B^T(i,q) = S[i*nq+q] */
        contract(nq * nq, nq, nm, 1, c->S, nq, 1, w, t0);
        contract(nq, nq, nm, nm, c->S, nq, 1, t0, t1);
        contract(1, nq, nm, nm * nm, c->S, nq, 1, t1, out);
    }
}

/*
 * dst = A src on one rank's local vector (no halo exchange).
 * color_offsets[n_colors+1] / color_cells[] : CSR list of cells per colour (NULL -> serial loop).
 * constrained[n_constrained] : owned local indices with dst[c] = src[c].
 */
void oracle_op_apply(int nm, int nq, int collocated, int flags, uint32_t n_cells, uint32_t n_local,
                     const double *shape_values, const double *co_shape_gradients,
                     const double *G, const double *JxW, const uint32_t *dof_indices,
                     int n_colors, const uint32_t *color_offsets, const uint32_t *color_cells,
                     uint32_t n_constrained, const uint32_t *constrained,
                     const double *src, double *dst)
{
    const size_t nm3 = (size_t)nm * nm * nm, nq3 = (size_t)nq * nq * nq;
    cell_ctx ctx = {nm, nq, collocated, flags, shape_values, co_shape_gradients};
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n_local; ++i) dst[i] = 0.0;
    const int ncol = color_offsets ? n_colors : 1;
    for (int col = 0; col < ncol; ++col) {
        const uint32_t cb = color_offsets ? color_offsets[col] : 0;
        const uint32_t ce = color_offsets ? color_offsets[col + 1] : n_cells;
#pragma omp parallel if (color_offsets != NULL)
        {
            double *ws = (double *)malloc(sizeof(double) * (7 * nq3 + 2 * nm3));
            double *u = ws + 7 * nq3, *o = u + nm3;
#pragma omp for schedule(static)
            for (uint32_t ci = cb; ci < ce; ++ci) {
                const uint32_t cell = color_cells ? color_cells[ci] : ci;
                const uint32_t *idx = dof_indices + (size_t)cell * nm3;
                for (size_t l = 0; l < nm3; ++l) u[l] = idx[l] == INVALID_INDEX ? 0.0 : src[idx[l]];
                cell_apply(&ctx, G ? G + (size_t)cell * 6 * nq3 : NULL, JxW ? JxW + (size_t)cell * nq3 : NULL,
                           u, o, ws);
                for (size_t l = 0; l < nm3; ++l)
                    if (idx[l] != INVALID_INDEX) dst[idx[l]] += o[l];
            }
            free(ws);
        }
    }
    for (uint32_t i = 0; i < n_constrained; ++i) dst[constrained[i]] = src[constrained[i]];
}

static double dot(const double *a, const double *b, size_t n)
{
    double s = 0.0;
#pragma omp parallel for reduction(+ : s) schedule(static)
    for (size_t i = 0; i < n; ++i) s += a[i] * b[i];
    return s;
}

/*
 * deal.II SolverCG with ReductionControl(max_it, abs_tol, rel_tol), x0 = 0, optional Jacobi
 * (inv_diag != NULL).  Single rank (n_ghost == 0).  Returns the iteration count; *converged = 0
 * mirrors SolverControl::NoConvergence.  work: 4*n_local doubles are allocated internally.
 */
int oracle_cg_solve(int nm, int nq, int collocated, int flags, uint32_t n_cells, uint32_t n_local,
                    const double *shape_values, const double *co_shape_gradients,
                    const double *G, const double *JxW, const uint32_t *dof_indices,
                    int n_colors, const uint32_t *color_offsets, const uint32_t *color_cells,
                    uint32_t n_constrained, const uint32_t *constrained,
                    const double *inv_diag, const double *b, double *x,
                    int max_it, double abs_tol, double rel_tol,
                    double *res0_out, double *resn_out, int *converged)
{
    const size_t n = n_local;
    double *r = (double *)malloc(sizeof(double) * n), *p = (double *)malloc(sizeof(double) * n),
           *v = (double *)malloc(sizeof(double) * n), *z = inv_diag ? (double *)malloc(sizeof(double) * n) : NULL;
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; ++i) { x[i] = 0.0; r[i] = b[i]; }
    double res = sqrt(dot(r, r, n));
    const double res0 = res;
    int it = 0, ok = res <= abs_tol;
    double rho = 0.0, rho_old;
    while (!ok) {
        ++it;
        rho_old = rho;
        if (inv_diag) {
#pragma omp parallel for schedule(static)
            for (size_t i = 0; i < n; ++i) z[i] = inv_diag[i] * r[i];
            rho = dot(r, z, n);
        } else
            rho = res * res;
        const double *d = inv_diag ? z : r;
        if (it == 1) {
#pragma omp parallel for schedule(static)
            for (size_t i = 0; i < n; ++i) p[i] = d[i];
        } else {
            const double beta = rho / rho_old;
#pragma omp parallel for schedule(static)
            for (size_t i = 0; i < n; ++i) p[i] = d[i] + beta * p[i];
        }
        oracle_op_apply(nm, nq, collocated, flags, n_cells, n_local, shape_values, co_shape_gradients, G, JxW,
                        dof_indices, n_colors, color_offsets, color_cells, n_constrained, constrained, p, v);
        const double alpha = rho / dot(p, v, n);
        double rr = 0.0;
#pragma omp parallel for reduction(+ : rr) schedule(static)
        for (size_t i = 0; i < n; ++i) {
            x[i] += alpha * p[i];
            r[i] -= alpha * v[i];
            rr += r[i] * r[i];
        }
        res = sqrt(fabs(rr));
        if (res <= abs_tol || res <= rel_tol * res0) { ok = 1; break; }
        if (it >= max_it) break;
    }
    free(r); free(p); free(v); free(z);
    if (res0_out) *res0_out = res0;
    if (resn_out) *resn_out = res;
    if (converged) *converged = ok;
    return it;
}

#ifdef __cplusplus
}
#endif
