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

/* torchrun exports OMP_NUM_THREADS=1 to every rank: the timed CPU arm sets its team size explicitly (bench.py) */
void oracle_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n >= 1) omp_set_num_threads(n);
#else
    (void)n;
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


/* ------------------------------------------------------------------------------------------------------------------
 * Fast path of the same operator for the timed CPU baseline: LANES cells are processed side by side, the lane index
 * being the fastest array dimension (the layout deal.II's CPU path uses: VectorizedArray over n_lanes cells,
 * bk3_dealii/check_bk3.cc:86-113), and the 1-D sizes are compile-time constants (the reference dispatches fe_degree / nq
 * at compile time as well, CEED_bp/src/bp3.cc:536-557).  Every lane performs exactly the arithmetic of cell_apply() in
 * the same order (no FMA contraction, no reassociation), so the results are bitwise those of the generic path.
 * ------------------------------------------------------------------------------------------------------------------ */
#define LANES 8
#define AINLINE static inline __attribute__((always_inline))

AINLINE void contract_k(const int na, const int nin, const int nout, const int nc, const double *M, const int rs, const int cs,
                        const double *restrict in, double *restrict out)
{
    /* same sum order as contract() (n ascending from 0.0), accumulated in a register per output entry */
    for (int a = 0; a < na; ++a)
        for (int m = 0; m < nout; ++m) {
            double *restrict o = out + ((size_t)a * nout + m) * nc;
            const double *restrict x = in + (size_t)a * nin * nc;
#pragma omp simd
            for (int c = 0; c < nc; ++c) {
                double sum = 0.0;
                for (int n = 0; n < nin; ++n) sum += M[m * rs + n * cs] * x[(size_t)n * nc + c];
                o[c] = sum;
            }
        }
}

/* u[nm^3][LANES] -> out[nm^3][LANES]; Gb[6][nq^3][LANES], Jb[nq^3][LANES] (lane-interleaved copies of the batch's
 * geometric factors); ws: 7*nq^3*LANES doubles */
AINLINE void cell_apply_lanes(const int nm, const int nq, const int collocated, const int flags, const double *S, const double *Dg,
                              const double *restrict Gb, const double *restrict Jb, const double *restrict u, double *restrict out,
                              double *restrict ws)
{
    const int L = LANES, n2 = nq * nq, n3 = n2 * nq, n3L = n3 * L;
    double *t0 = ws, *t1 = ws + n3L, *v = ws + 2 * n3L, *gr = ws + 3 * n3L, *gs = ws + 4 * n3L, *gt = ws + 5 * n3L, *w = ws + 6 * n3L;
    const double *vv = u;
    if (!collocated) {
        contract_k(1, nm, nq, nm * nm * L, S, 1, nq, u, t0);
        contract_k(nq, nm, nq, nm * L, S, 1, nq, t0, t1);
        contract_k(nq * nq, nm, nq, L, S, 1, nq, t1, v);
        vv = v;
    }
    for (int i = 0; i < n3L; ++i) w[i] = 0.0;
    if (flags & OP_LAPLACE) {
        contract_k(1, nq, nq, n2 * L, Dg, 1, nq, vv, gr);
        contract_k(nq, nq, nq, nq * L, Dg, 1, nq, vv, gs);
        contract_k(n2, nq, nq, L, Dg, 1, nq, vv, gt);
        const double *g0 = Gb, *g1 = Gb + n3L, *g2 = Gb + 2 * n3L, *g3 = Gb + 3 * n3L, *g4 = Gb + 4 * n3L, *g5 = Gb + 5 * n3L;
#pragma omp simd
        for (int i = 0; i < n3L; ++i) {
            const double qr = gr[i], qs = gs[i], qt = gt[i];
            gr[i] = g0[i] * qr + g1[i] * qs + g2[i] * qt;
            gs[i] = g1[i] * qr + g3[i] * qs + g4[i] * qt;
            gt[i] = g2[i] * qr + g4[i] * qs + g5[i] * qt;
        }
        contract_k(1, nq, nq, n2 * L, Dg, nq, 1, gr, t0);
        for (int i = 0; i < n3L; ++i) w[i] += t0[i];
        contract_k(nq, nq, nq, nq * L, Dg, nq, 1, gs, t0);
        for (int i = 0; i < n3L; ++i) w[i] += t0[i];
        contract_k(n2, nq, nq, L, Dg, nq, 1, gt, t0);
        for (int i = 0; i < n3L; ++i) w[i] += t0[i];
    }
    if (flags & OP_MASS)
        for (int i = 0; i < n3L; ++i) w[i] += Jb[i] * vv[i];
    if (collocated)
        memcpy(out, w, sizeof(double) * n3L);
    else {
        contract_k(nq * nq, nq, nm, L, S, nq, 1, w, t0);
        contract_k(nq, nq, nm, nm * L, S, nq, 1, t0, t1);
        contract_k(1, nq, nm, nm * nm * L, S, nq, 1, t1, out);
    }
}

/* compile-time (nm, nq, collocated) instances of the batch kernel: nq = nm (collocated or not) and nq = nm + 1, degrees 1..8 */
typedef void (*batch_fn)(int flags, const double *S, const double *Dg, const double *Gb, const double *Jb, const double *u, double *out,
                         double *ws);
#define LANES_DEF(NM, NQ, COLL)                                                                                                    \
    static void batch_##NM##_##NQ##_##COLL(int flags, const double *S, const double *Dg, const double *Gb, const double *Jb,        \
                                           const double *u, double *out, double *ws)                                               \
    {                                                                                                                              \
        cell_apply_lanes(NM, NQ, COLL, flags, S, Dg, Gb, Jb, u, out, ws);                                                           \
    }
#define LANES_DEGREE_DEF(NM, NP) LANES_DEF(NM, NM, 1) LANES_DEF(NM, NM, 0) LANES_DEF(NM, NP, 0)
LANES_DEGREE_DEF(2, 3) LANES_DEGREE_DEF(3, 4) LANES_DEGREE_DEF(4, 5) LANES_DEGREE_DEF(5, 6)
LANES_DEGREE_DEF(6, 7) LANES_DEGREE_DEF(7, 8) LANES_DEGREE_DEF(8, 9) LANES_DEGREE_DEF(9, 10)

static batch_fn batch_kernel(int nm, int nq, int collocated)
{
#define LANES_CASE(NM, NQ, COLL) if (nm == NM && nq == NQ && (collocated != 0) == COLL) return batch_##NM##_##NQ##_##COLL;
#define LANES_DEGREE(NM, NP) LANES_CASE(NM, NM, 1) LANES_CASE(NM, NM, 0) LANES_CASE(NM, NP, 0)
    LANES_DEGREE(2, 3) LANES_DEGREE(3, 4) LANES_DEGREE(4, 5) LANES_DEGREE(5, 6)
    LANES_DEGREE(6, 7) LANES_DEGREE(7, 8) LANES_DEGREE(8, 9) LANES_DEGREE(9, 10)
    return NULL;
}

/* all cells [cb, ce) of one colour, LANES at a time (cells of a colour share no DoF, so the lanes' scatters are disjoint) */
static void colour_loop_lanes(batch_fn kernel, int nm, int nq, int flags, const double *S, const double *Dg, const double *G,
                              const double *JxW, const uint32_t *dof_indices, const uint32_t *color_cells, uint32_t cb, uint32_t ce,
                              const double *src, double *dst)
{
    const int L = LANES;
    const size_t nm3 = (size_t)nm * nm * nm, nq3 = (size_t)nq * nq * nq;
    const uint32_t n_batches = (ce - cb + L - 1) / L;
#pragma omp parallel
    {
        double *ws = (double *)aligned_alloc(64, sizeof(double) * L * (7 * nq3 + 2 * nm3 + 7 * nq3));
        double *u = ws + 7 * nq3 * L, *o = u + nm3 * L, *Gb = o + nm3 * L, *Jb = Gb + 6 * nq3 * L;
#pragma omp for schedule(static)
        for (uint32_t bi = 0; bi < n_batches; ++bi) {
            uint32_t cell[LANES];
            int cnt = 0;
            for (int l = 0; l < L; ++l) {
                const uint32_t ci = cb + bi * L + l;
                if (ci < ce) { cell[l] = color_cells ? color_cells[ci] : ci; cnt = l + 1; }
                else cell[l] = cell[0];  /* padding lane: computed, never scattered */
            }
            const uint32_t *idx[LANES];
            const double *Gc[LANES], *Jc[LANES];
            for (int l = 0; l < L; ++l) {
                idx[l] = dof_indices + (size_t)cell[l] * nm3;
                Gc[l] = (flags & OP_LAPLACE) ? G + (size_t)cell[l] * 6 * nq3 : NULL;
                Jc[l] = (flags & OP_MASS) ? JxW + (size_t)cell[l] * nq3 : NULL;
            }
            /* lane-interleave the batch's inputs: LANES sequential read streams, contiguous writes */
            for (size_t k = 0; k < nm3; ++k)
                for (int l = 0; l < L; ++l) u[k * L + l] = idx[l][k] == INVALID_INDEX ? 0.0 : src[idx[l][k]];
            if (flags & OP_LAPLACE)
                for (size_t k = 0; k < 6 * nq3; ++k)
                    for (int l = 0; l < L; ++l) Gb[k * L + l] = Gc[l][k];
            if (flags & OP_MASS)
                for (size_t k = 0; k < nq3; ++k)
                    for (int l = 0; l < L; ++l) Jb[k * L + l] = Jc[l][k];
            kernel(flags, S, Dg, Gb, Jb, u, o, ws);
            for (size_t k = 0; k < nm3; ++k)
                for (int l = 0; l < cnt; ++l)
                    if (idx[l][k] != INVALID_INDEX) dst[idx[l][k]] += o[k * L + l];
        }
        free(ws);
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
        /* coloured call (the timed baseline): lane-batched fast path with compile-time sizes; same arithmetic per cell */
        batch_fn kernel = color_offsets != NULL ? batch_kernel(nm, nq, collocated) : NULL;
        if (kernel) {
            colour_loop_lanes(kernel, nm, nq, flags, shape_values, co_shape_gradients, G, JxW, dof_indices, color_cells, cb, ce, src, dst);
            continue;
        }
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
