/*
 * b200fe.h -- C ABI of the B200-native matrix-free operator path.
 *
 * Drop-in boundary for the hot path of dealii-X/benchmarks (paths below are relative to the
 * reference tree).  Every entry point takes plain pointers and sizes, returns an int status
 * (0 = success) and never throws.  Device pointers are named d_*, host pointers h_*.  Every
 * call that launches work takes the CUDA stream (cudaStream_t passed as void*) and performs no
 * hidden synchronisation unless its comment says so.  There is no CPU fallback: without a
 * CUDA device every compute call returns B200FE_ERR_CUDA.
 *
 * Index / layout conventions are the reference's own:
 *   element-local DoF  l = i*nm^2 + j*nm + k, k fastest (= x in deal.II lexicographic order)
 *   quadrature point   p*nq^2 + q*nq + r      (p <-> i, the slowest index)
 *   G  [cell][6][nq^3] components rr,rs,rt,ss,st,tt, r = direction of the slowest index
 *        (CEED_BK/include/kernels/BK3/templated_cuda_kernels.cuh:156-161,
 *         CEED_bp/include/portable_laplace_operator.h:294-295)
 *   JxW [cell][nq^3]
 */
#ifndef B200FE_H
#define B200FE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200FE_VERSION 100 /* 0.1.0 */

enum {
    B200FE_OK = 0,
    B200FE_ERR_INVALID_ARG = 1, /* null pointer, size mismatch, bad enum */
    B200FE_ERR_UNSUPPORTED = 2, /* degree outside 1..8 or nq not in {p+1, p+2} */
    B200FE_ERR_CUDA = 3,        /* CUDA runtime error (text in b200fe_last_error) */
    B200FE_ERR_COMM = 4,        /* NCCL / peer-access error */
    B200FE_ERR_NO_CONVERGENCE = 5 /* CG hit max_it (deal.II SolverControl::NoConvergence) */
};

#define B200FE_INVALID_INDEX 0xFFFFFFFFu /* numbers::invalid_unsigned_int */

int b200fe_version(void);
/* Text of the last error raised on the calling thread ("" if none). */
const char *b200fe_last_error(void);

/* ------------------------------------------------------------------------------------------
 * 1. Element-vector bake-off kernels (replace the <<<>>> launches of the standalone drivers).
 *    h_basis  [nq*nm]  basis[q*nm+i]           (CEED_BK/src/BK1/templated_cuda_benchmark.cc:52-58)
 *    h_dbasis [nq*nq]  dbasis[p*nq+n]          (CEED_BK/src/BK3/templated_cuda_benchmark.cc:60-66)
 *    The two small 1-D matrices are HOST pointers: they ride to the GPU as kernel parameters
 *    (constant bank), which replaces the reference's d_basis/d_dbasis device copies.
 * ------------------------------------------------------------------------------------------ */

/* out_e = B^T (JxW .* (B in_e)).  Replaces BK1::Parallel::MassOperator<T,nq><<<>>>
 * (CEED_BK/include/kernels/BK1/templated_cuda_kernels.cuh:10-195; launch at
 *  CEED_BK/src/BK1/templated_cuda_benchmark.cc:84).  nq = p+2 (the reference fixes nm = nq-1). */
int b200fe_bk1_apply(int p, int nq, uint32_t nelmt, const double *h_basis, const double *d_JxW,
                     const double *d_in, double *d_out, void *stream);

/* out_e = B^T D^T G D B in_e.  Replaces BK3::Parallel::LaplaceOperator<T,nq><<<>>>
 * (CEED_BK/include/kernels/BK3/templated_cuda_kernels.cuh:11-292; launch at
 *  CEED_BK/src/BK3/templated_cuda_benchmark.cc:95). */
int b200fe_bk3_apply(int p, int nq, uint32_t nelmt, const double *h_basis, const double *h_dbasis,
                     const double *d_G, const double *d_in, double *d_out, void *stream);

/* out_e = D^T G D in_e, nm = nq = p+1.  Replaces BK5::Parallel::LaplaceOperator<T,nq><<<>>>
 * (CEED_BK/include/kernels/BK5/templated_cuda_kernels.cuh:11-126). */
int b200fe_bk5_apply(int p, uint32_t nelmt, const double *h_dbasis, const double *d_G,
                     const double *d_in, double *d_out, void *stream);

/* sum(x[i]^2) -> *d_result (device scalar, overwritten).  Replaces the thrust::transform_reduce
 * checksum of the drivers (CEED_BK/src/BK3/templated_cuda_benchmark.cc:104-108). */
int b200fe_sum_squares(uint64_t n, const double *d_x, double *d_result, void *stream);

/* Launch configuration actually used by the E-/L-vector kernels (for the benchmark tables that
 * print nelmtPerBatch / numBlocks / threadsPerBlock, CEED_BK/include/benchmark_printer.hpp:8-81).
 * kind: 1 = BK1, 3 = BK3, 5 = BK5. */
int b200fe_bk_launch_info(int kind, int p, int nq, uint32_t nelmt, int *elems_per_block,
                          int *num_blocks, int *threads_per_block, int *smem_bytes);

/* ------------------------------------------------------------------------------------------
 * 2. 1-D bases (what deal.II's QGauss<1>/FE_Q/MatrixFree::ShapeInfo hand to the reference operator;
 *    SURVEY.md appendix A7).  All arrays are HOST arrays, any output pointer may be NULL.
 *      h_shape_values       [nm*nq]  shape_values[i*nq+q]        (bk3_kokkos_kernel.h:161)
 *      h_co_shape_gradients [nq*nq]  co_shape_gradients[n*nq+q]  (bk3_kokkos_kernel.h:243)
 *      h_shape_gradients    [nm*nq]  d/dx of shape i at point q
 *      h_points, h_weights  [nq]     on [0,1], ascending, weights sum to 1
 *    quad_kind GLL with nq = p+1 is the collocated BP5 setting: shape_values is the identity.
 * ------------------------------------------------------------------------------------------ */
enum { B200FE_QUAD_GAUSS = 0, B200FE_QUAD_GLL = 1 };

int b200fe_basis_1d(int p, int nq, int quad_kind, double *h_shape_values,
                    double *h_co_shape_gradients, double *h_shape_gradients, double *h_points,
                    double *h_weights);

/* ------------------------------------------------------------------------------------------
 * 3. Box mesh + FE_Q(p) DoF numbering + partition (host).  Stands in for what the reference
 *    drivers obtain from deal.II: subdivided_hyper_rectangle + refine_global
 *    (CEED_bp/src/bp3.cc:452-488; bp5_kokkos/create_triangulation.h:17-29), distribute_dofs,
 *    Dirichlet constraints on the whole boundary (bp3.cc:147-151) and the per-cell index table
 *    of setup_dirichlet_boundary_dofs_masks (CEED_bp/include/portable_laplace_operator.h:304-394).
 *    Every rank builds its own view; no communication.
 * ------------------------------------------------------------------------------------------ */
enum { B200FE_PARTITION_P4EST = 0, /* first cell of rank r = floor(N r / P) on the z-order curve */
       B200FE_PARTITION_BLOCKS = 1 /* active_cell_index / ceil(N/P), create_triangulation.h:44-51 */ };
enum { B200FE_GHOSTS_MINIMAL = 0,  /* DoFs of owned cells that another rank owns */
       B200FE_GHOSTS_RELEVANT = 1  /* every DoF of the one-cell ghost layer (deal.II locally relevant set) */ };

typedef struct {
    int subdivisions[3]; /* coarse cells per axis */
    int n_refine;        /* refine_global(n_refine) */
    double p1[3], p2[3]; /* box corners */
    int p;               /* FE_Q degree 1..8 */
    int n_ranks, rank;
    int partition;       /* B200FE_PARTITION_* */
    int ghosts;          /* B200FE_GHOSTS_* */
    int dirichlet;       /* 1: constrain every boundary DoF (boundary id 0), 0: no constraints */
} b200fe_boxmesh_desc;

typedef struct {
    uint64_t n_cells_global, n_dofs_global;
    uint64_t first_cell;  /* active index of the first owned cell */
    uint64_t owned_begin; /* first owned global DoF */
    uint32_t n_cells_local, n_owned, n_ghost, n_constrained;
    uint32_t cells[3];
    double h[3];
} b200fe_boxmesh_info_t;

typedef struct b200fe_boxmesh b200fe_boxmesh;

int b200fe_boxmesh_create(const b200fe_boxmesh_desc *desc, b200fe_boxmesh **out);
void b200fe_boxmesh_destroy(b200fe_boxmesh *mesh);
int b200fe_boxmesh_info(const b200fe_boxmesh *mesh, b200fe_boxmesh_info_t *info);
/* Copies out (any pointer may be NULL):
 *   h_dof_indices   [n_cells_local][nm^3] partitioner-local index or B200FE_INVALID_INDEX
 *   h_constrained   [n_constrained]       owned local indices with Dirichlet constraints (sorted)
 *   h_ghost_global  [n_ghost]             global index of each ghost (sorted => grouped by owner)
 *   h_ghost_owner   [n_ghost]             owning rank of each ghost
 *   h_cell_xyz      [n_cells_local][3]    integer cell coordinates in iterator (z-order) order
 *   h_rank_dof_begin[n_ranks+1]           owned ranges of all ranks */
int b200fe_boxmesh_fill(const b200fe_boxmesh *mesh, uint32_t *h_dof_indices, uint32_t *h_constrained,
                        uint64_t *h_ghost_global, int32_t *h_ghost_owner, int32_t *h_cell_xyz,
                        uint64_t *h_rank_dof_begin);

#ifdef __cplusplus
}
#endif
#endif /* B200FE_H */
