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

/* Debugging aid for tests: overwrite the shared memory of every SM with NaN patterns, so that a kernel
 * reading shared memory it did not write (e.g. the unused slots of a tail batch) shows up in results. */
int b200fe_debug_poison_smem(void *stream);

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
    double h[3];      /* cell size per axis */
    double origin[3]; /* p1 */
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

/* The same index table written by a DEVICE kernel into d_dof_indices[n_cells_local][(p+1)^3] (what the operator borrows):
 * the mesh object keeps, per cell, the local index of the first DoF of each of its 27 entities (vertices, lines, quads,
 * interior); this uploads those 27 numbers per cell and expands them on the GPU instead of building the (p+1)^3 entries per
 * cell on the host and copying them over (360 MB for the 64^3-cell, p = 6 mesh).  Bit-identical to b200fe_boxmesh_fill.
 * Replaces the host loop of CEED_bp/include/portable_laplace_operator.h:304-394.  Synchronises `stream`. */
int b200fe_boxmesh_dof_indices_device(const b200fe_boxmesh *mesh, uint32_t *d_dof_indices, void *stream);

/* Mapping support points of the owned cells on the DEVICE, d_nodes[cell][3][(p_geo+1)^3]
 * (node index c*ng^2 + b*ng + a, a <-> x), Gauss-Lobatto lattice of MappingQ(p_geo).
 * deform_kind 0: the box itself; 1: x += A sin(f y), y += A sin(f z), z += A sin(f x)
 * (smooth deformed mesh in the spirit of bk3_dealii/check_bk3.cc:50-52).  Synchronises `stream`. */
int b200fe_boxmesh_nodes(const b200fe_boxmesh *mesh, int p_geo, int deform_kind, double amplitude,
                         double frequency, double *d_nodes, void *stream);

/* ------------------------------------------------------------------------------------------
 * 3b. Box mesh with ONE extra level of local refinement (hanging nodes; BASELINE config C5).
 *    The reference has no such mesh; a deal.II driver would get these objects from
 *    Triangulation::execute_coarsening_and_refinement + DoFHandler::distribute_dofs +
 *    DoFTools::make_hanging_node_constraints + AffineConstraints.  Conventions: csrc/hangmesh.cc (H1-H5).
 *    The cells [refine_lo, refine_hi) of the box mesh (coordinates after refine_global) are refined once.
 *    Only B200FE_PARTITION_P4EST and B200FE_GHOSTS_MINIMAL are built.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    b200fe_boxmesh_desc box;
    int refine_lo[3], refine_hi[3];
} b200fe_hangmesh_desc;

typedef struct {
    uint64_t n_cells_global, n_dofs_global;
    uint64_t first_cell;  /* position of the first owned cell on the p4est curve */
    uint64_t owned_begin;
    uint32_t n_cells_local, n_owned, n_ghost, n_constrained;
    uint32_t n_hanging_rows, n_hanging_entries;
    uint32_t n_face_blocks; /* the same rows grouped by coarse face, see b200fe_hangmesh_fill_faces */
    uint32_t cells[3]; /* of the unrefined box mesh */
    double h[3];       /* cell size of the unrefined box mesh */
    double origin[3];
} b200fe_hangmesh_info_t;

typedef struct b200fe_hangmesh b200fe_hangmesh;

int b200fe_hangmesh_create(const b200fe_hangmesh_desc *desc, b200fe_hangmesh **out);
void b200fe_hangmesh_destroy(b200fe_hangmesh *mesh);
int b200fe_hangmesh_info(const b200fe_hangmesh *mesh, b200fe_hangmesh_info_t *info);
/* Copies out (any pointer may be NULL); the first six arrays as in b200fe_boxmesh_fill except
 *   h_dof_indices   hanging DoFs are ordinary entries (their values come from b200fe_op_set_constraints)
 *   h_constrained   owned Dirichlet AND owned hanging DoFs (rows that act as identity)
 *   h_cell_lxyz     [n_cells_local][4]  level (0 = unrefined, 1 = child), x, y, z at the cell's own level
 * and the hanging-node rows this rank needs (every hanging DoF of its cells, owned or ghost), CSR with
 * partitioner-local indices, Dirichlet parents dropped:
 *   h_hang_dof[n_hanging_rows], h_hang_row_ptr[n_hanging_rows+1], h_hang_col / h_hang_w[n_hanging_entries]:
 *   u[hang_dof[r]] = sum_k hang_w[k] * u[hang_col[k]],  k in [row_ptr[r], row_ptr[r+1]) */
int b200fe_hangmesh_fill(const b200fe_hangmesh *mesh, uint32_t *h_dof_indices, uint32_t *h_constrained,
                         uint64_t *h_ghost_global, int32_t *h_ghost_owner, int32_t *h_cell_lxyz,
                         uint64_t *h_rank_dof_begin, uint32_t *h_hang_dof, uint32_t *h_hang_row_ptr,
                         uint32_t *h_hang_col, double *h_hang_w);
/* Face-structured form of the same rows.  Every hanging DoF of such a mesh lies on a face shared by an unrefined
 * cell K and a refined cell, and the (2p+1)^2 fine nodes of the face are the tensor-product interpolation
 *   u_fine(a', b') = sum_{a,b} W[a'][a] W[b'][b] u_K(a, b),   W = b200fe_trace_weights (row-major [2p+1][p+1])
 * of the (p+1)^2 face nodes of K (a, a' run along the lower in-face axis).  Per block, partitioner-local indices:
 *   h_face_parents [n_face_blocks][(p+1)^2]   B200FE_INVALID_INDEX = Dirichlet / absent parent (value 0)
 *   h_face_children[n_face_blocks][(2p+1)^2]  B200FE_INVALID_INDEX = not a child of this block (coarse vertex,
 *        Dirichlet, not needed by this rank, or assigned to an earlier block: every hanging DoF appears once)
 * 370 scattered accesses per face at p = 8 instead of the 23,400 of its CSR rows. */
int b200fe_hangmesh_fill_faces(const b200fe_hangmesh *mesh, uint32_t *h_face_parents, uint32_t *h_face_children);
/* 1-D trace matrix of FE_Q(p): W[rel][j] = l_j((half + t_a)/2) at the fine lattice positions rel = half*p + a of a
 * coarse interval (t = Gauss-Lobatto nodes); exact unit rows where a fine node coincides with a coarse node. */
int b200fe_trace_weights(int p, double *h_W);
/* As b200fe_boxmesh_nodes, for the cells of this mesh (children are half the size). */
int b200fe_hangmesh_nodes(const b200fe_hangmesh *mesh, int p_geo, int deform_kind, double amplitude,
                          double frequency, double *d_nodes, void *stream);

/* ------------------------------------------------------------------------------------------
 * 4. Geometry and the L-vector operator (drop-in for Portable::LaplaceOperator,
 *    CEED_bp/include/portable_laplace_operator.h:17-96, and for bp5_kokkos' HelmholtzOperator,
 *    bp5_kokkos/benchmark.cc:141-292).
 * ------------------------------------------------------------------------------------------ */

/* compute_G_tensors (CEED_bp/include/portable_laplace_operator.h:239-302) from mapping support
 * points: d_G[cell][6][nq^3] = JxW * K K^T with K = J^-1 and the reference directions ordered
 * (r,s,t) = (z^,y^,x^) as the cell kernel pairs them, i.e. the mathematically consistent form
 * of bakeoff_problems_dealii/include/portable_laplace_operator.h:227-258 (on the reference's cube
 * cells both coincide: G = diag(h w_q)).  d_JxW[cell][nq^3] = det J * w.  Either output may be NULL. */
int b200fe_geometry_from_nodes(int p_geo, int nq, int quad_kind, uint32_t n_cells, const double *d_nodes,
                               double *d_G, double *d_JxW, void *stream);

/* The same factors from deal.II's own MatrixFree data (what the deal.II-kernel variant of the operator reads,
 * bakeoff_problems_dealii/include/portable_laplace_operator.h:227-258): d_inv_jacobian(q, cell, ref, real) =
 * d xi_ref / d x_real and d_JxW(q, cell) as Kokkos LayoutLeft views, i.e. element (q, cell, ref, real) at
 * q + nq^3 * (cell + n_cells * (ref + 3 * real)), q = x + nq (y + nq z).  d_G[cell][6][nq^3] as above. */
int b200fe_geometry_from_inv_jacobian(uint32_t n_cells, int nq, const double *d_inv_jacobian, const double *d_JxW,
                                      double *d_G, void *stream);

/* Per-cell constants for the on-the-fly geometry of affine (parallelepiped) cells from MappingQ1 support points
 * d_nodes[cell][3][2][2][2] (b200fe_boxmesh_nodes with p_geo = 1): d_cell_G[cell][8], see b200fe_op_desc. */
int b200fe_geometry_affine_from_nodes(uint32_t n_cells, const double *d_nodes, double *d_cell_G, void *stream);

enum { B200FE_OP_LAPLACE = 1, B200FE_OP_MASS = 2, B200FE_OP_HELMHOLTZ = 3 };

typedef struct {
    int p;          /* fe_degree 1..8 */
    int nq;         /* 1-D quadrature points: p+2 (BP1/BP3), p+1 ("bp35", bp5_kokkos, BP5) */
    int op_kind;    /* B200FE_OP_* */
    int collocated; /* 1: quadrature points are the GLL nodes (CEED BP5); shape_values ignored */
    uint32_t n_cells, n_owned, n_ghost;
    const double *h_shape_values;       /* [nm*nq] shape_values[i*nq+q]       (MatrixFree data) */
    const double *h_co_shape_gradients; /* [nq*nq] co_shape_gradients[n*nq+q] (MatrixFree data) */
    const uint32_t *d_dof_indices; /* DEVICE [n_cells][nm^3], B200FE_INVALID_INDEX = constrained; BORROWED */
    const double *d_G;             /* DEVICE [n_cells][6][nq^3]; BORROWED (may be NULL for MASS) */
    const double *d_JxW;           /* DEVICE [n_cells][nq^3]; BORROWED (MASS/HELMHOLTZ, rhs_one) */
    const uint32_t *h_constrained; /* HOST owned local indices of constrained DoFs; copied */
    uint32_t n_constrained;
    /* overlap split (deal.II colours with overlap_communication_computation = true): cells
     * [0,n_phase0) and [n_phase0+n_phase1, n_cells) touch no ghost DoF, cells of phase 1 may.
     * 0,0 = single colour, no overlap (the reference drivers' setting, bp3.cc:103). */
    uint32_t n_phase0, n_phase1;
    /* Optional: geometry evaluated on the fly for AFFINE cells ("next" row): when
     * d_cell_G != NULL the Laplace kernels do not stream d_G (48 nq^3 bytes per cell) but use
     * G(q) = cell_G * w_p w_q w_r with DEVICE d_cell_G[cell][8] = {det J * K K^T (rr,rs,rt,ss,st,tt), det J, 0}
     * (b200fe_geometry_affine_from_nodes) and the HOST 1-D weights h_weights[nq].  Results equal the
     * stored-G path to rounding on parallelepiped cells; algorithmic bytes drop to 4 nm^3 + 64 per cell.
     * d_G may still be given (needed by b200fe_op_diagonal on parallelepiped cells).  When every cell turns out to be an
     * axis-aligned box (checked on the device; b200fe_op_cartesian) the separable kernels run instead, and then
     *  - the mass and Helmholtz operators are accepted with d_cell_G as well (their mass term is det J (M x M x M); d_JxW
     *    may be NULL), on any other cell shape they are refused with B200FE_ERR_UNSUPPORTED;
     *  - b200fe_op_diagonal and b200fe_op_rhs_one need neither d_G nor d_JxW (unconstrained operators).  BORROWED. */
    const double *d_cell_G;
    const double *h_weights;
    /* Optional: geometry evaluated on the fly for TRILINEAR cells (general hexahedra, the MappingQ1 case of SURVEY.md 8f.1):
     * when d_cell_vertices != NULL the Laplace kernels rebuild the Jacobian at every quadrature point from the DEVICE array
     * d_cell_vertices[cell][3][2][2][2] (coordinate d, then the vertex index z, y, x with x fastest -- b200fe_boxmesh_nodes
     * with p_geo = 1) and form G = JxW J^-1 J^-T in registers; needs h_weights and the HOST 1-D points h_points[nq].
     * 192 bytes per cell instead of 48 nq^3; results equal the stored-G operator of the same MappingQ1 mesh to rounding.
     * Built for operators with the symmetric 1-D matrices of a real basis (b200fe_basis_1d).  BORROWED. */
    const double *d_cell_vertices;
    const double *h_points;
} b200fe_op_desc;

typedef struct b200fe_op b200fe_op;
typedef struct b200fe_halo b200fe_halo;

/* Borrowed device arrays must outlive the operator (the reference operator owns G and the masks;
 * here the caller's vectors are used in place so a 4 GB G is never duplicated). */
int b200fe_op_create(const b200fe_op_desc *desc, b200fe_op **out);
void b200fe_op_destroy(b200fe_op *op);
/* Attach the ghost exchange (NULL detaches).  Borrowed. */
int b200fe_op_set_halo(b200fe_op *op, b200fe_halo *halo);

/* Hanging-node constraints (AffineConstraints with homogeneous, chain-free rows; b200fe_hangmesh_fill):
 * the operator becomes C^T A C with identity on the constrained rows.  Inside every vmult-type call, after
 * the ghost update, u[hang_dof[r]] = sum_k w[k] u[col[k]] is written into src (the hanging entries of src
 * are scratch for the duration of the call and restored afterwards, like its ghost entries); after the cell
 * kernel, dst[col[k]] += w[k] dst[hang_dof[r]] and dst[hang_dof[r]] = 0 before compress(add).  Hanging DoFs
 * must also be listed in h_constrained of the descriptor.  The lists are copied.  n_rows = 0 detaches.
 * The overlap split (n_phase0/n_phase1) is ignored while constraints are attached. */
int b200fe_op_set_constraints(b200fe_op *op, uint32_t n_rows, const uint32_t *h_hang_dof,
                              const uint32_t *h_hang_row_ptr, const uint32_t *h_hang_col, const double *h_hang_w);
/* The same constraints in the face-structured form of b200fe_hangmesh_fill_faces: one CTA per coarse face applies
 * W (x) W (and its transpose) in shared memory -- (p+1)^2 + (2p+1)^2 scattered accesses per face instead of their
 * product.  The fast path for meshes whose hanging nodes come from one level of refinement (BP6 p = 8, 48 M DoFs:
 * 1.06 ms per apply against 1.16 ms with the CSR rows and 1.03 ms without constraints); the CSR rows remain the
 * general AffineConstraints path.  When set (n_blocks > 0) it replaces the rows inside vmult / distribute / rhs;
 * h_W = b200fe_trace_weights(p).  Copied. */
int b200fe_op_set_face_constraints(b200fe_op *op, int p, uint32_t n_blocks, const uint32_t *h_face_parents,
                                   const uint32_t *h_face_children, const double *h_W);
/* AffineConstraints::distribute on a local vector: fills the hanging entries of d_x from their parents
 * (call after the solve; ghost entries of d_x must be up to date when parents are ghosts). */
int b200fe_op_distribute(b200fe_op *op, double *d_x, void *stream);

/* LaplaceOperator::vmult (portable_laplace_operator.h:124-172): dst = 0; update ghosts of src;
 * cell kernel (gather / sum factorisation / atomic scatter); compress(add); zero ghosts of src;
 * dst[c] = src[c] on constrained DoFs.  Vectors hold n_owned + n_ghost doubles; the ghost entries
 * of src are scratch (as with deal.II's mutable ghost section). */
int b200fe_op_vmult(b200fe_op *op, double *d_dst, const double *d_src, void *stream);
/* Vector-valued problems (CEED BP2/BP4/BP6: the scalar operator applied to each of n_components
 * components): vectors are component-blocked, [component][n_owned + n_ghost].  One cell-kernel launch
 * per component (a fused multi-component kernel that streams G once is the planned next step). */
int b200fe_op_vmult_components(b200fe_op *op, int n_components, double *d_dst, const double *d_src, void *stream);
/* vmult with the inner product src . dst (summed over ranks) fused into the kernel -> *d_dot. */
int b200fe_op_vmult_dot(b200fe_op *op, double *d_dst, const double *d_src, double *d_dot, void *stream);
/* LaplaceOperator::vmult_dummy (portable_laplace_operator.h:175-235). */
int b200fe_op_vmult_dummy(b200fe_op *op, double *d_dst, const double *d_src, int ghost_exchange_on,
                          int computation_on, void *stream);
/* vmult with HOST vectors of n_owned doubles (H2D, apply, D2H; synchronises the stream). */
int b200fe_op_vmult_host(b200fe_op *op, double *h_dst, const double *h_src, void *stream);
/* HelmholtzOperator::compute_diagonal (bp5_kokkos/benchmark.cc:218-251): matrix diagonal, 1 on
 * constrained rows.  (The reciprocal of :240-250 is one elementwise op on the caller's side.)
 * With hanging-node constraints attached the result is diag(C^T A C) (MatrixFreeTools::compute_diagonal semantics): plain
 * cell diagonals on the cells without hanging DoFs; on the others the columns of the cell matrix (cell kernel on unit
 * vectors) with the local constraint matrix applied on both sides.  Needs the CSR rows (b200fe_op_set_constraints) on the
 * operator, also when the face-structured form is the one used by the apply; setup path, synchronises. */
int b200fe_op_diagonal(b200fe_op *op, double *d_diag, void *stream);
/* compute_rhs of the BP drivers (CEED_bp/src/bp3.cc:184-239): b_i = int phi_i * 1, constrained rows 0. */
int b200fe_op_rhs_one(b200fe_op *op, double *d_b, void *stream);
/* Per-launch CUDA-event timing of the cell kernel (what bench.py's roofline reads): enable with room
 * for max_launches launches (0 disables); read returns the summed kernel time and the number of
 * launches recorded since the last read (waits for them to finish). */
int b200fe_op_timing_enable(b200fe_op *op, int max_launches);
int b200fe_op_timing_read(b200fe_op *op, double *total_ms, int *launches);
/* Number of kernels this library has launched in this process (cell kernels, CG vector kernels, ...). */
unsigned long long b200fe_launch_count(void);
int b200fe_op_launch_info(b200fe_op *op, int *elems_per_block, int *num_blocks, int *threads_per_block,
                          int *smem_bytes, int *blocks_per_sm, int *regs_per_thread);
/* Which contraction form the cell kernel of this operator runs: *even_odd = 1 when the operator's 1-D matrices have the
 * point symmetry of a real basis and the even-odd kernel is built for its (degree, quadrature) -- deal.II's
 * evaluate_evenodd, halving the multiply-adds of every 1-D sweep -- else 0 (plain contractions; always the case for the
 * reference drivers' cos() test matrices, CEED_BK/src/BK3/templated_cuda_benchmark.cc:50-66).  B200FE_EVEN_ODD=0 in the
 * environment forces 0. */
int b200fe_op_kernel_variant(b200fe_op *op, int *even_odd);
/* *on = 1 when the scatter of this operator writes the DoFs at the interior positions of a cell (all three local indices in
 * 1..p-1) with plain stores instead of atomic adds: for FE_Q those DoFs belong to one cell only, which b200fe_op_create
 * VERIFIES on the dof_indices table it is given (one device pass; a table that breaks the property keeps atomics
 * everywhere, *on = 0).  The CG loop then skips their zero-fill as well: (p-1)^3 / p^3 of the result vector is written
 * once per apply instead of zeroed, fetched into L2 and added to.  p >= 3 only; B200FE_EXCL_INTERIOR=0 forces 0.
 * (deal.II's matrix-free loops get the same effect from their cell-interior "pre/post" ranges; no reference file -- the
 * reference kernels scatter everything with Kokkos::atomic_add, CEED_bp/include/bk3_kokkos_kernel.h:372-386.) */
int b200fe_op_exclusive_interior(b200fe_op *op, int *on);
/* *on = 1 when the operator runs the separable "cartesian" kernel: collocated Laplace operator with on-the-fly affine
 * geometry (d_cell_G) whose cells are ALL axis-aligned boxes (checked on the device at b200fe_op_create) -- deal.II
 * MatrixFree's cartesian cell type, and what every mesh of the reference drivers is (CEED_bp/src/bp3.cc:452-488: cube cells).
 * G is then diagonal and separable, and D^T G D u = c_rr (S x W x W) u + c_ss (W x S x W) u + c_tt (W x W x S) u with the
 * 1-D stiffness matrix S = D^T W D: three 1-D contractions per point instead of six.  B200FE_CARTESIAN=0 keeps the general
 * affine kernel. */
int b200fe_op_cartesian(b200fe_op *op, int *on);

/* ------------------------------------------------------------------------------------------
 * 5. Conjugate gradients (dealii::SolverCG + ReductionControl as called at CEED_bp/src/bp3.cc:266-285
 *    and bp5_kokkos/benchmark.cc:355-378).  x0 = 0.  d_inv_diag = NULL: PreconditionIdentity,
 *    else Jacobi with the given inverse diagonal.  Convergence is tested on the device every
 *    iteration; the host polls it every `check_every` iterations (iterations queued after
 *    convergence are no-ops, so the result and the iteration count are those of deal.II's loop).
 *    Returns B200FE_ERR_NO_CONVERGENCE (result still filled) when max_it is reached.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    int iterations; /* ReductionControl::last_step() */
    int converged;
    double initial_residual, final_residual; /* initial_value(), last_value() */
} b200fe_cg_result;

int b200fe_cg_solve(b200fe_op *op, double *d_x, const double *d_b, const double *d_inv_diag, double abs_tol,
                    double rel_tol, int max_it, int check_every, b200fe_cg_result *result, void *stream);
/* Vector-valued problems (CEED BP2/BP4/BP6): one CG on the block-diagonal system of n_components copies of the
 * scalar operator; x and b are component-blocked, [component][n_owned + n_ghost] (b200fe_op_vmult_components);
 * d_inv_diag (n_owned entries, or NULL) is shared by the components. */
int b200fe_cg_solve_components(b200fe_op *op, int n_components, double *d_x, const double *d_b, const double *d_inv_diag,
                               double abs_tol, double rel_tol, int max_it, int check_every, b200fe_cg_result *result,
                               void *stream);
/* CG preconditioned by a Chebyshev polynomial of the Jacobi-scaled operator -- dealii::PreconditionChebyshev<Operator, Vector,
 * DiagonalMatrix> as a SolverCG preconditioner (the step after Jacobi; the reference includes the multigrid transfer header for
 * it, CEED_bp/src/bp3.cc:27, without using it yet).  z = p_k(D^-1 A) D^-1 r with the polynomial of `degree` terms that is
 * optimal on [lambda_max / smoothing_range, lambda_max] (three-term recurrence, degree - 1 operator applications per CG
 * iteration).  lambda_max: b200fe_op_estimate_max_eigenvalue.  Scalar problems (one component). */
int b200fe_cg_solve_chebyshev(b200fe_op *op, double *d_x, const double *d_b, const double *d_inv_diag, int degree, double lambda_max,
                              double smoothing_range, double abs_tol, double rel_tol, int max_it, int check_every,
                              b200fe_cg_result *result, void *stream);
/* p-multigrid (polynomial global coarsening on the same cells): deal.II's MGTransferGlobalCoarsening between FE_Q(p_fine) and
 * FE_Q(p_coarse) -- the header the reference already includes for this step (CEED_bp/src/bp3.cc:27) -- and a V-cycle with
 * PreconditionChebyshev smoothers as a SolverCG preconditioner.
 *   ptransfer: the two levels' index tables [cell][(p+1)^3] (BORROWED; Dirichlet DoFs masked on both) and local vector lengths;
 *     prolongate_add: fine += P coarse (cell contributions weighted by 1 / valence); restrict_add: coarse += P^T fine.
 *   pmg: n_levels operators from fine to coarse on the SAME cells (single rank), their inverse diagonals and
 *     b200fe_op_estimate_max_eigenvalue values, the n_levels - 1 transfers; smoother_degree terms of the Chebyshev polynomial
 *     before and after the coarse correction, coarse_degree terms on the coarsest level.  vcycle: z = V(r) from a zero start
 *     (n_owned entries of the finest level); cg_solve_pmg: SolverCG on the finest operator with that preconditioner. */
typedef struct b200fe_ptransfer b200fe_ptransfer;
typedef struct b200fe_pmg b200fe_pmg;
int b200fe_ptransfer_create(int p_fine, int p_coarse, uint32_t n_cells, const uint32_t *d_idx_fine, const uint32_t *d_idx_coarse,
                            uint32_t n_local_fine, uint32_t n_local_coarse, b200fe_ptransfer **out);
void b200fe_ptransfer_destroy(b200fe_ptransfer *t);
int b200fe_ptransfer_prolongate_add(b200fe_ptransfer *t, double *d_fine, const double *d_coarse, void *stream);
int b200fe_ptransfer_restrict_add(b200fe_ptransfer *t, double *d_coarse, const double *d_fine, void *stream);
int b200fe_pmg_create(int n_levels, b200fe_op *const *ops, const double *const *d_inv_diag, const double *lambda_max,
                      b200fe_ptransfer *const *transfers, int smoother_degree, double smoothing_range, int coarse_degree,
                      b200fe_pmg **out);
void b200fe_pmg_destroy(b200fe_pmg *pmg);
int b200fe_pmg_vcycle(b200fe_pmg *pmg, double *d_z, const double *d_r, void *stream);
int b200fe_cg_solve_pmg(b200fe_pmg *pmg, double *d_x, const double *d_b, double abs_tol, double rel_tol, int max_it, int check_every,
                        b200fe_cg_result *result, void *stream);
/* Largest eigenvalue of D^-1 A by n_iterations power iterations from a fixed pseudo-random start vector, times deal.II's
 * safety factor 1.2 (PreconditionChebyshev::estimate_eigenvalues).  d_inv_diag = NULL: of A itself.  Synchronises. */
int b200fe_op_estimate_max_eigenvalue(b200fe_op *op, const double *d_inv_diag, int n_iterations, double *lambda_max, void *stream);
/* Same with HOST x and b (n_owned doubles): H2D of b, solve, D2H of x, synchronises. */
int b200fe_cg_solve_host(b200fe_op *op, double *h_x, const double *h_b, const double *d_inv_diag, double abs_tol,
                         double rel_tol, int max_it, int check_every, b200fe_cg_result *result, void *stream);

/* ------------------------------------------------------------------------------------------
 * 6. Ghost exchange between the GPUs of one node (NCCL over NVLink/NVSwitch; one process per GPU).
 *    Replaces update_ghost_values / compress(add) / zero_out_ghost_values
 *    (portable_laplace_operator.h:133,169,170) and p-halox's exchange round (p-halox/phalox.cc:104-126).
 *    Lists follow deal.II's Partitioner: the ghost segment is grouped by owner rank; the send list
 *    holds, per peer, the owned local indices that peer ghosts, in the peer's ghost order.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    int rank, n_ranks;
    const char *nccl_unique_id; /* 128 bytes from b200fe_comm_unique_id on rank 0, broadcast by the caller */
    uint32_t n_owned, n_ghost;
    int n_peers;
    const int *peers;
    const uint32_t *recv_offset, *recv_count; /* per peer: slice of the ghost segment (relative to n_owned) */
    const uint32_t *send_offset, *send_count; /* per peer: slice of h_send_indices */
    const uint32_t *h_send_indices;           /* owned local indices to pack */
    uint32_t n_send;
} b200fe_halo_desc;

/* The peer tables of b200fe_halo_desc for this rank, computed WITHOUT communication: the meshes above are replayable
 * on every rank, so the rank builds every other rank's view and reads off who ghosts which of its DoFs (deal.II's
 * Partitioner needs a message round for this).  desc->rank / desc->box.rank = this rank.
 * Arrays: h_peers/recv_offset/recv_count/send_offset/send_count [n_peers], h_send_indices [n_send]. */
typedef struct b200fe_exchange b200fe_exchange;
int b200fe_exchange_create_box(const b200fe_boxmesh_desc *desc, b200fe_exchange **out);
int b200fe_exchange_create_hang(const b200fe_hangmesh_desc *desc, b200fe_exchange **out);
void b200fe_exchange_destroy(b200fe_exchange *ex);
int b200fe_exchange_info(const b200fe_exchange *ex, int *n_peers, uint32_t *n_send, uint32_t *n_owned, uint32_t *n_ghost);
int b200fe_exchange_fill(const b200fe_exchange *ex, int32_t *h_peers, uint32_t *h_recv_offset, uint32_t *h_recv_count,
                         uint32_t *h_send_offset, uint32_t *h_send_count, uint32_t *h_send_indices);

int b200fe_comm_available(void); /* 1 if libnccl.so.2 could be bound */
int b200fe_comm_unique_id(char *id128);
/* Collective over all ranks (ncclCommInitRank). */
int b200fe_halo_create(const b200fe_halo_desc *desc, b200fe_halo **out);
void b200fe_halo_destroy(b200fe_halo *halo);
int b200fe_halo_update_ghosts(b200fe_halo *halo, double *d_v, void *stream);
int b200fe_halo_compress_add(b200fe_halo *halo, double *d_v, void *stream);
int b200fe_halo_zero_ghosts(b200fe_halo *halo, double *d_v, void *stream);
int b200fe_halo_allreduce_sum(b200fe_halo *halo, double *d_vals, int count, void *stream);
/* Transport of the calls above.  b200fe_halo_create maps one CUDA-IPC window per peer when every rank can reach every
 * other one by load/store (one NVLink/NVSwitch domain; B200FE_HALO_P2P=0 disables): exchanges are then peer stores + flags
 * issued by the pack kernel itself and the all-reduce is an all-gather of partial sums added in rank order (identical bits
 * on every rank) -- the "CUDA P2P over NVLink" transport next to "NCCL send/recv" (p-halox/README.md:62-68 compares MPI
 * transports the same way).  *p2p_available: windows exist; *p2p_in_use: current choice.  set_transport switches between
 * the two (collective: every rank must make the same choice before the next exchange).  b200fe_halo_status synchronises
 * the device and returns B200FE_ERR_COMM if a bounded wait of the P2P transport has expired (a peer never answered). */
int b200fe_halo_transport(b200fe_halo *halo, int *p2p_available, int *p2p_in_use);
int b200fe_halo_set_transport(b200fe_halo *halo, int use_p2p);
int b200fe_halo_status(b200fe_halo *halo);
/* One exchange round of p-halox (p-halox/phalox.cc:111-125: Irecv all, Isend all, Waitall) without
 * pack lists: per peer k, d_send[send_offset[k] .. +send_count[k]) goes to peers[k] and
 * d_recv[recv_offset[k] .. +recv_count[k]) is filled from it.  A halo created with
 * h_send_indices = NULL ("raw mode") supports only this call; the same peer may appear twice. */
int b200fe_halo_exchange_raw(b200fe_halo *halo, const double *d_send, double *d_recv, void *stream);
/* n_rounds consecutive rounds of the same exchange.  With the P2P transport and messages small enough for the low-latency
 * round (< 4096 doubles per direction on every rank) all rounds run inside ONE kernel launch -- the exchange as a
 * device-resident solver loop would issue it: a round then costs the NVLink hop and the polling instead of the launch-to-launch
 * gap of dependent kernels (p-halox "launch" mode: the latency floor of the fabric).  Otherwise the same as n_rounds calls of
 * b200fe_halo_exchange_raw. */
int b200fe_halo_exchange_raw_rounds(b200fe_halo *halo, const double *d_send, double *d_recv, int n_rounds, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* B200FE_H */
