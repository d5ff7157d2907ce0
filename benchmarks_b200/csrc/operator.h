// operator.h -- internal view of the L-vector operator object behind b200fe_op.
#pragma once
#include <vector>

#include "common.h"
#include "kernels.h"

namespace b200fe {

struct Halo;  // halo.cu

struct Operator {
    int p = 0, nm = 0, nq = 0, qop = 0;
    bool collocated = false;
    uint32_t n_cells = 0, n_owned = 0, n_ghost = 0, n_constrained = 0;
    // 1-D matrices in the kernel (BK) layout: B[q*nm+i], D[p*nq+n]; deal.II-layout copies for setup kernels
    std::vector<double> B, D, shape_values, co_shape_gradients;
    const uint32_t *d_idx = nullptr;  // borrowed
    const double *d_G = nullptr;      // borrowed
    const double *d_JxW = nullptr;    // borrowed
    const double *d_cellG = nullptr;  // borrowed; affine on-the-fly geometry: [cell][8]
    const double *d_cellX = nullptr;  // borrowed; trilinear on-the-fly geometry: [cell][3][2][2][2] vertex coordinates
    std::vector<double> W;            // 1-D quadrature weights (on-the-fly geometry), followed by the points (trilinear)
    // affine cells that are all axis-aligned boxes, collocated operator: the separable kernel (QOP_CARTESIAN) with the 1-D
    // stiffness matrix S = D^T W D (BK layout, in place of B); decided by b200fe_op_create from the per-cell constants
    bool cartesian = false;
    std::vector<double> S;     // K | M: the 1-D stiffness and mass matrices on the nodal basis, nm*nm doubles each
    std::vector<double> mvec;  // m = B^T w (int of the 1-D shape functions): rhs of the separable operators
    int otf_flag() const { return d_cellG ? (cartesian ? QOP_AFFINE | QOP_CARTESIAN : QOP_AFFINE) : d_cellX ? QOP_TRILINEAR : 0; }
    const double *otf_data() const { return d_cellG ? d_cellG : d_cellX; }
    int otf_stride() const { return d_cellG ? 8 : 24; }
    uint32_t *d_constrained = nullptr;  // owned
    double *d_mats = nullptr;           // owned: shape_values | co_shape_gradients | shape_gradients (setup kernels)
    // the same constraints grouped by coarse face (b200fe_op_set_face_constraints); takes precedence when set
    uint32_t n_face_blocks = 0;
    int face_nm = 0, face_nf = 0, face_save_comps = 0;
    uint32_t *d_face_parents = nullptr, *d_face_children = nullptr;
    double *d_face_W = nullptr, *d_face_save = nullptr;
    bool has_constraints() const { return n_hang != 0 || n_face_blocks != 0; }
    Halo *halo = nullptr;               // borrowed, optional
    // hanging-node rows (b200fe_op_set_constraints), owned copies: u[hang_dof[r]] = sum_k w[k] u[col[k]]
    uint32_t n_hang = 0;
    uint32_t *d_hang_dof = nullptr, *d_hang_ptr = nullptr, *d_hang_col = nullptr;
    double *d_hang_w = nullptr, *d_hang_save = nullptr;  // save: src values of the hanging entries during a vmult
    int hang_save_comps = 0;                              // components the save buffer has room for
    std::vector<uint32_t> h_hang_dof, h_hang_ptr, h_hang_col;  // host copies of the rows (compute_diagonal with constraints)
    std::vector<double> h_hang_w;
    // exclusive cell-interior DoFs (KArgs::excl_interior): bit i set = local DoF i sits at an interior position of exactly one
    // cell -- the cell kernel writes it with a plain store, so the zero-fill before an apply may skip it.  Owned; null when
    // the property does not hold on this index table (or p <= 2, or B200FE_EXCL_INTERIOR=0).
    uint32_t *d_excl_mask = nullptr;
    const int *d_skip = nullptr;        // set by the CG driver for the duration of a solve (see KArgs::skip)
    // overlap split: cells [0, n_phase0) and [n_phase0+n_phase1, n_cells) touch no ghost DoF
    uint32_t n_phase0 = 0, n_phase1 = 0;
    LaunchInfo last_launch{};
    // optional per-launch timing of the cell kernel (bench.py roofline): ring of event pairs
    bool timing = false;
    std::vector<cudaEvent_t> ev;   // 2 per launch
    size_t ev_used = 0;

    uint32_t n_local() const { return n_owned + n_ghost; }
    Operator() = default;
    Operator(const Operator &) = delete;
    Operator &operator=(const Operator &) = delete;
    ~Operator()
    {   // owned device resources only; borrowed arrays stay with the caller
        for (cudaEvent_t e : ev) cudaEventDestroy(e);
        cudaFree(d_constrained);
        cudaFree(d_mats);
        cudaFree(d_excl_mask);
        free_constraints();
        free_face_constraints();
    }
    void free_face_constraints()
    {
        cudaFree(d_face_parents); cudaFree(d_face_children); cudaFree(d_face_W); cudaFree(d_face_save);
        d_face_parents = d_face_children = nullptr;
        d_face_W = d_face_save = nullptr;
        n_face_blocks = 0;
        face_save_comps = 0;
    }
    void free_constraints()
    {
        cudaFree(d_hang_dof); cudaFree(d_hang_ptr); cudaFree(d_hang_col); cudaFree(d_hang_w); cudaFree(d_hang_save);
        d_hang_dof = d_hang_ptr = d_hang_col = nullptr;
        d_hang_w = d_hang_save = nullptr;
        n_hang = 0;
        hang_save_comps = 0;
        h_hang_dof.clear(); h_hang_ptr.clear(); h_hang_col.clear(); h_hang_w.clear();
    }
};

// dst = 0 on the local vector, cell kernel over cells [cell_begin, cell_end), optional fused dot.
// ncomp > 1: all components of component-blocked vectors in ONE launch (multi-component kernel; see op_has_mc_kernel)
int op_apply_cells(Operator &op, double *d_dst, const double *d_src, uint32_t cell_begin, uint32_t cell_end,
                   double *d_dot, cudaStream_t s, int ncomp = 1);
// vector-valued applies of this operator can use the multi-component kernel (G streamed once): collocated Laplace with
// stored geometric factors; B200FE_MULTI_COMPONENT=0 falls back to one launch per component
bool op_has_mc_kernel(const Operator &op);
// dst[c] = src[c] on owned constrained DoFs; optional dot += sum src[c]^2
int op_copy_constrained(Operator &op, double *d_dst, const double *d_src, double *d_dot, cudaStream_t s, int ncomp = 1);
// hanging-node rows: src[h] = sum w src[parents] (old values saved when `save`), and its transpose on dst
// (dst[parents] += w dst[h]; dst[h] = 0; src[h] restored when `restore`)
// (all `ncomp` components of component-blocked vectors in one launch)
int op_distribute(Operator &op, double *d_v, bool save, cudaStream_t s, int ncomp = 1);
int op_condense(Operator &op, double *d_dst, double *d_src_restore, cudaStream_t s, int ncomp = 1);
// full local vmult: zero, cells, constrained rows (+ halo exchange when attached)
// ncomp > 1: component-blocked vectors [component][n_local], the scalar operator on each
// dst_is_zero: the caller has already cleared dst (owned + ghost entries of every component), e.g. inside its own vector pass
int op_vmult(Operator &op, double *d_dst, const double *d_src, double *d_dot, bool ghost_on, bool compute_on,
             cudaStream_t s, int ncomp = 1, bool dst_is_zero = false);

// compute_diagonal of C^T A C (needs the constraint rows on the operator)
int op_diagonal_constrained(Operator &op, double *d_diag, cudaStream_t s);

// number of kernels this library has launched so far (all kinds)
extern unsigned long long g_launch_count;

// cg.cu: frees the CG workspace cached for this operator
void cg_release_work(Operator *op);

}  // namespace b200fe
