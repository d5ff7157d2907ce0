// sumfact.cuh -- types shared by the sum-factorised cell kernels for sm_100a (FP64); the kernel itself is sumfact2.cuh.
// (The first-generation kernel that used to live here -- latency-bound, 54 % of the HBM roofline, profiles/r01a_* -- was
// removed in round 2 once the second generation had replaced it everywhere.)
//
// One templated kernel covers the reference's whole kernel zoo for this path:
//   E-vector BK1 / BK3 / BK5   CEED_BK/include/kernels/BK{1,3,5}/templated_cuda_kernels.cuh
//   L-vector BP apply          CEED_bp/include/bk3_kokkos_kernel.h:27-389 (nq = p+2)
//                              bakeoff_problems_dealii/include/bk3_kokkos_kernel.h:32-434 (nq = p+1)
//   Helmholtz quad op          bp5_kokkos/benchmark.cc:62-137
//
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "eo_contract.h"

namespace b200fe {

constexpr uint32_t kInvalidIndex = 0xFFFFFFFFu;  // numbers::invalid_unsigned_int

enum : int { QOP_LAPLACE = 1, QOP_MASS = 2, QOP_HELMHOLTZ = 3,
             // geometry evaluated on the fly for affine cells: G(q) = cellG * w_p w_q w_r with six
             // per-cell constants instead of the streamed 6 nq^3 factors; sumfact2 kernel, L-vector operators only
             QOP_AFFINE = 4,
             // geometry evaluated on the fly for TRILINEAR cells (general hexahedra given by their 8 vertices): the Jacobian
             // at a quadrature point is rebuilt from 24 vertex coordinates per cell and G = JxW K K^T formed in registers
             QOP_TRILINEAR = 8,
             // with QOP_AFFINE, collocated operators: every cell is an axis-aligned box (deal.II's "cartesian" cell type; the
             // reference's own meshes) -- G is diagonal and separable, so D^T G D u collapses to
             //   c_rr (S x W x W) u + c_ss (W x S x W) u + c_tt (W x W x S) u,   S = D^T W D  (1-D stiffness matrix, passed as B):
             // three 1-D contractions per point instead of six, 8 instead of 16 shared-memory accesses per point
             QOP_CARTESIAN = 16 };

// 1-D matrices in the BK layout: B[q*NM+i] (CEED_BK BK1 serial_kernels.hpp:39),
// D[p*NQ+n] = derivative of collocation function n at point p (BK3 serial_kernels.hpp:98).
// EO = true: the even / odd halves of B and D instead (eo_contract.h) -- the kernels instantiated with it run every 1-D
// contraction through the even-odd split; chosen at launch time when the matrices have the symmetry of a real basis.
template <int NM, int NQ, bool EO = false>
struct Mats {
    static constexpr bool kEvenOdd = false;
    double B[NQ * NM];
    double D[NQ * NQ];
    double W[NQ];  // 1-D quadrature weights (on-the-fly geometry only)
    double X[NQ];  // 1-D quadrature points on [0,1] (trilinear on-the-fly geometry only)
};
template <int NM, int NQ>
struct Mats<NM, NQ, true> {
    static constexpr bool kEvenOdd = true;
    eo::EoMats<NM, NQ> E;
    double W[NQ];
    double X[NQ];
};

struct KArgs {
    uint32_t n_elems;       // elements handled by this launch
    const double *G;        // [e][6][NQ^3], components rr,rs,rt,ss,st,tt, r <-> slowest index
    const double *JxW;      // [e][NQ^3]
    const double *in;       // E-vector [e][NM^3]  | L-vector src (owned + ghosts)
    double *out;            // E-vector [e][NM^3]  | L-vector dst
    const uint32_t *idx;    // L-vector only: [e][NM^3], kInvalidIndex = constrained
    double *dot;            // L-vector only, optional: += sum_e u_e . (A_e u_e)
    const double *cellG;    // affine geometry: [e][8] = det J * K K^T (rr,rs,rt,ss,st,tt), det J, pad;
                            // trilinear geometry: [e][3][2][2][2] vertex coordinates (x fastest), 24 doubles per cell
    const int *skip;        // optional: when *skip != 0 the launch is a no-op (CG iterations replayed after convergence)
    // vector-valued operators (BP2/4/6 = the scalar operator on every component), multi-component kernels only: the
    // geometric factors of a batch are fetched once and used for all components; in/out of component c at + c*comp_stride
    int ncomp = 1;
    size_t comp_stride = 0;
    // L-vector only: the DoFs at the interior positions of a cell (all three local indices in 1..nm-2) belong to that cell
    // alone (FE_Q; verified on the index table at operator creation).  When set, they are written with plain stores -- no
    // zero-fill before the launch and no read-modify-write in L2 for (p-1)^3 / p^3 of the result vector.
    int excl_interior = 0;
};

constexpr __host__ __device__ int odd(int n) { return n | 1; }
constexpr __host__ __device__ int cmax(int a, int b) { return a > b ? a : b; }

}  // namespace b200fe
