// kernels.h -- host-side view of the templated cell kernels (declarations only).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "sumfact.cuh"

namespace b200fe {

struct LaunchInfo {
    int elems_per_block, num_blocks, threads_per_block, smem_bytes, blocks_per_sm, regs_per_thread;
    int even_odd;  // 1: the even-odd kernel ran (symmetric 1-D matrices), 0: plain contractions
};

// Defined (explicitly instantiated) in inst.cu, one translation unit per degree.
// hB: nq*nm doubles B[q*nm+i] (may be null when COLL), hD: nq*nq doubles D[p*nq+n] (may be null for mass).
template <int NM, int NQ, bool COLL, int QOP, bool LVEC>
cudaError_t launch_t(const double *hB, const double *hD, const double *hW, const KArgs &a, cudaStream_t s,
                     LaunchInfo *info, bool dry_run);

// Runtime dispatch over the instantiated set.  Returns cudaErrorInvalidValue for a combination
// that is not built (caller maps it to B200FE_ERR_UNSUPPORTED).
// hW: nq quadrature weights, only read for qop & QOP_AFFINE.
cudaError_t launch_sumfact(int nm, int nq, bool coll, int qop, bool lvec, const double *hB,
                           const double *hD, const KArgs &a, cudaStream_t s, LaunchInfo *info,
                           bool dry_run, const double *hW = nullptr);

// Separable kernel for axis-aligned cells, interpolated operators (sumfact_cart.cuh).  hKM: the 1-D stiffness matrix K
// followed by the 1-D mass matrix M, nm*nm doubles each, row-major.  launch_cart_t is instantiated per degree in inst.cu.
// qop: QOP_LAPLACE, QOP_MASS or QOP_HELMHOLTZ (mass term = det J * M x M x M, KArgs::cellG[e][6]).
template <int NM>
cudaError_t launch_cart_t(int qop, const double *hKM, const KArgs &a, cudaStream_t s, LaunchInfo *info, bool dry_run);
cudaError_t launch_cartesian(int nm, int qop, const double *hKM, const KArgs &a, cudaStream_t s, LaunchInfo *info, bool dry_run);

// grid multiplier for the persistent launches (env B200FE_GRID_MULT, default 1 = one resident wave)
int grid_multiplier();

}  // namespace b200fe
