// inst.cu -- explicit instantiation of every kernel variant of one polynomial degree.
// Compiled once per degree with -DINST_P=<p> (see Makefile) so the degrees build in parallel.
#include "launch_impl.cuh"

#ifndef INST_P
#error "compile with -DINST_P=<degree>"
#endif

namespace b200fe {
constexpr int P = INST_P, NM = P + 1;
#define INST(nm, nq, coll, qop, lvec)                                                            \
    template cudaError_t launch_t<nm, nq, coll, qop, lvec>(const double *, const double *,       \
                                                            const double *, const KArgs &,       \
                                                            cudaStream_t, LaunchInfo *, bool);
// E-vector bake-off kernels
INST(NM, NM + 1, false, QOP_MASS, false)     // BK1
INST(NM, NM + 1, false, QOP_LAPLACE, false)  // BK3
INST(NM, NM, true, QOP_LAPLACE, false)       // BK5
// L-vector operators
INST(NM, NM + 1, false, QOP_LAPLACE, true)   // BP3  (QGauss(p+2))
INST(NM, NM, false, QOP_LAPLACE, true)       // "bp35" (QGauss(p+1))
INST(NM, NM, true, QOP_LAPLACE, true)        // BP5  (GLL collocated)
INST(NM, NM + 1, false, QOP_MASS, true)      // BP1  (QGauss(p+2))
INST(NM, NM, false, QOP_HELMHOLTZ, true)     // bp5_kokkos Helmholtz (QGauss(p+1))
// L-vector Laplace operators with geometry evaluated on the fly (affine cells)
INST(NM, NM + 1, false, QOP_LAPLACE | QOP_AFFINE, true)
INST(NM, NM, false, QOP_LAPLACE | QOP_AFFINE, true)
INST(NM, NM, true, QOP_LAPLACE | QOP_AFFINE, true)
// ... collocated, axis-aligned cells (deal.II's "cartesian" cell type): separable operator, three 1-D contractions per point
INST(NM, NM, true, QOP_LAPLACE | QOP_AFFINE | QOP_CARTESIAN, true)
// ... and for trilinear cells (8 vertices per cell, G rebuilt at every quadrature point)
INST(NM, NM + 1, false, QOP_LAPLACE | QOP_TRILINEAR, true)
INST(NM, NM, false, QOP_LAPLACE | QOP_TRILINEAR, true)
INST(NM, NM, true, QOP_LAPLACE | QOP_TRILINEAR, true)
// ... interpolated operators on axis-aligned cells: separable kernel on the nodal values (sumfact_cart.cuh)
template cudaError_t launch_cart_t<NM>(int, const double *, const KArgs &, cudaStream_t, LaunchInfo *, bool);
}  // namespace b200fe
