// dispatch.cu -- runtime (nm, nq, variant) -> instantiated launcher.
#include <cstdlib>

#include "kernels.h"

namespace b200fe {

int grid_multiplier()
{
    static int mult = [] {
        const char *e = std::getenv("B200FE_GRID_MULT");
        int v = e ? std::atoi(e) : 1;
        return v < 1 ? 1 : v;
    }();
    return mult;
}

namespace {
template <int P>
cudaError_t by_degree(int nq, bool coll, int qop, bool lvec, const double *hB, const double *hD,
                      const KArgs &a, cudaStream_t s, LaunchInfo *info, bool dry, const double *hW)
{
    constexpr int NM = P + 1;
    if (!lvec) {
        if (!coll && nq == NM + 1 && qop == QOP_MASS) return launch_t<NM, NM + 1, false, QOP_MASS, false>(hB, hD, hW, a, s, info, dry);
        if (!coll && nq == NM + 1 && qop == QOP_LAPLACE) return launch_t<NM, NM + 1, false, QOP_LAPLACE, false>(hB, hD, hW, a, s, info, dry);
        if (coll && nq == NM && qop == QOP_LAPLACE) return launch_t<NM, NM, true, QOP_LAPLACE, false>(hB, hD, hW, a, s, info, dry);
    } else {
        if (!coll && nq == NM + 1 && qop == QOP_LAPLACE) return launch_t<NM, NM + 1, false, QOP_LAPLACE, true>(hB, hD, hW, a, s, info, dry);
        if (!coll && nq == NM && qop == QOP_LAPLACE) return launch_t<NM, NM, false, QOP_LAPLACE, true>(hB, hD, hW, a, s, info, dry);
        if (coll && nq == NM && qop == QOP_LAPLACE) return launch_t<NM, NM, true, QOP_LAPLACE, true>(hB, hD, hW, a, s, info, dry);
        if (!coll && nq == NM + 1 && qop == QOP_MASS) return launch_t<NM, NM + 1, false, QOP_MASS, true>(hB, hD, hW, a, s, info, dry);
        if (!coll && nq == NM && qop == QOP_HELMHOLTZ) return launch_t<NM, NM, false, QOP_HELMHOLTZ, true>(hB, hD, hW, a, s, info, dry);
        constexpr int LA = QOP_LAPLACE | QOP_AFFINE;
        if (!coll && nq == NM + 1 && qop == LA) return launch_t<NM, NM + 1, false, LA, true>(hB, hD, hW, a, s, info, dry);
        if (!coll && nq == NM && qop == LA) return launch_t<NM, NM, false, LA, true>(hB, hD, hW, a, s, info, dry);
        if (coll && nq == NM && qop == LA) return launch_t<NM, NM, true, LA, true>(hB, hD, hW, a, s, info, dry);
        constexpr int LC = QOP_LAPLACE | QOP_AFFINE | QOP_CARTESIAN;
        if (coll && nq == NM && qop == LC) return launch_t<NM, NM, true, LC, true>(hB, hD, hW, a, s, info, dry);
        constexpr int LT = QOP_LAPLACE | QOP_TRILINEAR;
        if (!coll && nq == NM + 1 && qop == LT) return launch_t<NM, NM + 1, false, LT, true>(hB, hD, hW, a, s, info, dry);
        if (!coll && nq == NM && qop == LT) return launch_t<NM, NM, false, LT, true>(hB, hD, hW, a, s, info, dry);
        if (coll && nq == NM && qop == LT) return launch_t<NM, NM, true, LT, true>(hB, hD, hW, a, s, info, dry);
    }
    return cudaErrorInvalidValue;
}
}  // namespace

cudaError_t launch_sumfact(int nm, int nq, bool coll, int qop, bool lvec, const double *hB,
                           const double *hD, const KArgs &a, cudaStream_t s, LaunchInfo *info,
                           bool dry, const double *hW)
{
    switch (nm - 1) {
        case 1: return by_degree<1>(nq, coll, qop, lvec, hB, hD, a, s, info, dry, hW);
        case 2: return by_degree<2>(nq, coll, qop, lvec, hB, hD, a, s, info, dry, hW);
        case 3: return by_degree<3>(nq, coll, qop, lvec, hB, hD, a, s, info, dry, hW);
        case 4: return by_degree<4>(nq, coll, qop, lvec, hB, hD, a, s, info, dry, hW);
        case 5: return by_degree<5>(nq, coll, qop, lvec, hB, hD, a, s, info, dry, hW);
        case 6: return by_degree<6>(nq, coll, qop, lvec, hB, hD, a, s, info, dry, hW);
        case 7: return by_degree<7>(nq, coll, qop, lvec, hB, hD, a, s, info, dry, hW);
        case 8: return by_degree<8>(nq, coll, qop, lvec, hB, hD, a, s, info, dry, hW);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_cartesian(int nm, int qop, const double *hKM, const KArgs &a, cudaStream_t s, LaunchInfo *info, bool dry)
{
    switch (nm - 1) {
        case 1: return launch_cart_t<2>(qop, hKM, a, s, info, dry);
        case 2: return launch_cart_t<3>(qop, hKM, a, s, info, dry);
        case 3: return launch_cart_t<4>(qop, hKM, a, s, info, dry);
        case 4: return launch_cart_t<5>(qop, hKM, a, s, info, dry);
        case 5: return launch_cart_t<6>(qop, hKM, a, s, info, dry);
        case 6: return launch_cart_t<7>(qop, hKM, a, s, info, dry);
        case 7: return launch_cart_t<8>(qop, hKM, a, s, info, dry);
        case 8: return launch_cart_t<9>(qop, hKM, a, s, info, dry);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace b200fe
