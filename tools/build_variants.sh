#!/bin/bash
# Builds the tuning variants of libb200fe.so into benchmarks_b200/variants/ (git-ignored, shipped by gpurun).  No GPU needed.
#   eo      even-odd decomposition of the 1-D contractions (symmetric, i.e. real, 1-D matrices only)
#   nopad3  unpadded quadrature planes for nq <= 3 (occupancy of the p = 1, 2 kernels)
#   eo_nopad3  both
#   eo_r112, eo_r136  even-odd + a lower register floor for nq >= 7 (more resident CTAs; the even-odd kernels need 98-188 registers)
set -e
cd "$(dirname "$0")/../benchmarks_b200/csrc"
mkdir -p ../variants
make -j8 BUILD=build_eo LIB=../variants/libb200fe_eo.so EXTRA_NVFLAGS="-DB200FE_EVEN_ODD"
make -j8 BUILD=build_nopad3 LIB=../variants/libb200fe_nopad3.so EXTRA_NVFLAGS="-DB200FE_V2_NOPAD_MAXNQ=3"
make -j8 BUILD=build_eo_nopad3 LIB=../variants/libb200fe_eo_nopad3.so EXTRA_NVFLAGS="-DB200FE_EVEN_ODD -DB200FE_V2_NOPAD_MAXNQ=3"
make -j8 BUILD=build_eo_r112 LIB=../variants/libb200fe_eo_r112.so EXTRA_NVFLAGS="-DB200FE_EVEN_ODD -DB200FE_V2_RMIN_HI=112"
make -j8 BUILD=build_eo_r136 LIB=../variants/libb200fe_eo_r136.so EXTRA_NVFLAGS="-DB200FE_EVEN_ODD -DB200FE_V2_RMIN_HI=136"
ls -la ../variants
