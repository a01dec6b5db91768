#!/bin/bash
# A/B builds of the QP kernel's stage pipeline (only k_qp_warp.cu differs): tum_control_b200/csrc/build/variants/libb200nmpc_<name>.so
#   cpa1 = default (cp.async, one record per commit group)     cpa2 = cp.async, two records per commit group
#   tma1 = TMA bulk copy + mbarrier, one record per phase      tma2 = TMA bulk copy, two records per mbarrier phase
# select at run time with B200NMPC_LIB=<path> (tum_control_b200/_lib.py)
set -e
cd "$(dirname "$0")/../tum_control_b200/csrc"
make -s
mkdir -p build/variants
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
ARCH="-gencode arch=compute_100a,code=sm_100a"
FLAGS="$ARCH -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v --fmad=true"
OTHERS="build/b200nmpc.o build/k_linearize.o build/k_snmpc.o build/k_sim.o build/k_misc.o"
build() {   # name, extra flags
    $NVCC $FLAGS $2 -c k_qp_warp.cu -o build/variants/k_qp_warp_$1.o 2> build/variants/k_qp_warp_$1.ptxas.log
    $NVCC $ARCH -shared -o build/variants/libb200nmpc_$1.so build/variants/k_qp_warp_$1.o $OTHERS -lcudart
    echo "$1: $(grep -A1 'k_qp_solve_warpILb0ELb0' build/variants/k_qp_warp_$1.ptxas.log | grep -o 'Used [0-9]* registers' | head -1)"
}
# k_linearize compiled for 3 / 4 resident blocks per SM (168 / 128 registers) instead of 2 (255)
build_lin() {   # name, extra flags
    $NVCC $FLAGS $2 -c k_linearize.cu -o build/variants/k_linearize_$1.o 2> build/variants/k_linearize_$1.ptxas.log
    $NVCC $ARCH -shared -o build/variants/libb200nmpc_$1.so build/k_qp_warp.o build/b200nmpc.o build/variants/k_linearize_$1.o build/k_snmpc.o build/k_sim.o build/k_misc.o -lcudart
    echo "$1: $(grep -A2 'Function properties for _ZN4b20011k_linearize' build/variants/k_linearize_$1.ptxas.log | grep -o 'Used [0-9]* registers' | head -1)"
}
if [ "$1" = lin ]; then build_lin lin2d "-DLIN_DIRECT_STORE"; build_lin lin3d "-DLIN_MINB=3 -DLIN_DIRECT_STORE"; build_lin lin4d "-DLIN_MINB=4 -DLIN_DIRECT_STORE"; exit 0; fi
if [ "$1" = occ ]; then build w16 "-DQP_MINB=4"; build w20 "-DQP_MINB=5"; exit 0; fi
if [ "$1" = dmma ]; then build cpa1 ""; build vecpipe "-DQP_DMMA=0"; build tensor1 "-DQP_DMMA=1"; exit 0; fi
build cpa1 ""
build vecpipe "-DQP_DMMA=0"
build cpa2 "-DPIPE_GROUP=2"
build tma1 "-DPIPE_TMA"
build tma2 "-DPIPE_TMA -DPIPE_GROUP=2"
