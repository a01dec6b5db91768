#!/bin/bash
# GPU-box profiling pass of round 2 (run via gpurun from the repo root):  tools/gpu_profile.sh <tag> [variants...]
#   launch list of the headline bench command, then one `ncu --set full` capture of k_qp_solve_warp and one of k_linearize
#   for the shipped library, and one k_qp_solve_warp capture per extra variant (tools/build_variants.sh) for same-version A/Bs.
TAG=${1:-r02}; shift
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/launches_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_qp_solve_warp -s 2 -c 1 -o gpurun_out/qp_$TAG -f python tools/prof_one.py nominal > gpurun_out/ncu_qp_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_linearize -s 2 -c 1 -o gpurun_out/lin_$TAG -f python tools/prof_one.py nominal > gpurun_out/ncu_lin_$TAG.log 2>&1
for v in "$@"; do
  B200NMPC_LIB=$PWD/tum_control_b200/csrc/build/variants/libb200nmpc_$v.so ncu --set full --clock-control none --import-source on -k regex:k_qp_solve_warp -s 2 -c 1 -o gpurun_out/qp_${TAG}_$v -f python tools/prof_one.py nominal > gpurun_out/ncu_qp_${TAG}_$v.log 2>&1
done
ls -la gpurun_out | tail -12
