#!/bin/bash
# A/B of kernel builds on the GPU box: tools/gpu_ab.sh "<variants>" "<workloads>" [steps]  ->  gpurun_out/ab_<tag>.jsonl + table
# variants: names built by tools/build_variants.sh (or "default" = the shipped library)
V=${1:-"default"}; W=${2:-"cfg2"}; K=${3:-20}; TAG=${4:-ab}
mkdir -p gpurun_out
for rep in 1 2; do
for v in $V; do
  for w in $W; do
    if [ "$v" = default ]; then unset B200NMPC_LIB; else export B200NMPC_LIB=$PWD/tum_control_b200/csrc/build/variants/libb200nmpc_$v.so; fi
    python bench.py --steps $K --warmup 3 --no-cpu --workload $w 2>gpurun_out/ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
d['variant']='$v'
print(json.dumps(d), file=open('gpurun_out/${TAG}.jsonl','a'))
f=r.get('first_solve_of_a_handle',{}).get('solves_per_s',0); wm=r.get('warm_rti',{}).get('solves_per_s',0)
print('%-8s %-8s value %9d  kernel_ms %7.3f  lin_ms %6.3f  e2e %9d  first %9d  warm %9d' % ('$v','$w',d['value'],r['kernel_ms'],r['linearize_ms'],d['e2e']['value'],f,wm))
" || tail -3 gpurun_out/ab.err
  done
done
done
