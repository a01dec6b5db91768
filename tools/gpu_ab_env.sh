#!/bin/bash
# A/B of run-time knobs on the GPU box: tools/gpu_ab_env.sh "<env settings>" "<workloads>" [steps] [tag]
# every setting is a comma-separated list of VAR=value (or "none"), e.g. "none B200NMPC_QP_WPB=1 BENCH_E2E_DEPTH=1,BENCH_E2E_SPLIT=4"
V=${1:-"none"}; W=${2:-"cfg2"}; K=${3:-20}; TAG=${4:-abenv}
mkdir -p gpurun_out
for rep in 1 2; do
for v in $V; do
  for w in $W; do
    ( if [ "$v" != none ]; then export ${v//,/ }; fi
    python bench.py --steps $K --warmup 3 --no-cpu --workload $w 2>gpurun_out/abenv.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
d['variant']='$v'
print(json.dumps(d), file=open('gpurun_out/${TAG}.jsonl','a'))
f=r.get('first_solve_of_a_handle',{}).get('solves_per_s',0); wm=r.get('warm_rti',{}).get('solves_per_s',0)
print('%-40s %-8s value %9d  kernel_ms %7.3f  lin_ms %6.3f  e2e %9d  sync %9d  first %9d  warm %9d' % ('$v','$w',d['value'],r['kernel_ms'],r['linearize_ms'],d['e2e']['value'],d['e2e'].get('synchronous',{}).get('value',0),f,wm))
" || tail -3 gpurun_out/abenv.err )
  done
done
done
