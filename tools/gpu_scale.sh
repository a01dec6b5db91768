#!/bin/bash
# multi-GPU weak-scaling runs on one box (run via `gpurun --gpus 8`): headline workload at 1/2/4/8 GPUs, then the two BASELINE
# configurations that are defined multi-GPU (configs[2] SNMPC 1 -> 8, configs[4] R2NMPC 2048 per GPU x 8) at 1 and 8 GPUs
mkdir -p gpurun_out
OUT=gpurun_out/scaling_${1:-r02}.jsonl
run() {   # gpus workload
  if [ "$1" = 1 ]; then python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu --workload $2 >> $OUT 2>gpurun_out/scale.err || tail -3 gpurun_out/scale.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29500 + $1)) bench.py --gpus $1 --steps 20 --warmup 3 --no-cpu --workload $2 >> $OUT 2>gpurun_out/scale.err || tail -3 gpurun_out/scale.err; fi
}
if [ "$2" = quick ]; then for w in cfg2 cfg3 cfg5; do run 8 $w; done      # (the 1-GPU lines come from the single-GPU runs)
else
for n in 1 2 4 8; do run $n cfg2; done
for w in cfg3 cfg5; do for n in 1 8; do run $n $w; done; done
fi
python - <<PY
import json
for l in open("$OUT"):
    if not l.startswith("{"): continue
    d = json.loads(l); r = d["roofline"]
    print("%-6s gpus %d  value %10.0f  e2e %10.0f  ms/step %.3f  allgather_ms %s" % (d["config"]["name"], d["n_gpus"], d["value"], d["e2e"]["value"], d["ms_per_step"], r.get("allgather_ms")))
PY
