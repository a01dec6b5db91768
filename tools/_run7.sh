python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_r02f.log
bash tools/gpu_profile.sh r02f > gpurun_out/profile_r02f.log 2>&1; tail -4 gpurun_out/profile_r02f.log
bash tools/gpu_sanitize.sh 2>&1 | tail -12
for w in cfg2 cfg3 cfg4-20 cfg4-40 cfg4-80 cfg5; do python bench.py --steps 30 --warmup 5 --workload $w >> gpurun_out/final_configs_r02f.jsonl 2>gpurun_out/final.err || tail -3 gpurun_out/final.err; done
python tools/latency_small_batch.py > gpurun_out/batch_sweep_r02f.jsonl 2>&1
python tools/dropin_loop_latency.py > gpurun_out/dropin_r02f.jsonl 2>&1
tail -2 gpurun_out/batch_sweep_r02f.jsonl | cut -c1-150
