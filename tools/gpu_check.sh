# GPU-box check used during development: the -m gpu suite, then headline bench runs (run via gpurun from the repo root)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
for i in 1 2; do python bench.py --steps 20 --warmup 3 --no-cpu | tee -a gpurun_out/bench_check.jsonl | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('value',round(d['value']),'kernel_ms',round(r['kernel_ms'],3),'lin_ms',round(r['linearize_ms'],3),'e2e',round(d['e2e']['value']),'warm',round(r['warm_rti']['solves_per_s']),'it',r['solver']['qp_iter_mean'],r['warm_rti']['qp_iter_mean'])"; done
