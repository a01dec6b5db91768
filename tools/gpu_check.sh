# GPU-box check used during development: the -m gpu suite, then two headline bench runs (run via gpurun from the repo root)
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for i in 1 2; do python bench.py --steps 20 --warmup 3 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value',round(d['value']),'kernel_ms',d['roofline']['kernel_ms'],'e2e',round(d['e2e']['value']),'warm',round(d['solver']['warm_solves_per_s']))"; done
