python -m pytest tests/test_gpu_snmpc.py tests/test_gpu_baseline_sizes.py -m gpu -x -q 2>&1 | tail -4
bash tools/gpu_ab.sh "default" "cfg3" 20 ab_headfac
