python -m pytest tests/test_gpu_parity.py tests/test_gpu_snmpc.py tests/test_gpu_baseline_sizes.py tests/test_gpu_edge_cases.py -m gpu -x -q 2>&1 | tail -3
bash tools/gpu_ab.sh "prev default" "cfg2 cfg3 cfg4-80" 20 ab_fastpipe
