python -m pytest tests/test_gpu_api.py -m gpu -x -q 2>&1 | tail -4
bash tools/gpu_ab_env.sh "none" "cfg2 cfg5" 20 abenv3
