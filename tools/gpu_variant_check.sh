#!/bin/bash
# parity of an A/B kernel build (tools/build_variants.sh) on the GPU box: the oracle-parity tests with B200NMPC_LIB pointing at it
for v in "$@"; do
  echo "== $v"
  B200NMPC_LIB=$PWD/tum_control_b200/csrc/build/variants/libb200nmpc_$v.so python -m pytest tests/test_gpu_parity.py tests/test_gpu_snmpc.py tests/test_gpu_edge_cases.py -m gpu -x -q 2>&1 | tail -3
done
