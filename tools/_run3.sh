for v in default ns3 ns4; do
  if [ "$v" = default ]; then unset B200NMPC_LIB; else export B200NMPC_LIB=$PWD/tum_control_b200/csrc/build/variants/libb200nmpc_$v.so; fi
  echo "== $v"; python tools/latency_small_batch.py 1 16 148 592 1184 2>&1 | cut -c1-200
done
bash tools/gpu_ab.sh "default ns3 ns4" "cfg3 cfg2" 20 ab_nslot
