mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_qp_solve_warp -s 2 -c 1 -o gpurun_out/qp_snmpc_r02b -f python tools/prof_one.py snmpc > gpurun_out/ncu_qp_snmpc_r02b.log 2>&1
tail -3 gpurun_out/ncu_qp_snmpc_r02b.log
