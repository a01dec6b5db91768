# GPU-box profiling pass: other BASELINE configurations, headline bench, ncu launch list, one ncu --set full capture of k_qp_solve_warp
python bench_configs.py > gpurun_out/configs_v12.jsonl 2> gpurun_out/configs_v12.err
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_v12.json 2> gpurun_out/bench_v12.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v12.csv python bench.py --steps 2 --warmup 1 > gpurun_out/b12.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_qp_solve_warp -s 2 -c 1 -o gpurun_out/qp_v12 -f python tools/prof_one.py nominal > gpurun_out/ncu12.log 2>&1
tail -c 600 gpurun_out/configs_v12.err; ls -la gpurun_out | tail -6
