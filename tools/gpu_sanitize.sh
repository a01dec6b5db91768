# compute-sanitizer pass over every kernel of libb200nmpc on a small case (run via gpurun from the repo root): once as
# launched (the small-batch instantiation of the QP kernel), once with the throughput instantiation forced, once with the
# throughput instantiation in the long-horizon layout of the per-stage shared block
mkdir -p gpurun_out
for mode in latency throughput longh; do
  if [ $mode = throughput ]; then export B200NMPC_FORCE_THROUGHPUT=1; fi
  if [ $mode = longh ]; then export B200NMPC_FORCE_THROUGHPUT=1 B200NMPC_QP_LONGH=1; fi
  for tool in memcheck racecheck synccheck; do
    timeout 400 compute-sanitizer --tool $tool python tools/sanitize_case.py > gpurun_out/san_${mode}_$tool.log 2>&1
    echo "$mode $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/san_${mode}_$tool.log | tail -1)"
  done
done
tail -3 gpurun_out/san_latency_memcheck.log
