# compute-sanitizer pass over every kernel of libb200nmpc on a small case (run via gpurun from the repo root)
for tool in memcheck racecheck synccheck; do
  timeout 280 compute-sanitizer --tool $tool python tools/sanitize_case.py > gpurun_out/san_$tool.log 2>&1
  echo "$tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/san_$tool.log | tail -1)"
done
tail -3 gpurun_out/san_memcheck.log
