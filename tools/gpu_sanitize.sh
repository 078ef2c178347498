set -x
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python tools/sanitize_run.py > gpurun_out/san_$tool.log 2>&1; echo "$tool rc=$?" >> gpurun_out/san_summary.txt
  tail -4 gpurun_out/san_$tool.log >> gpurun_out/san_summary.txt
done
cat gpurun_out/san_summary.txt
