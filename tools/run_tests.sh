set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python tools/sanitize.py 160 2>&1 | tail -3
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize.py 20 > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|SANITIZE_WORKLOAD_OK|Error|error" gpurun_out/sanitizer_$tool.log | head -8
done
