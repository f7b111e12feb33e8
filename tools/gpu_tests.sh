#!/bin/bash
# Runs the GPU test-suite in isolated processes (a trapping kernel poisons its CUDA context),
# each under its own timeout; logs to gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu_info.txt 2>&1
status=0
for sel in "$@"; do
  name=$(echo "$sel" | tr -c 'A-Za-z0-9_\n' '_')
  timeout 600 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -k "$sel" -s \
      > "gpurun_out/test_${name}.log" 2>&1
  rc=$?
  echo "== $sel -> rc=$rc: $(tail -n 1 gpurun_out/test_${name}.log)"
  grep -E "max abs err|FAILED|Error|error|timed out" "gpurun_out/test_${name}.log" | head -60
  [ $rc -ne 0 ] && status=1
done
exit $status
