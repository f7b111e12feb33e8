#!/bin/bash
# Runs every GPU test file in its own process (a trapping kernel poisons its CUDA context), each
# under its own timeout; full logs in gpurun_out/, a digest on stdout.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu_info.txt 2>&1
status=0
files=${@:-tests/test_bwd_ops_gpu.py tests/test_train_gpu.py tests/test_net_gpu.py tests/test_ops_gpu.py tests/test_frontend_gpu.py tests/test_ddp_gpu.py}
for f in $files; do
  name=$(basename "$f" .py)
  timeout 1200 python -m pytest "$f" -m gpu -q --no-header -p no:cacheprovider -s \
      > "gpurun_out/${name}.log" 2>&1
  rc=$?
  echo "== $f -> rc=$rc: $(tail -n 1 gpurun_out/${name}.log)"
  grep -E "^FAILED|^ERROR|Error:|error:|timed out|rel-L2|worst per-parameter|loss .* vs oracle" "gpurun_out/${name}.log" | cut -c1-220 | head -80
  [ $rc -ne 0 ] && status=1
done
exit $status
