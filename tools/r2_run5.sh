#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_tests_files.sh tests/test_bwd_ops_gpu.py tests/test_train_gpu.py tests/test_ops_gpu.py tests/test_net_gpu.py > gpurun_out/tests_digest.txt 2>&1
grep -E "^==|FAILED|Error|timed out|VInpainter" gpurun_out/tests_digest.txt | head -40
python tools/time_train.py > gpurun_out/train_profile2.txt 2>&1; grep -v Warn gpurun_out/train_profile2.txt | head -64
python tools/graph_profile.py cfg2 4 0 2>/dev/null > gpurun_out/graph_cfg2_b.txt; head -30 gpurun_out/graph_cfg2_b.txt
