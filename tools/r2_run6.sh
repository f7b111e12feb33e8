#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_tests_files.sh tests/test_bwd_ops_gpu.py tests/test_train_gpu.py > gpurun_out/tests_digest.txt 2>&1
grep -E "^==|FAILED|Error|timed out" gpurun_out/tests_digest.txt | head -40
python tools/time_train.py > gpurun_out/train_profile3.txt 2>&1; grep -v Warn gpurun_out/train_profile3.txt | head -12; grep -A14 "backward graph:" gpurun_out/train_profile3.txt
python tools/time_gemm_k1.py 2>&1 | tee gpurun_out/gemm_k1.txt
