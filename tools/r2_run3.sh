#!/bin/bash
mkdir -p gpurun_out
python tools/time_train.py > gpurun_out/train_profile.txt 2>&1; cat gpurun_out/train_profile.txt | grep -v Warn | head -80
bash tools/ncu_capture.sh conv_gemm r2_ncu_conv_L7 conv 2048 1024 1024 3 0 1
bash tools/ncu_capture.sh conv_gemm r2_ncu_conv_L5 conv 8192 512 512 3 0 1
bash tools/ncu_capture.sh attention r2_ncu_attn attn 8 8 1024
bash tools/ncu_capture.sh gn_silu r2_ncu_gn_silu gnsilu 8 256 1024
bash tools/ncu_capture.sh ln_film r2_ncu_ln_film lnfilm 8 256 1024
bash tools/ncu_capture.sh mid_conv r2_ncu_mid_conv64 mid 8 16384 64
bash tools/ncu_capture.sh mid_conv r2_ncu_mid_conv32 mid 8 65536 32
grep -E "tensor|time_duration|dram__bytes|issue_active" gpurun_out/r2_ncu_conv_L7.txt gpurun_out/r2_ncu_attn.txt | head -40
