#!/bin/bash
mkdir -p gpurun_out
NCU="ncu --clock-control none"
$NCU --set full --import-source on -k regex:conv_gemm2 -s 3 -c 1 -f -o gpurun_out/conv2_L7 python tools/one_kernel.py conv 2048 1024 1024 3 > /dev/null 2>&1
$NCU --set full --import-source on -k regex:conv_gemm2 -s 3 -c 1 -f -o gpurun_out/conv2_L1 python tools/one_kernel.py conv 524288 32 32 3 > /dev/null 2>&1
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_r1b.csv python tools/one_eval.py 2 8 > gpurun_out/one_eval.log 2>&1
ls -la gpurun_out/*.ncu-rep
