#!/bin/bash
# round-1 final profiles: launch list of one net evaluation + full captures of the top GEMMs
mkdir -p gpurun_out
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_r1c.csv python tools/one_eval.py 3 8 > gpurun_out/one_eval.log 2>&1
$NCU --set full --import-source on -k regex:conv_gemm -s 3 -c 1 -f -o gpurun_out/conv_L7_v4 python tools/one_kernel.py conv 2048 1024 1024 3 0 1 > /dev/null 2>&1
$NCU --set full --import-source on -k regex:conv_gemm -s 3 -c 1 -f -o gpurun_out/conv_L3_v4 python tools/one_kernel.py conv 32768 128 128 3 0 1 > /dev/null 2>&1
ls -la gpurun_out/ | grep -E "ncu-rep|launches_r1c"
