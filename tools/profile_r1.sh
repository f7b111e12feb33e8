#!/bin/bash
# ncu launch list of one net evaluation + full captures of the top kernels -> gpurun_out/
mkdir -p gpurun_out
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_r1.csv python tools/one_eval.py 3 8 > gpurun_out/one_eval.log 2>&1
$NCU --set full --import-source on -k regex:conv_gemm -s 3 -c 1 -f -o gpurun_out/conv_L7 python tools/one_kernel.py conv 2048 1024 1024 3 > /dev/null 2>&1
$NCU --set full --import-source on -k regex:conv_gemm -s 3 -c 1 -f -o gpurun_out/conv_L1 python tools/one_kernel.py conv 524288 32 32 3 > /dev/null 2>&1
$NCU --set full --import-source on -k regex:attention -s 3 -c 1 -f -o gpurun_out/attn_1024 python tools/one_kernel.py attn 8 8 1024 > /dev/null 2>&1
$NCU --set full --import-source on -k regex:ln_film -s 3 -c 1 -f -o gpurun_out/lnfilm_L1 python tools/one_kernel.py lnfilm 8 65536 32 > /dev/null 2>&1
ls -la gpurun_out/
