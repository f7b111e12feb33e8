#!/bin/bash
# ncu launch list (gpu__time_duration per launch) of eager net evaluations -> gpurun_out/$1
mkdir -p gpurun_out
ncu --clock-control none --metrics gpu__time_duration.sum --csv --log-file gpurun_out/$1 python tools/one_eval.py 2 8 > gpurun_out/one_eval.log 2>&1
tail -1 gpurun_out/one_eval.log
