#!/bin/bash
# ncu launch list (gpu__time_duration per launch) of eager net evaluations of the README config
mkdir -p gpurun_out
ncu --clock-control none --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_final.csv python tools/one_eval.py 3 8 > gpurun_out/one_eval.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_final.csv > gpurun_out/launches_final_summary.txt 2>&1
rm -f gpurun_out/launches_final.csv
head -4 gpurun_out/launches_final_summary.txt | cut -c1-200
