#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_tests_files.sh tests/test_bwd_ops_gpu.py tests/test_train_gpu.py > gpurun_out/tests_digest.txt 2>&1
grep -E "^==|FAILED|Error|timed out" gpurun_out/tests_digest.txt | head -40
python tools/time_train.py > gpurun_out/train_profile4.txt 2>&1; grep -v Warn gpurun_out/train_profile4.txt | head -8; grep -A12 "backward graph:" gpurun_out/train_profile4.txt
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_cfg2_b.json 2> gpurun_out/bench_cfg2_b.err; python -c "
import json; d=json.load(open('gpurun_out/bench_cfg2_b.json')); r=d['roofline']; print('cfg2', d['value'], d['e2e']['value'], d['ms_per_net_eval'], r['kernel'], r['frac'], r['kernel_us'], r.get('achieved_eager_events'), r['step_frac'], d['train_step'])"; tail -3 gpurun_out/bench_cfg2_b.err
