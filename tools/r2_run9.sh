#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_net_gpu.py -m gpu -q --no-header -p no:cacheprovider -s > gpurun_out/net.log 2>&1
echo "net rc=$? $(tail -n 1 gpurun_out/net.log)"; grep -E "^FAILED|^ERROR|rel-L2" gpurun_out/net.log | cut -c1-160 | head -30
python bench.py --steps 3 --warmup 3 --no-train --no-cpu-baseline > gpurun_out/bench_cfg2_c.json 2> gpurun_out/bench_cfg2_c.err; python -c "
import json; d=json.load(open('gpurun_out/bench_cfg2_c.json')); r=d['roofline']; print('cfg2', d['value'], d['e2e']['value'], d['ms_per_net_eval'], r['kernel'], r['frac'], r['kernel_us'], r.get('kernel_busy_us_per_net_eval'), d['gpu_launches'])"; tail -3 gpurun_out/bench_cfg2_c.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 3
