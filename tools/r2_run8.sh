#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_net_gpu.py -m gpu -q --no-header -p no:cacheprovider > gpurun_out/ops_net.log 2>&1
echo "ops+net rc=$? $(tail -n 1 gpurun_out/ops_net.log)"; grep -E "^FAILED|^ERROR" gpurun_out/ops_net.log | head
python tools/graph_profile.py cfg2 4 0 2>/dev/null > gpurun_out/graph_cfg2_c.txt; head -3 gpurun_out/graph_cfg2_c.txt; grep narrow_conv gpurun_out/graph_cfg2_c.txt
python tools/graph_profile.py cfg2 4 1 2>/dev/null | head -2
