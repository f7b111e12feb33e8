#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q --no-header -p no:cacheprovider -k "pairs" -s > gpurun_out/pairs.log 2>&1
echo "pairs rc=$? $(tail -n 1 gpurun_out/pairs.log)"; grep -E "pair GEMM|FAILED|Error|timed out|differs" gpurun_out/pairs.log | cut -c1-200 | head -30
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_net_gpu.py -m gpu -q --no-header -p no:cacheprovider > gpurun_out/ops_net.log 2>&1
echo "ops+net rc=$? $(tail -n 1 gpurun_out/ops_net.log)"; grep -E "^FAILED|^ERROR" gpurun_out/ops_net.log | head
python tools/time_gemm_bn.py > gpurun_out/gemm_bn_pairs.txt 2>&1; cat gpurun_out/gemm_bn_pairs.txt
python - <<'PY' > gpurun_out/gemm_pairs_ab.txt 2>&1
import sys; sys.path.insert(0, '.')
from audio_diffusion_pytorch_b200 import _lib
from tools.time_gemm import run
L = _lib.lib()
for name, M, K, N, taps in [("L7 conv3", 2048, 1024, 1024, 3), ("L8 conv3", 1024, 1024, 1024, 3), ("L5 conv3", 8192, 512, 512, 3),
                            ("L6 conv3", 4096, 512, 512, 3), ("L7 qkv", 2048, 1024, 1536, 1), ("L5 qkv", 8192, 512, 1536, 1),
                            ("cfg5 L7", 4096, 1024, 1024, 3), ("cfg3 L7", 8192, 1024, 1024, 3)]:
    row = []
    for dis in (1, 0):
        L.adp_debug_set(0, dis)
        us, tf = run(M, K, N, taps, 128, res=(taps == 3), stats=False)
        row.append(f"{'single' if dis else 'pairs '}: {us:6.1f}us {tf:5.0f}TF")
    L.adp_debug_set(0, 0)
    print(f"{name:9s} bn128 | " + " | ".join(row), flush=True)
PY
cat gpurun_out/gemm_pairs_ab.txt
python bench.py --steps 3 --warmup 3 --no-train --no-cpu-baseline > gpurun_out/bench_pairs.json 2> gpurun_out/bench_pairs.err; python -c "
import json; d=json.load(open('gpurun_out/bench_pairs.json')); print(d['value'], d['ms_per_net_eval'], d['roofline']['frac'], d['roofline'].get('frac_in_graph'))"
