#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_tests_files.sh > gpurun_out/tests_digest.txt 2>&1
tail -n 60 gpurun_out/tests_digest.txt | grep -E "^==|FAILED|Error" 
python tools/time_gemm_bn.py > gpurun_out/gemm_bn.txt 2>&1; cat gpurun_out/gemm_bn.txt
python bench.py --steps 3 --warmup 3 --profile-ops > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; tail -c 3000 gpurun_out/bench_cfg2.json; grep "^#" gpurun_out/bench_cfg2.err | head -40
python bench.py --config cfg3 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err; tail -c 1500 gpurun_out/bench_cfg3.json; tail -n 5 gpurun_out/bench_cfg3.err
python bench.py --config cfg5 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg5.json 2> gpurun_out/bench_cfg5.err; tail -c 1500 gpurun_out/bench_cfg5.json; tail -n 5 gpurun_out/bench_cfg5.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 4
