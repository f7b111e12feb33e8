#!/bin/bash
# usage: tools/r2_multi.sh N   (inside a gpurun --gpus N box)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
if [ "$N" = "2" ]; then
  timeout 600 python -m pytest tests/test_ddp_gpu.py -m gpu -q --no-header -p no:cacheprovider -s > gpurun_out/ddp.log 2>&1
  echo "ddp rc=$? $(tail -n 1 gpurun_out/ddp.log)"; grep -E "DP gradient|FAILED|Error" gpurun_out/ddp.log | cut -c1-300 | head
fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
python -c "
import json; d=json.load(open('gpurun_out/bench_n$N.json')); print('N=$N value', d['value'], 'e2e', d['e2e']['value'], 'train', d['train_step'])"
tail -n 4 gpurun_out/bench_n$N.err | cut -c1-300
