"""Runs a few eager (non-graph) net evaluations of the BASELINE config for `ncu` launch lists.
usage: python tools/one_eval.py [evals] [batch]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import audio_diffusion_pytorch_b200 as adp  # noqa: E402
from bench import README  # noqa: E402

evals = int(sys.argv[1]) if len(sys.argv) > 1 else 3
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 8
torch.manual_seed(0)
model = adp.DiffusionModel(net_t=adp.UNetV0, **README).cuda()
model.net.use_cuda_graph = False
x = torch.randn(batch, 2, 2 ** 18, device="cuda")
sig = torch.full((batch,), 0.5, device="cuda")
with torch.no_grad():          # the inference plan (with autograd recording, net() is the training forward)
    for _ in range(evals):
        v = model.net(x, sig)
torch.cuda.synchronize()
print("done", float(v.abs().mean()))
