"""Per-kernel tables of one net evaluation with and without the GroupNorm-fused GEMM.
usage: python tools/compare_fusion.py [batch]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import audio_diffusion_pytorch_b200 as adp  # noqa: E402
from bench import README  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 8
torch.manual_seed(0)
model = adp.DiffusionModel(net_t=adp.UNetV0, **README).cuda()
net = model.net
x = torch.randn(batch, 2, 2 ** 18, device="cuda")
sig = torch.full((batch,), 0.5, device="cuda")
for fuse in (False, True):
    net.fuse_groupnorm = fuse
    net._plans.clear()
    net.use_cuda_graph = False
    for _ in range(2):
        net(x, sig)
    plan = next(iter(net._plans.values()))
    table = net.profile_plan(plan, iters=5)
    tot = sum(r["ms_total"] for r in table.values())
    print(f"==== fuse_groupnorm={fuse}: {tot:.3f} ms eager-summed, {sum(r['count'] for r in table.values())} launches")
    for r in sorted(table.values(), key=lambda r: -r["ms_total"]):
        if r["ms_total"] < 0.02:
            continue
        print(f"{r['name']:64s} x{r['count']:3d} avg {r['ms_avg'] * 1e3:8.1f} us total {r['ms_total'] * 1e3:8.1f} us "
              f"{r['flops'] / max(r['ms_avg'], 1e-9) / 1e9:7.1f} TF/s {r['bytes'] / max(r['ms_avg'], 1e-9) / 1e6:7.1f} GB/s")
