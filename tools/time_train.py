import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import audio_diffusion_pytorch_b200 as adp
from audio_diffusion_pytorch_b200 import training, ops
from bench import UPSAMPLER, LENGTH
dev = torch.device("cuda")
model = adp.DiffusionUpsampler(net_t=adp.UNetV0, **UPSAMPLER).to(dev)
opt = torch.optim.AdamW(model.parameters(), lr=1e-4, fused=True)
audio = torch.randn(4, 2, LENGTH, device=dev)
def step():
    opt.zero_grad(set_to_none=True); loss = model(audio); loss.backward(); opt.step()
for _ in range(4): step()
torch.cuda.synchronize()
net = model.net
def timeit(name, fn, n=5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); print(f"{name:30s} {(time.perf_counter()-t0)/n*1e3:8.2f} ms", flush=True)
plan = net._plans[("train", 4, LENGTH)]
timeit("full step", step)
def repack():
    net._packed_version = -1; net.packed()
    with torch.no_grad():
        for r in plan.refreshers: r()
timeit("repack fwd+dgrad weights", repack)
timeit("forward graph", lambda: plan.graph_f.replay())
timeit("backward graph", lambda: plan.graph_b.replay())
timeit("finals", lambda: [f() for f in plan.finals])
timeit("optimizer step", lambda: opt.step())
timeit("reupsample", lambda: model.reupsample(audio))
with ops.trace(timing=True) as tr:
    [f() for f in plan.fwd]
tb = tr.table(); print("fwd kernels", sum(r["count"] for r in tb.values()), "sum ms", sum(r["ms_total"] for r in tb.values()))
with ops.trace(timing=True) as tr:
    plan.flat.zero_(); plan.backward0(); plan.cond_backward()
tb = tr.table(); print("bwd kernels", sum(r["count"] for r in tb.values()), "sum ms", sum(r["ms_total"] for r in tb.values()))
for r in sorted(tb.values(), key=lambda r: -r["ms_total"])[:14]:
    print(f"  {r['name']:50s} x{r['count']:3d} total {r['ms_total']:7.3f} ms")
