"""In-graph kernel durations of the README sampling step via torch.profiler (CUPTI sees the
kernels of a CUDA-graph replay).  usage: python tools/graph_profile.py [steps]"""
import os
import sys
from collections import defaultdict

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import audio_diffusion_pytorch_b200 as adp  # noqa: E402
from bench import README  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
pdl = int(sys.argv[2]) if len(sys.argv) > 2 else 1
from audio_diffusion_pytorch_b200 import _lib  # noqa: E402
_lib.lib().adp_debug_set(6, pdl)        # programmatic dependent launch on/off (before capture)
torch.manual_seed(0)
model = adp.DiffusionModel(net_t=adp.UNetV0, **README).cuda()
x = torch.randn(8, 2, 2 ** 18, device="cuda")
for _ in range(2):
    model.sample(x, num_steps=3)          # eager + capture
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
model.sample(x, num_steps=20)
e1.record()
torch.cuda.synchronize()
print(f"pdl={pdl}: un-profiled {e0.elapsed_time(e1) * 1e3 / 20:.1f} us per step")
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    model.sample(x, num_steps=steps)
    torch.cuda.synchronize()
agg = defaultdict(lambda: [0, 0.0])
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
t0 = min(e.time_range.start for e in evs)
t1 = max(e.time_range.end for e in evs)
for e in evs:
    name = e.name.split("(")[0].replace("void ", "").replace("adp::", "")
    a = agg[name]
    a[0] += 1
    a[1] += e.time_range.end - e.time_range.start
tot = sum(a[1] for a in agg.values())
print(f"{steps} steps: span {(t1 - t0) / steps:.1f} us/step, kernel-busy {tot / steps:.1f} us/step, {sum(a[0] for a in agg.values()) / steps:.1f} kernels/step")
for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{name[:90]:90s} x{n / steps:6.1f} {us / steps:9.1f} us/step  avg {us / n:7.2f} us")
