"""Experiment: the B = 8 sampling step as two independent B = 4 halves on two CUDA streams (every
kernel of the deep levels is latency-bound with idle SMs; two concurrent chains overlap each
other's launch / fill / drain skeletons).  Two model instances (separate weights: pessimistic for
L2) stand in for a two-plan implementation.  usage: python tools/exp_split_streams.py"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import audio_diffusion_pytorch_b200 as adp
from bench import README
dev = "cuda"
torch.manual_seed(0)
full = adp.DiffusionModel(net_t=adp.UNetV0, **README).to(dev)
halves = [adp.DiffusionModel(net_t=adp.UNetV0, **README).to(dev) for _ in range(2)]
for h in halves:
    h.load_state_dict(full.state_dict())
x = torch.randn(8, 2, 2 ** 18, device=dev)
streams = [torch.cuda.Stream(), torch.cuda.Stream()]
def run_full():
    return full.sample(x, num_steps=50)
def run_split():
    cur = torch.cuda.current_stream()
    outs = []
    for i, (m, s) in enumerate(zip(halves, streams)):
        s.wait_stream(cur)
        with torch.cuda.stream(s):
            outs.append(m.sample(x[4 * i:4 * i + 4], num_steps=50))
    for s in streams:
        cur.wait_stream(s)
    return torch.cat(outs)
def timeit(fn, n=3):
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for _ in range(4):
    a = run_full(); b = run_split()
print("max |full - split|", float((a - b).abs().max()))
for _ in range(2):
    print(f"full B=8: {timeit(run_full):7.1f} ms   two B=4 halves on two streams: {timeit(run_split):7.1f} ms", flush=True)
one = halves[0]
print(f"one B=4 half alone: {timeit(lambda: one.sample(x[:4], num_steps=50)):7.1f} ms")
