"""A/B of the thin-level ConvBlock kernels: 256 threads x 2 blocks/SM vs 128 threads x 4 blocks/SM
(adp_debug_set(8, ..)).  usage: python tools/time_mid_threads.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audio_diffusion_pytorch_b200 import ops, _lib

dev = "cuda"
G = 8


def run(B, T, C, fused, threads, iters=40):
    _lib.lib().adp_debug_set(8, threads)
    torch.manual_seed(0)
    x = torch.randn(B, T, C, device=dev).to(torch.bfloat16)
    y = torch.empty_like(x)
    st = torch.zeros(B, G, 2, dtype=torch.float64, device=dev)
    ops.gn_stats(x, st, G)
    so = torch.zeros_like(st)
    gamma, beta = torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev) * 0.1
    w, b = torch.randn(C, C, 3, device=dev) * 0.05, torch.randn(C, device=dev) * 0.1
    wp = ops.pack_mid_conv(w)
    ss = torch.randn(B, 2 * C, device=dev) * 0.1
    kw = dict(residual=x, scale_shift=ss, ss_stride=2 * C) if fused else {}
    big = torch.empty(64 * 1024 * 1024, device=dev)
    for _ in range(3):
        ops.narrow_conv(x, y, st, gamma, beta, w, b, G, stats_out=so, w_packed=wp, **kw)
    tot = 0.0
    for _ in range(iters):
        big.zero_()                                   # flush L2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.narrow_conv(x, y, st, gamma, beta, w, b, G, stats_out=so, w_packed=wp, **kw)
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters * 1e3, y.float()


for B in (8, 16):
    for T, C in ((65536, 32), (16384, 64)):
        for fused in (False, True):
            t256, y256 = run(B, T, C, fused, 256)
            t128, y128 = run(B, T, C, fused, 128)
            same = float((y256 - y128).abs().max())
            print(f"B={B} T={T} C={C} {'+res+film' if fused else '         '}: 256x2 {t256:7.1f} us   128x4 {t128:7.1f} us   "
                  f"ratio {t128 / t256:.3f}   max |diff| {same:.2e}", flush=True)
_lib.lib().adp_debug_set(8, 256)
