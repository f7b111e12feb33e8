"""Launches one kernel shape repeatedly (for `ncu --set full`).
usage: python tools/one_kernel.py conv M K N taps [block_n [stats]] | attn B H T | lnfilm B T C | gnsilu B T C | mid B T C"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audio_diffusion_pytorch_b200 import ops  # noqa: E402

kind = sys.argv[1]
a = [int(v) for v in sys.argv[2:]]
dev = "cuda"
reps = 6
if kind == "conv":
    M, K, N, taps = a[:4]
    bn = a[4] if len(a) > 4 else 0
    with_stats = len(a) > 5 and a[5] != 0
    B = 8
    x = torch.randn(B, M // B, K, device=dev).bfloat16()
    w = torch.randn(N, K, taps, device=dev) * (K * taps) ** -0.5
    out = torch.empty(B, M // B, N, device=dev, dtype=torch.bfloat16)
    res = torch.randn_like(out)
    bias = torch.randn(N, device=dev)
    wp = ops.pack_conv(w)
    tp = (-1, 0, 1) if taps == 3 else (0,)
    st = torch.zeros(B, 8, 2, device=dev, dtype=torch.float64) if with_stats else None
    for _ in range(reps):
        ops.conv_gemm(x, wp, out, c_in=K, n_valid=N, taps=tp, bias=bias, residual=res, stats=st, block_n=bn)
elif kind == "attn":
    B, H, T = a
    qkv = torch.randn(B, T, 3 * H * 64, device=dev).bfloat16()
    o = torch.empty(B, T, H * 64, device=dev, dtype=torch.bfloat16)
    m = H * 64
    for _ in range(reps):
        ops.attention(qkv[..., :m], qkv[..., m:2 * m], qkv[..., 2 * m:], o, H, 0.125)
elif kind == "lnfilm":
    B, T, C = a
    x = torch.randn(B, T, C, device=dev).bfloat16()
    y = torch.empty_like(x)
    ss = torch.randn(B, 2 * C, device=dev)
    st = torch.zeros(B, 8, 2, device=dev, dtype=torch.float64)
    for _ in range(reps):
        ops.ln_film(x, y, ss, 2 * C, st, 8, 1e-6)
elif kind == "gnsilu":
    B, T, C = a
    x = torch.randn(B, T, C, device=dev).bfloat16()
    y = torch.empty_like(x)
    st = torch.zeros(B, 8, 2, device=dev, dtype=torch.float64)
    ops.gn_stats(x, st, 8)
    g, bb = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    for _ in range(reps):
        ops.gn_silu(x, y, st, g, bb, 8, 1e-5)
elif kind == "mid":
    B, T, C = a
    x = torch.randn(B, T, C, device=dev).bfloat16()
    y = torch.empty_like(x)
    res = torch.randn_like(x)
    st = torch.zeros(B, 8, 2, device=dev, dtype=torch.float64)
    ops.gn_stats(x, st, 8)
    so = torch.zeros(B, 8, 2, device=dev, dtype=torch.float64)
    g, bb = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    w = torch.randn(C, C, 3, device=dev) * (3 * C) ** -0.5
    bias = torch.randn(C, device=dev)
    ss = torch.randn(B, 2 * C, device=dev) * 0.3
    for _ in range(reps):
        ops.narrow_conv(x, y, st, g, bb, w, bias, 8, residual=res, scale_shift=ss, ss_stride=2 * C,
                        stats_out=so)
torch.cuda.synchronize()
print("done")
