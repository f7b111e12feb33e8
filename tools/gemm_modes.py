"""Diagnostic: which smem-descriptor convention makes row-shifted (multi-tap, single A box)
UMMA operands correct?  usage: python tools/gemm_modes.py <single_load> <base_off_mode>"""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audio_diffusion_pytorch_b200 import _lib, ops  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
single, mode = int(sys.argv[1]), int(sys.argv[2])
L = _lib.lib()
L.adp_debug_set(1, single)
L.adp_debug_set(2, mode)
for (B, T, C, co) in [(2, 512, 64, 64), (2, 300, 128, 128), (1, 256, 32, 32), (1, 128, 16, 16),
                      (2, 1000, 256, 256)]:
    g = torch.Generator().manual_seed(0)
    x = torch.randn(B, T, C, generator=g).cuda().bfloat16()
    w = (torch.randn(co, C, 3, generator=g) * (3 * C) ** -0.5).cuda().bfloat16()
    out = torch.empty(B, T, co, device="cuda", dtype=torch.bfloat16)
    ops.conv_gemm(x, ops.pack_conv(w), out, c_in=C, n_valid=co, taps=(-1, 0, 1))
    ref = F.conv1d(x.float().transpose(1, 2), w.float(), padding=1).transpose(1, 2)
    err = (out.float() - ref).abs().max().item()
    print(f"single={single} base_off_mode={mode} C={C} T={T}: max err {err:.4f} "
          f"{'OK' if err < 0.05 else 'WRONG'}", flush=True)
for (B, T, ci, co, f) in [(2, 256, 64, 32, 4), (1, 128, 256, 128, 2), (2, 256, 32, 8, 4)]:
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, T, ci, generator=g).cuda().bfloat16()
    w = (torch.randn(co, ci, 3, generator=g) * (3 * ci) ** -0.5).cuda().bfloat16()
    out = torch.empty(B, T * f, co, device="cuda", dtype=torch.bfloat16)
    ops.conv_gemm(x, ops.pack_upsample_conv(w, f), out.view(B, T, f * co), c_in=ci, n_valid=co,
                  up_factor=f)
    up = F.interpolate(x.float().transpose(1, 2), scale_factor=f, mode="nearest")
    ref = F.conv1d(up, w.float(), padding=1).transpose(1, 2)
    err = (out.float() - ref).abs().max().item()
    print(f"single={single} base_off_mode={mode} up ci={ci} f={f}: max err {err:.4f} "
          f"{'OK' if err < 0.08 else 'WRONG'}", flush=True)
