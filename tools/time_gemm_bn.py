"""Graph-timed A/B of the N tile (and drain-warp variant) of adp_conv_gemm on the deep-level shapes
of the README net at B=8.  usage: python tools/time_gemm_bn.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audio_diffusion_pytorch_b200 import _lib
from tools.time_gemm import run
L = _lib.lib()
shapes = [("L7 conv3", 2048, 1024, 1024, 3), ("L8 conv3", 1024, 1024, 1024, 3), ("L5 conv3", 8192, 512, 512, 3),
          ("L6 conv3", 4096, 512, 512, 3), ("L4 conv3", 16384, 256, 256, 3), ("L3 conv3", 32768, 128, 128, 3),
          ("L7 qkv", 2048, 1024, 1536, 1), ("L5 qkv", 8192, 512, 1536, 1), ("L7 out", 2048, 512, 1024, 1),
          ("L5 out", 8192, 512, 512, 1)]
for name, M, K, N, taps in shapes:
    row = []
    for bn in (64, 128, 256):
        if N % bn:
            continue
        for noew8 in (0, 1):
            L.adp_debug_set(7, noew8)
            try:
                us, tf = run(M, K, N, taps, bn, res=True, stats=True)
                row.append(f"bn{bn}{'/ew4' if noew8 else ''}: {us:6.1f}us {tf:5.0f}TF")
            except Exception as e:
                row.append(f"bn{bn}: ERR {str(e)[:30]}")
    L.adp_debug_set(7, 0)
    print(f"{name:9s} | " + " | ".join(row), flush=True)
