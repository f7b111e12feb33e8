"""Graph-timed N-tile sweep of the 1-tap projection GEMMs of the README net (no stats / residual,
as the q|k|v projections run them).  usage: python tools/time_gemm_k1.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.time_gemm import run
shapes = [("L7 qkv", 2048, 1024, 1536), ("L8 qkv", 1024, 1024, 1536), ("L5 qkv", 8192, 512, 1536),
          ("L6 qkv", 4096, 512, 1536), ("L7 out", 2048, 512, 1024), ("L8 out", 1024, 512, 1024),
          ("L5 out", 8192, 512, 512), ("L6 out", 4096, 512, 512), ("L7 down", 2048, 1024, 1024),
          ("L8 down", 1024, 2048, 1024)]
for name, M, K, N in shapes:
    row = []
    for bn in (64, 128, 256):
        if N % bn:
            continue
        for res in (False, True):
            try:
                us, tf = run(M, K, N, 1, bn, res=res, stats=False)
                row.append(f"bn{bn}{'+res' if res else ''}: {us:5.1f}us")
            except Exception as e:
                row.append(f"bn{bn}: ERR")
    print(f"{name:8s} M={M:5d} K={K:4d} N={N:4d} | " + " | ".join(row), flush=True)
