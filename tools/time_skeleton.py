"""Where does a conv GEMM launch spend its time?  Graph-timed (20 launches per graph) with the
adp_debug_set(4, bits) experiments: 1 skip MMAs, 2 skip TMA loads, 4 skip the TMEM drain,
8 return at kernel entry (launch floor).  usage: python tools/time_skeleton.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audio_diffusion_pytorch_b200 import _lib
from tools.time_gemm import run  # noqa

L = _lib.lib()
shapes = [("L1 conv3", 524288, 32, 32, 3, 32), ("L2 conv3", 131072, 64, 64, 3, 64),
          ("L7 conv3", 2048, 1024, 1024, 3, 128), ("L8 conv3", 1024, 1024, 1024, 3, 64),
          ("L5 conv3", 8192, 512, 512, 3, 128), ("L7 qkv", 2048, 1024, 1536, 1, 128),
          ("L7 k1 1024", 2048, 1024, 1024, 1, 128), ("L3 conv3", 32768, 128, 128, 3, 128)]
modes = [("full", 0), ("noMMA", 1), ("noLOAD", 2), ("neither", 3), ("neither+nodrain", 7), ("floor", 8)]
for name, M, K, N, taps, bn in shapes:
    for res, st in ((False, False), (True, True)):
        row = []
        for mname, dbg in modes:
            L.adp_debug_set(4, dbg)
            us, tf = run(M, K, N, taps, bn, res=res, stats=st)
            row.append(f"{mname}: {us:6.1f}us")
        L.adp_debug_set(4, 0)
        print(f"{name:11s} bn={bn:3d} res/stats={int(res)} | " + " | ".join(row), flush=True)
