import sys; sys.path.insert(0, '.')
from audio_diffusion_pytorch_b200 import _lib
from tools.time_gemm import run
L = _lib.lib()
for name, M, K, N, taps in [("L7 conv3", 2048, 1024, 1024, 3), ("L5 conv3", 8192, 512, 512, 3)]:
    for dis in (1, 0):
        for kc in (0, 2):
            row = []
            for dname, dbg in (("full", 0), ("noMMA", 1), ("noLOAD", 2), ("neither", 3), ("neither+nodrain", 7), ("floor", 8)):
                L.adp_debug_set(0, dis); L.adp_debug_set(4, dbg); L.adp_debug_set(5, kc)
                try:
                    us, tf = run(M, K, N, taps, 128, res=True, stats=False)
                    row.append(f"{dname}: {us:5.1f}")
                except Exception as e:
                    row.append(f"{dname}: ERR")
            L.adp_debug_set(0, 0); L.adp_debug_set(4, 0); L.adp_debug_set(5, 0)
            print(f"{name} {'single' if dis else 'pairs '} kc={kc or 'auto'} | " + " | ".join(row), flush=True)
