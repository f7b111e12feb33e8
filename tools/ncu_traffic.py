"""Builds profiles/ncu_traffic.json -- DRAM bytes per launch of the profiled kernel shapes -- from
the committed `ncu --set full` summaries (`name = value unit` lines written by tools/ncu_raw.sh /
tools/ncu_capture.sh).  bench.py reads `roofline.traffic` from that file.

    python tools/ncu_traffic.py            # re-scan profiles/*.txt

The mapping capture-file -> bench kernel label is the table below (a capture is taken for one
kernel shape with tools/one_kernel.py, whose arguments are that shape)."""
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROF = os.path.join(ROOT, "profiles")
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}

# bench / graph-profile label -> capture summary (newest generation of that shape first)
CAPTURES = {
    "conv_gemm[k3 M=2048 K=1024 N=1024x1]": ["r2_ncu_conv_L7.txt", "r1_ncu_conv_L7_v4.txt"],
    "conv_gemm[k3 M=32768 K=128 N=128x1]": ["r2_ncu_conv_L3.txt", "r1_ncu_conv_L3_v4.txt"],
    "conv_gemm[k3 M=8192 K=512 N=512x1]": ["r2_ncu_conv_L5.txt"],
    "narrow_conv[M=524288 C=32 +res+film]": ["r2_ncu_mid_conv32.txt", "r1_ncu_mid_conv32.txt"],
    "narrow_conv[M=131072 C=64 +res+film]": ["r2_ncu_mid_conv64.txt", "r1_ncu_mid_conv64.txt"],
    "attention[B=8 H=8 Tq=1024 Tk=1024]": ["r2_ncu_attn.txt"],
    "gn_silu[M=2048 C=1024]": ["r2_ncu_gn_silu.txt"],
    "ln_film[M=2048 C=1024]": ["r2_ncu_ln_film.txt"],
}


def dram_bytes(path):
    tot, seen = 0.0, 0
    for line in open(path):
        m = re.match(r"\s*dram__bytes_(read|write)\.sum\s*=?\s*([0-9.eE+-]+)\s*(\w+)?", line)
        if m and (m.group(3) or "byte") in UNIT:
            tot += float(m.group(2)) * UNIT[m.group(3) or "byte"]
            seen += 1
            continue
        m = re.match(r"\s*dram__bytes_(read|write)\.sum\s+(\w+)\s+([0-9.eE+-]+)", line)   # column form
        if m and m.group(2) in UNIT:
            tot += float(m.group(3)) * UNIT[m.group(2)]
            seen += 1
    return tot if seen >= 2 else None


def main():
    out = {}
    for label, files in CAPTURES.items():
        for f in files:
            p = os.path.join(PROF, f)
            if os.path.exists(p):
                b = dram_bytes(p)
                if b is not None:
                    out[label] = {"dram_bytes": b, "source": "profiles/" + f}
                    break
    with open(os.path.join(PROF, "ncu_traffic.json"), "w") as fh:
        json.dump(out, fh, indent=1, sort_keys=True)
    for k, v in out.items():
        print(f"{k:46s} {v['dram_bytes'] / 1e6:9.3f} MB  {v['source']}")


if __name__ == "__main__":
    main()
