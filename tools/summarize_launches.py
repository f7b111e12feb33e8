"""Summarise an ncu launch-list CSV (last net evaluation only) -> text table."""
import collections, csv, re, sys
path, out = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None
rows = list(csv.DictReader(l for l in open(path) if not l.startswith("==")))
names = [(r["Kernel Name"], float(r["Metric Value"]), r["Grid Size"]) for r in rows]
idx = [i for i, (k, _, _) in enumerate(names) if "FillFunctor<double>" in k]
last = [x for x in names[idx[-1]:] if "adp::" in x[0]]
tot = sum(v for _, v, _ in last)
agg = collections.defaultdict(lambda: [0, 0.0])
for k, v, g in last:
    k = re.sub(r"\(.*", "", k).replace("void ", "")[:52] + " grid" + g.replace(" ", "")
    agg[k][0] += 1
    agg[k][1] += v
lines = [f"one eager net evaluation (README config, B=8, T=2^18): {len(last)} adp kernels, "
         f"sum of gpu__time_duration {tot / 1000:.1f} us"]
fam = collections.defaultdict(float)
for k, (c, v) in agg.items():
    fam[k.split("<")[0].split(" ")[0]] += v
lines.append("by family: " + ", ".join(f"{k} {v / 1000:.0f} us" for k, v in sorted(fam.items(), key=lambda kv: -kv[1])))
for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append(f"{k:84s} x{c:3d} {v / 1000:9.1f} us {100 * v / tot:5.1f}%  avg {v / c / 1000:7.1f}")
text = "\n".join(lines)
print("\n".join(lines[:40]))
if out:
    open(out, "w").write(text + "\n")
