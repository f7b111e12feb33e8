#!/bin/bash
# Round-2 validation on one B200: every GPU test file in its own process, smoke(), the three bench
# configurations, the in-graph per-shape tables and the ncu captures of the thin-level kernels.
mkdir -p gpurun_out
bash tools/gpu_tests_files.sh > gpurun_out/tests_digest.txt 2>&1
grep -E "^==|FAILED|Error|timed out" gpurun_out/tests_digest.txt | head -40
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep smoke
python bench.py > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; tail -c 600 gpurun_out/bench_cfg2.err
python bench.py --config cfg3 --steps 2 > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err
python bench.py --config cfg5 --steps 2 > gpurun_out/bench_cfg5.json 2> gpurun_out/bench_cfg5.err
python - <<'PY'
import json
for c in ("cfg2", "cfg3", "cfg5"):
    try:
        d = json.load(open(f"gpurun_out/bench_{c}.json")); r = d["roofline"]
        print(c, "value %.1f e2e %.1f ms/eval %.3f | top %s frac %.3f (%.1f us; eager-event frac %.3f) step_frac %.3f busy %.0f us | cpu %s | train %s" % (
            d["value"], d["e2e"]["value"], d["ms_per_net_eval"], r["kernel"], r["frac"], r["kernel_us"],
            r["achieved_eager_events"] / r["peak"], r["step_frac"], r.get("kernel_busy_us_per_net_eval", -1),
            d.get("cpu_baseline", {}).get("value"), d.get("train_step", {}).get("ms_per_step")))
    except Exception as e:
        print(c, "ERR", e)
PY
if [ "$1" = "full" ]; then     # per-shape tables and ncu captures of the thin-level kernels
  for c in cfg2 cfg3 cfg5; do python tools/graph_profile.py $c 3 0 2>/dev/null > gpurun_out/graph_$c.txt; done
  bash tools/ncu_capture.sh mid_conv r2_ncu_mid_conv64 mid 8 16384 64
  bash tools/ncu_capture.sh mid_conv r2_ncu_mid_conv32 mid 8 65536 32
fi
