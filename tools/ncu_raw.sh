#!/bin/bash
# one `ncu --set full` capture printed as raw CSV (no .ncu-rep to bring back)
# usage: tools/ncu_raw.sh <kernel-regex> <out-name> <one_kernel.py args...>
mkdir -p gpurun_out
re=$1; out=$2; shift 2
ncu --clock-control none --set full -k regex:$re -s 2 -c 1 --page raw --csv python tools/one_kernel.py "$@" 2>/dev/null | grep -v "^==" > gpurun_out/$out.csv
python - "$out" <<'PY'
import csv, sys
rows = list(csv.reader(open(f"gpurun_out/{sys.argv[1]}.csv")))
rows = [r for r in rows if len(r) > 10]
hdr, units, vals = rows[0], rows[1], rows[2]
want = ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size",
        "launch__occupancy_limit", "launch__registers", "sm__warps_active.avg.pct", "smsp__pcsamp_warps_issue_stalled",
        "sm__inst_executed_pipe", "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "smsp__issue_active.avg.pct",
        "l1tex__data_bank_conflicts", "smsp__inst_executed_op_shared", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum")
with open(f"gpurun_out/{sys.argv[1]}.txt", "w") as f:
    for h, u, v in zip(hdr, units, vals):
        if h.startswith(want) and not h.endswith("_not_issued") and "per_second" not in h and "pct_of_peak_sustained_elapsed" not in h:
            f.write(f"{h} = {v} {u}\n")
PY
rm -f gpurun_out/$out.csv
wc -l gpurun_out/$out.txt
