#!/bin/bash
# `ncu --set full` of one kernel shape -> gpurun_out/<name>.txt (name = value unit lines of the
# metrics the review asks for).  usage: tools/ncu_capture.sh <kernel-regex> <name> <one_kernel.py args...>
mkdir -p gpurun_out
re=$1; out=$2; shift 2
ncu --clock-control none --set full -k regex:$re -s 3 -c 1 --page raw --csv python tools/one_kernel.py "$@" 2>/dev/null | grep -v "^==" > gpurun_out/$out.csv
python - "$out" "$*" <<'PY'
import csv, sys
rows = [r for r in csv.reader(open(f"gpurun_out/{sys.argv[1]}.csv")) if len(r) > 10]
hdr, units, vals = rows[0], rows[1], rows[2]
want = ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "launch__grid_size", "launch__block_size", "launch__occupancy_limit",
        "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active", "sm__pipe_tensor_op", "sm__inst_executed_pipe_tensor", "sm__inst_executed_pipe_uniform",
        "sm__cycles_elapsed.max", "sm__cycles_active.avg", "smsp__issue_active.avg.pct", "smsp__inst_executed.sum",
        "smsp__pcsamp_warps_issue_stalled", "l1tex__data_bank_conflicts_pipe_lsu.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__throughput.avg.pct", "gpu__compute_memory_throughput",
        "dram__throughput.avg.pct", "lts__throughput.avg.pct", "sm__inst_executed_pipe_xu", "sm__inst_executed_pipe_fma",
        "sm__pipe_fma_cycles_active", "sm__pipe_alu_cycles_active", "sm__pipe_xu_cycles_active")
with open(f"gpurun_out/{sys.argv[1]}.txt", "w") as f:
    f.write(f"# ncu --set full --clock-control none, one launch of: tools/one_kernel.py {sys.argv[2]}\n")
    for h, u, v in zip(hdr, units, vals):
        if h.startswith(want) and not h.endswith("_not_issued") and "per_second" not in h:
            f.write(f"{h} = {v} {u}\n")
PY
rm -f gpurun_out/$out.csv
wc -l gpurun_out/$out.txt
