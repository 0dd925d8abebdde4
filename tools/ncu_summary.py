"""Condense an `ncu --set full` report into the handful of counters DESIGN.md argues from.
    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > raw.csv ; python tools/ncu_summary.py raw.csv > profiles/rNN/x.md
"""
import csv
import sys

KEYS = [("gpu__time_duration.sum", "duration"),
        ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs/thread"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "active lanes / inst"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe active %"),
        ("sm__sass_thread_inst_executed_op_dfma_pred_on.sum", "DFMA thread-inst"),
        ("sm__sass_thread_inst_executed_op_dmul_pred_on.sum", "DMUL thread-inst"),
        ("sm__sass_thread_inst_executed_op_dadd_pred_on.sum", "DADD thread-inst"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_active", "L1TEX throughput %"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
        ("lts__t_requests_srcunit_tex_op_red.sum", "L2 RED requests"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM written"),
        ("sass__inst_executed_local_loads", "local loads (inst)"), ("sass__inst_executed_local_stores", "local stores (inst)")]

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
print("| kernel | " + " | ".join(n for _, n in KEYS) + " |")
print("|---|" + "---|" * len(KEYS))
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")].replace("void ", "").replace("(ClassArgs)", "").replace("(DigestArgs)", "")
    cells = []
    for k, _ in KEYS:
        if k in hdr:
            i = hdr.index(k)
            v = r[i]
            try:
                v = f"{float(v.replace(',', '')):.4g}"
            except ValueError:
                pass
            cells.append(f"{v} {units[i]}".strip())
        else:
            cells.append("n/a")
    print(f"| `{name}` | " + " | ".join(cells) + " |")
