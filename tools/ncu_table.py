"""Summarise an `ncu --page raw --csv` export: one row per kernel, the counters the roofline discussion needs.
(units as ncu 2025 exports them here: ms, Gbyte, Kbyte/block -- check row 2 of the CSV when the tool version changes)"""
import csv, re, sys

rows = list(csv.reader(open(sys.argv[1])))
h, data = rows[0], rows[2:]
col = {c: i for i, c in enumerate(h)}


def get(r, name, scale=1.0, fmt="{:.1f}"):
    if name not in col or r[col[name]] in ("", "n/a"):
        return "-"
    try:
        return fmt.format(float(r[col[name]].replace(",", "")) * scale)
    except ValueError:
        return r[col[name]]


COLS = [("ms", "gpu__time_duration.sum", 1, "{:.3f}"),
        ("tensor pipe %", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 1, "{:.1f}"),
        ("imma inst %", "sm__inst_executed_pipe_tensor_subpipe_imma.avg.pct_of_peak_sustained_active", 1, "{:.1f}"),
        ("dmma inst %", "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", 1, "{:.1f}"),
        ("fp64 pipe %", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", 1, "{:.1f}"),
        ("issue %", "smsp__issue_active.avg.pct_of_peak_sustained_active", 1, "{:.1f}"),
        ("warps act %", "sm__warps_active.avg.pct_of_peak_sustained_active", 1, "{:.1f}"),
        ("DRAM rd GB", "dram__bytes_read.sum", 1, "{:.3f}"),
        ("DRAM wr GB", "dram__bytes_write.sum", 1, "{:.3f}"),
        ("DRAM %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1, "{:.1f}"),
        ("L2 %", "lts__throughput.avg.pct_of_peak_sustained_elapsed", 1, "{:.1f}"),
        ("L1 %", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", 1, "{:.1f}"),
        ("regs", "launch__registers_per_thread", 1, "{:.0f}"),
        ("smem KB/CTA", "launch__shared_mem_per_block_dynamic", 1, "{:.1f}")]
print("| kernel | " + " | ".join(c[0] for c in COLS) + " |")
print("|---|" + "---|" * len(COLS))
for r in data:
    name = re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("mcacq::", "").replace("void ", "")
    print(f"| `{name[:44]}` | " + " | ".join(get(r, c[1], c[2], c[3]) for c in COLS) + " |")
if len(sys.argv) > 2:   # dump every column matching a regex for the kernels whose name matches argv[3]
    pat, kpat = re.compile(sys.argv[2]), re.compile(sys.argv[3] if len(sys.argv) > 3 else ".")
    for r in data:
        if not kpat.search(r[col["Kernel Name"]]):
            continue
        print("\n##", r[col["Kernel Name"]][:80])
        for c, i in col.items():
            if pat.search(c) and r[i] not in ("", "0", "n/a"):
                print(f"  {c} = {r[i]} {rows[1][i]}")
