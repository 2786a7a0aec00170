"""Summarise an `ncu --page raw --csv` export: one row per kernel launch with the metrics the roofline needs."""
import csv, sys
fn = sys.argv[1]
rows = list(csv.reader(open(fn)))
hdr = rows[0]; units = rows[1]; data = rows[2:]
col = {h: i for i, h in enumerate(hdr)}
want = [
 ("gpu__time_duration.sum", "dur_us"),
 ("dram__bytes_read.sum", "dram_rd"),
 ("dram__bytes_write.sum", "dram_wr"),
 ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
 ("lts__t_bytes.sum", "l2_bytes"),
 ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
 ("sm__inst_executed.avg.per_cycle_elapsed", "ipc"),
 ("smsp__issue_active.avg.pct", "issue%"),
 ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
 ("launch__registers_per_thread", "regs"),
 ("launch__grid_size", "grid"),
 ("launch__block_size", "blk"),
 ("smsp__inst_executed.sum", "winst"),
 ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
 ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu%"),
 ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma%"),
 ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%"),
 ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "bankconf"),
]
def num(s):
    try: return float(s.replace(",", ""))
    except: return float("nan")
def scale(v, u):
    u = u.lower()
    if u in ("ms", "msecond"): return v * 1e3
    if u in ("ns", "nsecond"): return v / 1e3
    if u in ("s", "second"): return v * 1e6
    if u == "kbyte": return v * 1e3
    if u == "mbyte": return v * 1e6
    if u == "gbyte": return v * 1e9
    return v
print("| kernel | " + " | ".join(n for _, n in want if _ in col) + " |")
print("|---|" + "---|" * sum(1 for m, _ in want if m in col))
for r in data:
    name = r[col["Kernel Name"]][:40]
    out = []
    for m, n in want:
        if m not in col: continue
        v = scale(num(r[col[m]]), units[col[m]])
        out.append(f"{v:.4g}")
    print(f"| {name} | " + " | ".join(out) + " |")
stall = [h for h in hdr if h.startswith("smsp__average_warp") and "issue_stalled" in h and h.endswith("_per_warp_active.pct")] or \
        [h for h in hdr if "warps_issue_stalled" in h and h.endswith("per_warp_active.pct")]
if stall:
    print("\nTop stall reasons (pct of warp-active cycles) per kernel:")
    for r in data:
        name = r[col["Kernel Name"]][:40]
        s = sorted(((num(r[col[h]]), h.split("issue_stalled_")[1].split("_per")[0]) for h in stall), reverse=True)[:4]
        print(f"- {name}: " + ", ".join(f"{n} {v:.0f}%" for v, n in s))
