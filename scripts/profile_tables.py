"""Turn the ncu exports of a profiling session into the small, tracked tables under profiles/rN/.

usage: python scripts/profile_tables.py <launches.csv> <full_raw.csv> <chunks per launch> <commit> <out dir>

  launch_list.md     per kernel: launches, mean / min device time per launch, share of the wave (ncu --metrics gpu__time_duration.sum
                     --clock-control none: cold-cache and serialised -- compare SHARES with bench.py's kernels_ms, not absolutes)
  ncu_full_summary.md  one row per kernel of one wave from `ncu --set full`: duration, DRAM bytes, IPC, occupancy, pipe utilisation, top stalls
  ncu_traffic.json   DRAM bytes per launch of the frontend kernels (bench.py reads it for roofline.traffic) and DRAM bytes per chunk of the wave
"""
import csv
import json
import os
import sys

launches, full_raw, chunks, commit, out = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4], sys.argv[5]
os.makedirs(out, exist_ok=True)


def num(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return float("nan")


def short(name):
    n = name.split("(")[0].replace("void ", "").replace("bn::", "").replace("<unnamed>::", "").replace("(anonymous namespace)::", "")
    return n.strip()


# ---- launch list ----------------------------------------------------------------------------------------------------
rows = [r for r in csv.reader(open(launches)) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
col = {h: i for i, h in enumerate(hdr)}
per = {}
order = []
for r in rows:
    if r is hdr or len(r) <= col["Metric Value"] or r[col["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = short(r[col["Kernel Name"]])
    v = num(r[col["Metric Value"]])
    unit = r[col["Metric Unit"]].lower()
    us = v / 1e3 if unit.startswith("n") else (v * 1e3 if unit.startswith("m") else v)
    key = (name, r[col["Grid Size"]] if "Grid Size" in col else "")
    if key not in per:
        per[key] = []
        order.append(key)
    per[key].append(us)
tot = sum(sum(v) / len(v) for v in per.values())
with open(os.path.join(out, "launch_list.md"), "w") as fh:
    fh.write(f"# ncu launch list, commit {commit}: `python profiles/run_wave.py {chunks} 3` (one wave of {chunks} chunks per repetition)\n\n")
    fh.write("| kernel | grid | launches | mean us / launch | min us | share of the wave |\n|---|---|---|---|---|---|\n")
    for k in order:
        v = per[k]
        m = sum(v) / len(v)
        fh.write(f"| `{k[0]}` | {k[1]} | {len(v)} | {m:.1f} | {min(v):.1f} | {100 * m / tot:.1f} % |\n")
    fh.write(f"\nSum of the means: {tot / 1e3:.2f} ms per wave = {tot / chunks * 1e3:.0f} ns per chunk (cold-cache, serialised launches).\n")

# ---- full summary ---------------------------------------------------------------------------------------------------
rows = list(csv.reader(open(full_raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
want = [("gpu__time_duration.sum", "dur us"), ("dram__bytes_read.sum", "DRAM rd MB"), ("dram__bytes_write.sum", "DRAM wr MB"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"), ("sm__inst_executed.avg.per_cycle_elapsed", "IPC"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ %"), ("launch__registers_per_thread", "regs"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("smsp__inst_executed.sum", "warp inst"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu %"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma %"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu %")]


def scaled(r, m):
    v, u = num(r[col[m]]), units[col[m]].lower()
    if m == "gpu__time_duration.sum":
        return v * 1e3 if u.startswith("ms") else (v / 1e3 if u.startswith("ns") else (v * 1e6 if u in ("s", "second") else v))
    if "bytes" in m:
        f = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
        return v * f / 1e6
    return v


stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("per_issue_active.ratio")]
traffic = {}
total_dram = 0.0
with open(os.path.join(out, "ncu_full_summary.md"), "w") as fh:
    fh.write(f"# `ncu --set full --clock-control none`, one wave of {chunks} chunks, commit {commit}\n\n")
    fh.write("| kernel | " + " | ".join(n for m, n in want if m in col) + " | top stalls (cycles per issue) |\n")
    fh.write("|---|" + "---|" * (sum(1 for m, _ in want if m in col) + 1) + "\n")
    for r in data:
        name = short(r[col["Kernel Name"]])
        cells = []
        for m, n in want:
            if m in col:
                v = scaled(r, m)
                cells.append(f"{v:.4g}")
        st = sorted(((num(r[col[h]]), h.split("stalled_")[1].split("_per")[0]) for h in stall), reverse=True)[:4]
        fh.write(f"| `{name}` | " + " | ".join(cells) + " | " + ", ".join(f"{n} {v:.2f}" for v, n in st) + " |\n")
        rd, wr = scaled(r, "dram__bytes_read.sum") * 1e6, scaled(r, "dram__bytes_write.sum") * 1e6
        total_dram += rd + wr
        traffic.setdefault(name, {"dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr, "chunks_per_launch": chunks,
                                  "duration_us": scaled(r, "gpu__time_duration.sum")})
    fh.write(f"\nDRAM traffic of the whole wave: {total_dram / 1e9:.2f} GB = {total_dram / chunks / 1e3:.0f} KB per chunk "
             f"(algorithmic: 144.4 KB per chunk).\n")
alias = {"K1q_stft_quant": "k_stft_q<0>", "K2q_head": "k_head_q", "K1_stft": "k_stft_mag<1, 0>", "K2tc_head": "k_head_tc"}
outj = {"commit": commit, "what": f"ncu --set full --clock-control none, python profiles/run_wave.py {chunks} 1", "dram_bytes_per_chunk_whole_wave": total_dram / chunks}
for bench_name, kname in alias.items():
    for k, v in traffic.items():
        if k.replace(" ", "").startswith(kname.replace(" ", "")):
            outj[bench_name] = v
json.dump(outj, open(os.path.join(out, "ncu_traffic.json"), "w"), indent=1)
print(open(os.path.join(out, "launch_list.md")).read())
