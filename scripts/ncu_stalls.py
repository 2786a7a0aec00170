"""Aggregate stall reasons of an `ncu --page source --print-source sass --csv` export; list the hottest SASS lines."""
import csv, gzip, sys
fn = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
op = gzip.open if fn.endswith(".gz") else open
rows = list(csv.reader(op(fn, "rt")))
hdr = rows[1]; col = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) >= len(hdr) - 2 and r[col["Instructions Executed"]].isdigit()]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = {s: 0 for s in stalls}
for r in data:
    for s in stalls:
        try: tot[s] += int(r[col[s]])
        except: pass
allsamp = sum(tot.values())
tot_inst = sum(int(r[col["Instructions Executed"]]) for r in data)
print(f"# {rows[0][1][:90]} lines={len(data)} warp-inst={tot_inst} samples={allsamp}")
print("  ".join(f"{s[6:]}={100*v/allsamp:.1f}%" for s, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v * 100 > allsamp))
# opcode histogram
ops = {}
for r in data:
    src = r[col["Source"]].strip()
    toks = src.split()
    o = toks[1] if toks[0].startswith("@") else toks[0]
    o = o.split(".")[0]
    ops[o] = ops.get(o, 0) + int(r[col["Instructions Executed"]])
print("  ".join(f"{o}={100*v/tot_inst:.1f}%" for o, v in sorted(ops.items(), key=lambda kv: -kv[1])[:24]))
idx = sorted(range(len(data)), key=lambda i: -int(data[i][col["# Samples"]]))[:top]
for i in sorted(idx):
    r = data[i]
    s = int(r[col["# Samples"]]); ie = int(r[col["Instructions Executed"]])
    why = sorted(((int(r[col[x]] or 0), x[6:]) for x in stalls), reverse=True)[:2]
    print(f"{i:5d} inst={100.0*ie/tot_inst:5.2f}% samp={100.0*s/allsamp:5.2f}% {why[0][1]}/{why[1][1]}  {r[col['Source']].strip()[:90]}")
