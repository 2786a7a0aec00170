"""Print SASS lines of an `ncu --page source --print-source sass --csv` export with executed counts and stall samples."""
import csv, gzip, sys
fn = sys.argv[1]
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3]) if len(sys.argv) > 3 else 10**9
op = gzip.open if fn.endswith(".gz") else open
rows = list(csv.reader(op(fn, "rt")))
hdr = rows[1]; col = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
tot_inst = sum(int(r[col["Instructions Executed"]]) for r in data)
tot_samp = sum(int(r[col["# Samples"]]) for r in data)
print(f"# {rows[0][1][:80]}  lines={len(data)} inst={tot_inst} samples={tot_samp}")
for i, r in enumerate(data):
    if i < lo or i >= hi: continue
    ie = int(r[col["Instructions Executed"]]); s = int(r[col["# Samples"]])
    print(f"{i:5d} {ie:10d} {100.0*ie/tot_inst:5.2f}% s={100.0*s/max(tot_samp,1):5.2f}%  {r[col['Source']].strip()}")
