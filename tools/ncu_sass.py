"""Per-SASS-instruction samples and stall reasons from an .ncu-rep, aggregated by phase markers.
usage: ncu_sass.py report.ncu-rep [lo_idx hi_idx]   (instruction index range to list)"""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None; data = []
for r in rows:
    if r and r[0] == "Address": hdr = r; continue
    if hdr and len(r) == len(hdr): data.append(r)
ix = {k: hdr.index(k) for k in hdr}
stalls = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
tot = sum(int(r[ix["# Samples"]]) for r in data)
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3]) if len(sys.argv) > 3 else len(data)
print("total samples", tot, "instructions", len(data))
agg = {}
for i, r in enumerate(data[lo:hi]):
    n = int(r[ix["# Samples"]])
    top = sorted(((int(r[ix[s]]), s[6:]) for s in stalls), reverse=True)[:3]
    for v, s in top:
        pass
    for s in stalls:
        agg[s[6:]] = agg.get(s[6:], 0) + int(r[ix[s]])
    if len(sys.argv) > 4:
        print(f"{lo+i:5d} {n:5d} {int(r[ix['Instructions Executed']]):9d} {r[ix['Source']].strip()[:70]:70s} " + " ".join(f"{s}:{v}" for v, s in top if v))
print("range samples:", sum(int(r[ix["# Samples"]]) for r in data[lo:hi]), {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
if len(sys.argv) > 4 and sys.argv[4] == "blocks":
    step = 50
    for b in range(0, len(data), step):
        blk = data[b:b + step]
        n = sum(int(r[ix["# Samples"]]) for r in blk)
        ie = sum(int(r[ix["Instructions Executed"]]) for r in blk)
        ops = {}
        for r in blk:
            op = r[ix["Source"]].split()[0] if r[ix["Source"]].split() else "?"
            if op.startswith("@"): op = r[ix["Source"]].split()[1]
            ops[op] = ops.get(op, 0) + 1
        top = sorted(ops.items(), key=lambda kv: -kv[1])[:4]
        if n: print(f"{b:5d} smp {n:5d} ({n/tot*100:4.1f}%) inst {ie:10d}  {top}")
