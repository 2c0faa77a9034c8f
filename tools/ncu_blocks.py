"""SASS of one kernel of an .ncu-rep grouped into blocks of equal execution count (= loop nests): instructions, samples,
top opcodes.  usage: ncu_blocks.py report.ncu-rep kernel-regex [min_share]"""
import csv, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]; thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.004
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
h = rows[hi]; ix = {k: i for i, k in enumerate(h)}
st = [c for c in h if c.startswith("stall_") and "Not" not in c]
blocks = []
for r in rows[hi + 1:]:
    if len(r) != len(h): continue
    ex = int(r[ix["Instructions Executed"]]); sm = int(r[ix["# Samples"]]); op = r[ix["Source"]].split()
    op = [o for o in op if not o.startswith("@")][0] if op else "?"
    if not blocks or blocks[-1]["ex"] != ex: blocks.append({"ex": ex, "n": 0, "smp": 0, "ops": {}, "addr": r[0], "st": {}})
    b = blocks[-1]; b["n"] += 1; b["smp"] += sm; k = op.split(".")[0]; b["ops"][k] = b["ops"].get(k, 0) + 1
    for c in st: b["st"][c[6:]] = b["st"].get(c[6:], 0) + int(r[ix[c]])
ti = sum(b["ex"] * b["n"] for b in blocks); ts = sum(b["smp"] for b in blocks)
print(f"total warp-inst {ti}  samples {ts}")
tot = {}
for b in blocks:
    for k, v in b["st"].items(): tot[k] = tot.get(k, 0) + v
print("stalls:", ", ".join(f"{k} {v / max(sum(tot.values()), 1) * 100:.1f}%" for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]))
for i, b in enumerate(blocks):
    w = b["ex"] * b["n"]
    if w / ti > thr or b["smp"] / ts > thr:
        top = sorted(b["ops"].items(), key=lambda kv: -kv[1])[:5]
        s3 = sorted(b["st"].items(), key=lambda kv: -kv[1])[:3]
        print(f"{i:4d} {b['addr'][-5:]} n={b['n']:4d} ex={b['ex']:9d} inst={w / ti * 100:5.1f}% smp={b['smp'] / ts * 100:5.1f}%  {top}  {[(k, v) for k, v in s3]}")
