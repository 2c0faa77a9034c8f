"""Executed warp-instructions by SASS opcode (top N) from an .ncu-rep.  usage: ncu_opcodes.py rep [N]"""
import csv, subprocess, sys, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None; cnt = collections.Counter(); tot = 0
for r in rows:
    if "Source" in r and "Instructions Executed" in r: hdr = r; continue
    if hdr and len(r) == len(hdr):
        try: n = int(r[hdr.index("Instructions Executed")])
        except ValueError: continue
        toks = r[hdr.index("Source")].split()
        op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")
        cnt[op.split(".")[0] + ("." + op.split(".")[1] if "." in op and op.split(".")[0] in ("MUFU","LDS","STS","LDG","STG","RED","ATOMS","SHFL") else "")] += n; tot += n
print("total", tot)
for op, n in cnt.most_common(top): print(f"{n/tot*100:5.1f}%  {n:>12}  {op}")
