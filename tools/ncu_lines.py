"""Per-source-line hot spots from an .ncu-rep (needs -lineinfo + --import-source on).
usage: ncu_lines.py report.ncu-rep [topN] [file-filter]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30; flt = sys.argv[3] if len(sys.argv) > 3 else ""
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file = None; hdr = None; data = []
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur_file = r[1]; continue
    if r and r[0] == "Line No": hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0].isdigit():
        ix = hdr.index("Instructions Executed"); sx = hdr.index("# Samples")
        try: data.append((cur_file.split("/")[-1], int(r[0]), r[1].strip()[:100], int(r[ix]), int(r[sx])))
        except ValueError: pass
ti = sum(d[3] for d in data) or 1; ts = sum(d[4] for d in data) or 1
print(f"total warp-inst {ti}  samples {ts}")
data = [d for d in data if flt in d[0]]
for d in sorted(data, key=lambda d: -d[4])[:top]:
    print(f"{d[4]/ts*100:5.1f}% smp {d[3]/ti*100:5.1f}% inst | {d[0]}:{d[1]} | {d[2]}")
