"""Turn gpurun_out/ captures into the tracked summaries under profiles/ (round tag as argv[1])."""
import csv, os, subprocess, sys, json
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
os.makedirs("profiles", exist_ok=True)
# 1. launch list of the bench step
src = "gpurun_out/launches.csv"
if os.path.exists(src):
    rows = list(csv.reader(open(src)))
    hdr = [r for r in rows if "Kernel Name" in r][0]
    i0 = rows.index(hdr); kn = hdr.index("Kernel Name"); mv = hdr.index("Metric Value")
    body = rows[i0 + 1:]
    # one full step = from one k2_count_valid to the next
    idx = [i for i, r in enumerate(body) if "k2_count_valid" in r[kn]]
    step = body[idx[0]:idx[1]] if len(idx) >= 2 else body
    tot = sum(float(r[mv].replace(",", "")) for r in step)
    with open(f"profiles/{tag}_launch_list.md", "w") as f:
        f.write(f"# {tag}: every launch of ONE bench step (cfg2 G-A, B=16, C=150), `ncu --metrics gpu__time_duration.sum --clock-control none`\n\n")
        f.write("command: `python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline` (per-launch times are cold-cache and serialised: compare shares)\n\n")
        f.write("| us | share | kernel |\n|---:|---:|---|\n")
        for r in step:
            t = float(r[mv].replace(",", "")) / 1000
            f.write(f"| {t:.1f} | {t * 1000 / tot * 100:.1f}% | `{r[kn][:110]}` |\n")
        f.write(f"\nsum = {tot / 1000:.1f} us over {len(step)} launches\n")
    os.system(f"cp {src} profiles/{tag}_launch_list.csv")
# 2. ncu --set full summaries
for name, rep in (("k2", "gpurun_out/k2_r1.ncu-rep"), ("k3full", "gpurun_out/k3full_r1.ncu-rep"), ("k1", "gpurun_out/k1_r1.ncu-rep")):
    if not os.path.exists(rep): continue
    out = subprocess.run([sys.executable, "tools/ncu_summary.py", rep], capture_output=True, text=True).stdout
    ops = subprocess.run([sys.executable, "tools/ncu_opcodes.py", rep, "16"], capture_output=True, text=True).stdout
    lines = subprocess.run([sys.executable, "tools/ncu_lines.py", rep, "15"], capture_output=True, text=True).stdout
    with open(f"profiles/{tag}_{name}_ncu_full.txt", "w") as f:
        f.write(f"# {tag} {name}: ncu --set full --clock-control none --import-source on (one launch)\n\n## metrics\n{out}\n## executed SASS opcodes\n{ops}\n## hottest source lines (stall samples)\n{lines}")
print(os.listdir("profiles"))
