"""per-source-line execution shares of one launch in an .ncu-rep (kernels compiled with -lineinfo, captured with --import-source on):
python tools/src_hot.py rep launch_index [min_pct]"""
import csv, subprocess, sys
rep, k = sys.argv[1], int(sys.argv[2]); thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda", "--launch-skip", str(k), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No" and len(r) > 3)
H = rows[hi]
ie, ism = H.index("Instructions Executed"), H.index("# Samples")
it = H.index("Thread Instructions Executed")
R = [r for r in rows[hi + 1:] if r and r[0].isdigit() and len(r) > ie]
tot = sum(int(r[ie] or 0) for r in R); tots = sum(int(r[ism] or 0) for r in R)
print("warp instructions", tot, "samples", tots)
for r in R:
    e = int(r[ie] or 0); s = int(r[ism] or 0)
    if tot and (100.0 * e / tot >= thr or (tots and 100.0 * s / tots >= thr)):
        print(f"{r[0]:>5} {100.0*e/tot:5.2f}% thr {int(r[it] or 0)/max(e,1):5.1f} smp {100.0*s/max(tots,1):5.2f}%  {r[1].strip()[:110]}")
