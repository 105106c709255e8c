"""per-instruction execution shares of one launch in an .ncu-rep: python tools/sass_hot.py rep launch_index [min_pct]"""
import csv, subprocess, sys
rep, k = sys.argv[1], int(sys.argv[2]); thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.1
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--launch-skip", str(k), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
print(rows[0][1][:100])
H = rows[1]; R = [r for r in rows[2:] if r and r[0].startswith("0x")]
R = R[:len(R) // 2] if len(R) > 1 and R[0][0] == R[len(R) // 2][0] else R
ia, ie, it, ism = H.index("Source"), H.index("Instructions Executed"), H.index("Thread Instructions Executed"), H.index("# Samples")
tot = sum(int(r[ie]) for r in R); tots = sum(int(r[ism]) for r in R)
print("warp instructions", tot, "samples", tots, "static instructions", len(R))
for i, r in enumerate(R):
    e = int(r[ie])
    if 100.0 * e / tot >= thr or 100.0 * int(r[ism]) / tots >= 4 * thr:
        print(f"{i:5d} {100.0*e/tot:5.2f}% thr {int(r[it])/max(e,1):5.1f} smp {100.0*int(r[ism])/tots:5.2f}%  {r[ia].strip()[:100]}")
