"""per-source-line shares of one launch of an .ncu-rep when ncu's own source page carries no line correlation: the SASS
page (execution counts per instruction, in order) is zipped with `nvdisasm -g` of the kernel's cubin (the same
instructions in the same order, annotated with //## File ..., line N).
usage: python tools/line_shares.py rep launch_index cubin kernel_substring [min_pct]"""
import csv, re, subprocess, sys
rep, k, cubin, kern = sys.argv[1], int(sys.argv[2]), sys.argv[3], sys.argv[4]
thr = float(sys.argv[5]) if len(sys.argv) > 5 else 1.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--launch-skip", str(k), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
H = rows[1]
R = [r for r in rows[2:] if r and r[0].startswith("0x")]
R = R[:len(R) // 2] if len(R) > 1 and R[0][0] == R[len(R) // 2][0] else R
ie, it, ism = H.index("Instructions Executed"), H.index("Thread Instructions Executed"), H.index("# Samples")
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
# the kernel's section
start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and kern in l)
lines, cur, fn = [], None, None
for l in dis[start + 1:]:
    if l.startswith(".text.") or l.startswith("//--------------------- .nv"):
        if lines:
            break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        lines.append(cur)
assert len(lines) >= len(R), (len(lines), len(R))
agg = {}
for r, loc in zip(R, lines):
    a = agg.setdefault(loc, [0, 0, 0])
    a[0] += int(r[ie] or 0); a[1] += int(r[it] or 0); a[2] += int(r[ism] or 0)
tot = sum(a[0] for a in agg.values()); tots = sum(a[2] for a in agg.values())
print(f"{rows[0][1][:80]}: warp instructions {tot}, samples {tots}")
src = {}
for loc, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    if loc is None or 100.0 * a[0] / tot < thr and 100.0 * a[2] / max(tots, 1) < thr:
        continue
    f, n = loc
    if f not in src:
        try:
            src[f] = open(f"/root/repo/bonej2_b200/csrc/{f}").read().splitlines()
        except OSError:
            src[f] = []
    text = src[f][n - 1].strip()[:100] if 0 < n <= len(src[f]) else ""
    print(f"{f}:{n:<5d} {100.0*a[0]/tot:5.1f}% lanes {a[1]/max(a[0],1):4.1f} smp {100.0*a[2]/max(tots,1):5.1f}%  {text}")
