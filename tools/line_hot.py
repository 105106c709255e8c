"""per-SOURCE-LINE execution shares of one launch in an .ncu-rep, by joining ncu's SASS page (instruction order) with
nvdisasm's line info of the same function in the built library:
python tools/line_hot.py rep launch_index cubin_name kernel_substring [min_pct]
e.g. python tools/line_hot.py gpurun_out/x.ncu-rep 1 paint_disc disc_kernel 1.0"""
import csv, os, re, subprocess, sys, tempfile
rep, k, cub, ker = sys.argv[1], int(sys.argv[2]), sys.argv[3], sys.argv[4]
thr = float(sys.argv[5]) if len(sys.argv) > 5 else 1.0
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", cub + ".sm_100a.cubin", os.path.join(root, "bonej2_b200", "libbonejthickness.so")], cwd=tmp, check=True,
               stdout=subprocess.DEVNULL)
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub + ".sm_100a.cubin")], capture_output=True, text=True).stdout
lines, cur, infn = [], None, False
for ln in dis.splitlines():
    m = re.match(r"\s*\.text\.(\S+):", ln) or re.match(r"\s*//-+ \.text\.(\S+)", ln)
    if m:
        infn = ker in m.group(1)
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4}\*/", ln):
        lines.append(cur)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--launch-skip", str(k), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
H = rows[1]; R = [r for r in rows[2:] if r and r[0].startswith("0x")]
R = R[:len(R) // 2] if len(R) > 1 and R[0][0] == R[len(R) // 2][0] else R
ie, it, ism = H.index("Instructions Executed"), H.index("Thread Instructions Executed"), H.index("# Samples")
print(rows[0][1][:90], "| static instructions: ncu", len(R), "nvdisasm", len(lines))
agg = {}
for i, r in enumerate(R):
    key = lines[i] if i < len(lines) else None
    a = agg.setdefault(key, [0, 0, 0])
    a[0] += int(r[ie]); a[1] += int(r[it]); a[2] += int(r[ism])
tot = sum(a[0] for a in agg.values()); tots = sum(a[2] for a in agg.values())
print("warp instructions", tot, "samples", tots)
src = {}
for key, a in sorted(agg.items(), key=lambda kv: (kv[0] or ("", 0))):
    if 100.0 * a[0] / tot >= thr or 100.0 * a[2] / tots >= thr:
        text = ""
        if key:
            if key[0] not in src:
                p = os.path.join(root, "bonej2_b200", "csrc", key[0])
                src[key[0]] = open(p).read().splitlines() if os.path.exists(p) else []
            text = src[key[0]][key[1] - 1].strip()[:90] if key[1] - 1 < len(src[key[0]]) else ""
        print(f"{str(key):32s} {100.0*a[0]/tot:5.2f}% thr {a[1]/max(a[0],1):5.1f} smp {100.0*a[2]/tots:5.2f}%  {text}")
