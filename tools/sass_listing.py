"""SASS of the hot kernels out of the built library into profiles/ (one file per kernel, encodings stripped):
python tools/sass_listing.py rNN"""
import os, re, subprocess, sys
R = sys.argv[1] if len(sys.argv) > 1 else "r02"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "bonej2_b200", "libbonejthickness.so")
want = {"edt_tile_z_tma_kernelILb1ELi16": "edt_tile_z_tma_kernel", "ridge_tma_kernel": "ridge_tma_kernel", "finish_tma_kernel": "finish_tma_kernel",
        "disc_kernel": "disc_kernel", "zfront_kernel": "zfront_kernel", "copy2d_batch_kernel": "copy2d_batch_kernel",
        "edt_x_kernelILi1ELb1": "edt_x_kernel", "edt_tile_kernelItLi16": "edt_tile_kernel_y",
        "edt_tile16_kernelItLb0ELi3ELi16": "edt_tile16_kernel_y", "edt_tile16_kernelIjLb1ELi3ELi16": "edt_tile16_kernel_z"}
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
cur, buf = None, {}
for ln in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        cur = next((v for k, v in want.items() if k in m.group(1)), None)
        if cur:
            buf[cur] = ["Function : " + m.group(1)]
        continue
    if cur and re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
        buf[cur].append(re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", ln).rstrip())
for name, lines in buf.items():
    p = os.path.join(ROOT, "profiles", f"{R}_sass_{name}.txt")
    marks = sorted({t for l in lines for t in ("UTMALDG", "SYNCS", "ATOMS", "VIADDMNMX.U16x2", "VIMNMX3", "VOTE", "LDS", "LDG", "STG", "MUFU") if t in l})
    with open(p, "w") as f:
        f.write(f"# cuobjdump -sass bonej2_b200/libbonejthickness.so (sm_100a), {len(lines) - 1} instructions; mnemonics present: {' '.join(marks)}\n")
        f.write("\n".join(lines) + "\n")
    print(name, len(lines) - 1, " ".join(marks))
