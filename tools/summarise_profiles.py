"""Turn the ncu outputs of tools/profile_round.sh into tracked summaries under profiles/.
usage: python tools/summarise_profiles.py rNN"""
import csv, re, json, os, shutil, subprocess, sys
R = sys.argv[1] if len(sys.argv) > 1 else "r01"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)
f = lambda s: float(str(s).replace(",", "") or 0)

# 1. launch list of the bench command -> per-kernel totals and shares
src = os.path.join(G, f"{R}_bench_launches.csv")
rows = list(csv.reader(open(src)))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[hdr]
iK, iM, iV = H.index("Kernel Name"), H.index("Metric Name"), H.index("Metric Value")
agg, order = {}, []
for r in rows[hdr + 1:]:
    if len(r) <= iV or r[iM] != "gpu__time_duration.sum":
        continue
    k = r[iK].replace("bjt::", "").replace("<unnamed>::", "")
    k = re.sub(r"\((bool|int)\)", "", k)                        # template arguments print as (int)-1
    k = k.split("(")[0]
    if k not in agg:
        agg[k] = [0, 0.0]; order.append(k)
    agg[k][0] += 1; agg[k][1] += f(r[iV]) / 1e6
tot = sum(v[1] for v in agg.values())
shutil.copy(src, os.path.join(P, f"{R}_bench_launches.csv"))
with open(os.path.join(P, f"{R}_bench_launch_summary.txt"), "w") as o:
    o.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none, command: python bench.py --steps 1 --warmup 1 --no-cpu\n")
    o.write(f"# cold-cache serialised launch times: compare SHARES, not absolutes.  total {tot:.1f} ms over {sum(v[0] for v in agg.values())} launches\n")
    o.write(f"{'kernel':70s} {'launches':>8s} {'ms':>10s} {'share':>7s}\n")
    for k in sorted(agg, key=lambda k: -agg[k][1]):
        o.write(f"{k[:70]:70s} {agg[k][0]:8d} {agg[k][1]:10.2f} {100*agg[k][1]/tot:6.1f}%\n")

# 2. full capture -> selected metrics per kernel
rep = os.path.join(G, f"{R}_full_512sp.ncu-rep")
if os.path.exists(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(out.splitlines()))
    H = rr[0]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
            "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed.avg.per_cycle_active",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
            "launch__registers_per_thread", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
            "launch__grid_size", "launch__block_size"]
    with open(os.path.join(P, f"{R}_full_512sp_summary.txt"), "w") as o:
        o.write("# ncu --set full --clock-control none, command: python tools/one_run.py 512 1 1 (512^3 phantom, Tb.Sp run)\n")
        units = rr[1]
        for r in rr[2:]:
            o.write("== " + r[H.index("Kernel Name")][:100] + "\n")
            for h, u, v in zip(H, units, r):
                if h in want:
                    o.write(f"   {h:65s} {v:>16s} {u}\n")
print(open(os.path.join(P, f"{R}_bench_launch_summary.txt")).read())
