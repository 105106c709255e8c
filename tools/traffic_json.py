"""profiles/ystage_traffic.json from the captures of tools/capture_traffic.sh: python tools/traffic_json.py SIZE"""
import csv, json, os, re, subprocess, sys
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows_out = []
for inv in (0, 1):
    rep = os.path.join(ROOT, "gpurun_out", f"traffic_y_{inv}.ncu-rep")
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    H, U = rows[0], rows[1]
    for r in rows[2:]:
        if not re.search(r"sep_dyn_kernel<[^,]*, (\(int\))?2,", r[H.index("Kernel Name")]):
            continue                                              # x-stage launches of the same template
        def val(name):
            i = H.index(name)
            v = float(r[i].replace(",", ""))
            unit = U[i].lower()
            return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1)
        rows_out.append({"run": "Tb.Sp" if inv else "Tb.Th", "kernel": r[H.index("Kernel Name")][:60],
                         "ms": float(r[H.index("gpu__time_duration.sum")].replace(",", "")),
                         "dram_read": val("dram__bytes_read.sum"), "dram_write": val("dram__bytes_write.sum")})
tot = sum(x["dram_read"] + x["dram_write"] for x in rows_out)
rec = {"kernel": "sep_dyn_kernel<PACKED,2,+-1>", "volume": [n, n, n], "launches": len(rows_out),
       "dram_bytes_per_launch": tot / max(1, len(rows_out)),
       "source": "ncu --set full --clock-control none, tools/capture_traffic.sh (one Tb.Th and one Tb.Sp run)", "per_launch": rows_out}
json.dump(rec, open(os.path.join(ROOT, "profiles", "ystage_traffic.json"), "w"), indent=1)
print(json.dumps({k: v for k, v in rec.items() if k != "per_launch"}))
