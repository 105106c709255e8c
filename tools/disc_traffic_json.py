"""profiles/disc_traffic.json from the captures of tools/profile_round.sh (ncu dram bytes of disc_kernel launches of one
Tb.Th and one Tb.Sp run at the bench size): python tools/disc_traffic_json.py rNN SIZE"""
import csv, json, os, sys
R = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
per = []
for inv in (0, 1):
    rows = list(csv.reader(open(os.path.join(ROOT, "gpurun_out", f"{R}_disc_traffic_{inv}.csv"))))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[hdr]
    iID, iM, iU, iV = H.index("ID"), H.index("Metric Name"), H.index("Metric Unit"), H.index("Metric Value")
    launches = {}
    for r in rows[hdr + 1:]:
        if len(r) <= iV:
            continue
        v = float(r[iV].replace(",", ""))
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(r[iU], 1)
        launches.setdefault(r[iID], {})[r[iM]] = v
    for d in launches.values():
        per.append({"run": "Tb.Sp" if inv else "Tb.Th", "ms": d.get("gpu__time_duration.sum"),
                    "dram_read": d.get("dram__bytes_read.sum"), "dram_write": d.get("dram__bytes_write.sum")})
tot = sum(x["dram_read"] + x["dram_write"] for x in per)
rec = {"kernel": "disc_kernel", "volume": [n, n, n], "launches": len(per), "dram_bytes_per_launch": tot / max(1, len(per)),
       "build": R, "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none, tools/profile_round.sh "
       "(the first launches of one Tb.Th and one Tb.Sp run: chunks of 128 planes)", "per_launch": per}
json.dump(rec, open(os.path.join(ROOT, "profiles", "disc_traffic.json"), "w"), indent=1)
print(json.dumps({k: v for k, v in rec.items() if k != "per_launch"}))
