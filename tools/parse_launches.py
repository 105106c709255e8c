"""summarise an `ncu --csv` launch list: python tools_parse_launches.py file.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
H = rows[hdr]; data = rows[hdr + 1:]
iK, iM, iV, iID = H.index('Kernel Name'), H.index('Metric Name'), H.index('Metric Value'), H.index('ID')
per = {}
for r in data:
    if len(r) <= iV: continue
    per.setdefault(r[iID], {'k': r[iK]})[r[iM]] = r[iV]
f = lambda d, k: float(d.get(k, '0').replace(',', ''))
tot = sum(f(d, 'gpu__time_duration.sum') for d in per.values())
for id_, d in per.items():
    us = f(d, 'gpu__time_duration.sum') / 1000
    name = d['k'].replace('unnamed>::', '').replace('void ', '')[:58]
    print(f"{id_:>3} {name:58s} {us:10.1f} us {100*us*1000/tot:5.1f}%  regs {d.get('launch__registers_per_thread','-'):>3}  "
          f"rd {f(d,'dram__bytes_read.sum')/1e6:8.1f} MB wr {f(d,'dram__bytes_write.sum')/1e6:8.1f} MB  warps {d.get('sm__warps_active.avg.pct_of_peak_sustained_active','-')}")
print(f"total {tot/1e6:.3f} ms over {len(per)} launches")
