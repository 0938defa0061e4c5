"""Per-kernel table from the ncu metric pass over the preprocess / decode kernels (tools/gpu_profile_stages.sh):
launches, time, DRAM bytes and achieved DRAM GB/s against the measured HBM copy peak.
usage: python tools/ncu_stage_table.py gpurun_out/r02_ncu_stages.csv [peak GB/s]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
peak = float(sys.argv[2]) if len(sys.argv) > 2 else 6542.7
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
col = {h: i for i, h in enumerate(rows[hi])}
per = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) < len(col):
        continue
    d = per.setdefault((r[col["ID"]], r[col["Kernel Name"]]), {})
    d[r[col["Metric Name"]]] = (float(r[col["Metric Value"]].replace(",", "")), r[col["Metric Unit"]])
SC = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "nsecond": 1e-3, "us": 1, "usecond": 1, "ms": 1e3,
      "msecond": 1e3}
agg = collections.OrderedDict()
for (_, k), d in per.items():
    name = k.split("(")[0].replace("void ", "")
    a = agg.setdefault(name, [0, 0.0, 0.0, 0.0])
    t = d["gpu__time_duration.sum"]
    a[0] += 1
    a[1] += t[0] * SC.get(t[1], 1)
    a[2] += d["dram__bytes_read.sum"][0] * SC.get(d["dram__bytes_read.sum"][1], 1)
    a[3] += d["dram__bytes_write.sum"][0] * SC.get(d["dram__bytes_write.sum"][1], 1)
print(f"{'kernel':34s} launches  total us  us/launch  DRAM rd MB  DRAM wr MB   GB/s  of {peak:.0f}")
tot = 0.0
for k, a in agg.items():
    gbs = (a[2] + a[3]) / a[1] / 1e3 if a[1] else 0.0
    tot += a[1]
    print(f"{k[:34]:34s} {a[0]:8d} {a[1]:9.1f} {a[1] / a[0]:10.1f} {a[2] / 1e6:11.2f} {a[3] / 1e6:11.2f} {gbs:6.0f}  {gbs / peak:5.2f}")
print(f"total {tot:.1f} us over {sum(a[0] for a in agg.values())} launches (ncu per-launch times: cold cache, serialised)")
