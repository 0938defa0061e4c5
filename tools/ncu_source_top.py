"""Top stall sites of an ncu report's SASS source page.  usage: python tools/ncu_source_top.py <rep> [n] [kernel-index]
Prints the n instructions with the most warp-stall samples, with +-2 instructions of context, so that the wait each
role (TMA / MMA / epilogue warp) sits in can be read off."""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 15
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks = out.split('"Kernel Name"')
kidx = int(sys.argv[3]) if len(sys.argv) > 3 else 1
blk = '"Kernel Name"' + blocks[kidx]
lines = blk.splitlines()
print(lines[0][:160])
rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[1:] if len(r) == len(hdr)]
tot = sum(int(r[col["# Samples"]] or 0) for r in data)
order = sorted(range(len(data)), key=lambda i: -int(data[i][col["# Samples"]] or 0))[:n]
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print(f"total samples {tot}, instructions {len(data)}")
for i in order:
    r = data[i]
    s = int(r[col["# Samples"]] or 0)
    st = sorted(((int(r[col[h]] or 0), h) for h in stall_cols), reverse=True)[:3]
    print(f"--- {100.0 * s / max(tot, 1):5.1f}%  samples {s}  " + ", ".join(f"{h}={v}" for v, h in st if v))
    for j in range(max(0, i - 2), min(len(data), i + 3)):
        mark = ">>" if j == i else "  "
        print(f"   {mark} {data[j][col['Address']][-5:]} {data[j][col['Source']][:110]}")
