"""Per-kernel counts of the SASS mnemonics that prove a Blackwell-native kernel (B200_PROFILING.md): tcgen05.mma ->
UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG / UTMASTG / UTMAREDG / UTMAPF / UBLKCP, plus MUFU and the packed
fp32 pipe.  The .so is git-ignored (built artefact), so this summary is the committed evidence.

usage: python tools/sass_summary.py [vietvoice-tts_b200/libvvb200.so] > profiles/r02_sass_summary.txt"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "vietvoice-tts_b200/libvvb200.so"
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
WATCH = ["UTCHMMA", "UTCQMMA", "UTCMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTMAPF", "UBLKCP",
         "UTCBAR", "SYNCS", "MUFU.EX2", "MUFU.TANH", "FFMA2", "FADD2", "FMUL2", "HMMA", "USETMAXREG", "BAR"]
per = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        per[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        per[cur]["_total"] += 1
        for w in WATCH:
            if op.startswith(w):
                per[cur][w] += 1
                break
demangle = subprocess.run(["c++filt"], input="\n".join(per), capture_output=True, text=True).stdout.splitlines()
print(f"# SASS summary of {lib} (cuobjdump -sass; sm_100a only)")
print("# columns: instructions | " + " ".join(WATCH))
tot = collections.Counter()
for (name, c), dn in zip(per.items(), demangle):
    dn = re.sub(r"CUtensorMap_st", "TMap", dn)
    cols = " ".join(f"{w}={c[w]}" for w in WATCH if c[w])
    print(f"{dn[:110]:110s} | {c['_total']:6d} | {cols}")
    tot.update(c)
print("# TOTAL | " + " ".join(f"{w}={tot[w]}" for w in WATCH if tot[w]))
print("# legacy tensor path (HMMA / mma.sync) instructions:", tot["HMMA"])
