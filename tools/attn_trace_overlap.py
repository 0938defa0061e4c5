"""Overlap analysis of the attention kernel's exp2 phases on ONE SM.
Build the trace variant (`make -C vietvoice-tts_b200/csrc VARIANT=_TR DEFS=-DVV_ATTN_TRACE=11`, smid 10), run
`VVB200_LIB=.../libvvb200_TR.so python tools/prof_kernels.py attn 1` on the GPU (writes gpurun_out/attn_trace.csv), then
`python tools/attn_trace_overlap.py gpurun_out/attn_trace.csv`: per scheduler, the share of time with 0 / 1 / 2
softmax warps inside their exp2 loop."""
import collections
import csv
import sys

rows = list(csv.DictReader(open(sys.argv[1])))
R = sorted(((int(r["cta"]), int(r["warp"]), int(r["kv"]), int(r["start"]), int(r["end"])) for r in rows), key=lambda x: x[3])
launches = [[R[0]]]
for a, b in zip(R, R[1:]):
    if b[3] - a[3] > 200000:
        launches.append([])
    launches[-1].append(b)
L = launches[-1]
t0 = min(x[3] for x in L)
print(f"last launch: {max(x[4] for x in L) - t0} cycles, {len(set(x[0] for x in L))} CTAs on this SM")
for w in range(4):
    ev = [x for x in L if x[1] == w]
    pts = sorted([(x[3], 1) for x in ev] + [(x[4], -1) for x in ev])
    cur, last, dur = 0, t0, collections.Counter()
    for t, d in pts:
        dur[cur] += t - last
        last = t
        cur += d
    tot = sum(dur.values())
    gaps = []
    by_cta = collections.defaultdict(list)
    for x in ev:
        by_cta[x[0]].append(x)
    for c in by_cta.values():
        c.sort(key=lambda x: x[3])
        gaps += [b[3] - a[4] for a, b in zip(c, c[1:])]
    print(f"scheduler {w}: {len(ev)} exp2 phases, mean {sum(x[4] - x[3] for x in ev) / len(ev):.0f} cycles, mean gap between "
          f"phases of a CTA {sum(gaps) / max(len(gaps), 1):.0f}; warps in exp2: " +
          ", ".join(f"{k}: {100 * v / tot:.1f} %" for k, v in sorted(dur.items())))
