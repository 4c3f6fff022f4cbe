"""Summarise a GOPS_B200_TIMELINE dump (rollout_tc2 kernel): mean cycles between consecutive stamps, per thread role."""
import sys
from collections import defaultdict

rows = [tuple(map(int, ln.split())) for ln in open(sys.argv[1])]
for who in (0, 1):
    ev = [(i, c) for w, i, c in rows if w == who]
    if not ev:
        continue
    seg = defaultdict(list)
    for (i0, c0), (i1, c1) in zip(ev, ev[1:]):
        seg[(i0, i1)].append(c1 - c0)
    total = ev[-1][1] - ev[0][1]
    print(f"== {'owner' if who == 0 else 'helper'}: {len(ev)} stamps, {total} cycles")
    for k, v in sorted(seg.items(), key=lambda kv: -sum(kv[1])):
        print(f"  {k[0]:>3} -> {k[1]:<3} n={len(v):4d} mean={sum(v) / len(v):8.0f}  share={sum(v) / total * 100:5.1f}%")
