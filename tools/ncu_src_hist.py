#!/usr/bin/env python
"""Stall-sample histogram per opcode class from an `ncu --page source --csv` export of one kernel.
usage: ncu_src_hist.py file.csv"""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
hdr = rows[hi]; ix = {k: i for i, k in enumerate(hdr)}
seen = set(); data = []
for r in rows[hi + 1:]:
    if len(r) == len(hdr) and r[0] not in seen:
        seen.add(r[0]); data.append(r)
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
stalls = [k for k in hdr if k.startswith('stall_') and 'Not' not in k]
tot = sum(f(r, '# Samples') for r in data)
def cls(src):
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
    op = m.group(2) if m else '?'
    if op.startswith('IMAD.WIDE'):
        return 'IMAD.WIDE.X' if '.X' in op else ('IMAD.WIDE+cc' if re.search(r"IMAD\.WIDE\.U32 R\d+, P\d", src) else 'IMAD.WIDE')
    if op.startswith('IMAD'): return 'IMAD.other'
    return op.split('.')[0]
agg = collections.defaultdict(lambda: collections.Counter())
cnt = collections.Counter(); execd = collections.Counter()
for r in data:
    c = cls(r[1]); cnt[c] += 1; execd[c] += f(r, 'Instructions Executed')
    agg[c]['samples'] += f(r, '# Samples')
    for s in stalls: agg[c][s] += f(r, s)
print(f"total samples {tot:.0f}; instructions {len(data)}")
print(f"{'class':14s} {'static':>6s} {'exec%':>6s} {'samp%':>6s}  top stalls")
te = sum(execd.values())
for c, a in sorted(agg.items(), key=lambda x: -x[1]['samples'])[:14]:
    top = sorted(((s, v) for s, v in a.items() if s != 'samples'), key=lambda x: -x[1])[:4]
    print(f"{c:14s} {cnt[c]:6d} {100*execd[c]/te:6.1f} {100*a['samples']/tot:6.1f}  " + ", ".join(f"{s[6:]}={100*v/tot:.1f}" for s, v in top))
