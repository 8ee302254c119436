#!/usr/bin/env python
"""Top source lines by warp-stall samples from `ncu --page source --csv --print-source cuda,sass`."""
import collections
import csv
import sys


def num(x):
    try:
        return float(x.replace(',', ''))
    except Exception:
        return 0.0


rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
sections, cur = [], None
for r in rows:
    if r and r[0] == "File Path":
        cur = {'file': r[1], 'rows': []}
        sections.append(cur)
    elif r and r[0] == "Line No" and cur is not None:
        cur['hdr'] = r
    elif cur is not None and 'hdr' in cur and len(r) == len(cur['hdr']):
        cur['rows'].append(r)
for s in sections:
    H = s['hdr']
    si, li, ie = H.index("# Samples"), H.index("Line No"), H.index("Instructions Executed")
    srcs = [i for i, h in enumerate(H) if h == "Source"]
    tot = sum(num(r[si]) for r in s['rows'])
    print("==", s['file'], "total samples", tot, "rows", len(s['rows']))
    agg = collections.defaultdict(lambda: [0, 0, ''])
    for r in s['rows']:
        k = r[li]
        agg[k][0] += num(r[si])
        agg[k][1] += num(r[ie])
        agg[k][2] = r[srcs[0]][:120]
    for k, (sm, ins, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top_n]:
        print("  line %5s samples %8.0f (%4.1f%%) inst %10.0f  %s" % (k, sm, 100 * sm / max(tot, 1), ins, src))
