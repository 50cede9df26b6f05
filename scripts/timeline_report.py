"""Per-kernel breakdown of a timeline CSV written by scripts/timeline.py:  python scripts/timeline_report.py <csv> [min_kernels]"""
import collections
import csv
import re
import sys


def short(n):
    n = n.replace('void ', '').replace('(anonymous namespace)::', '')
    m = re.match(r'([\w:]+)(<[^(]*>)?', n)
    return (m.group(1) + (m.group(2) or ''))[:64] if m else n[:64]


def main():
    rows = list(csv.DictReader(open(sys.argv[1])))
    mink = int(sys.argv[2]) if len(sys.argv) > 2 else 300
    ev = sorted((float(r['start_us']), float(r['dur_us']), r['stream'], r['name']) for r in rows)
    segs, cur, end = [], [ev[0]], ev[0][0] + ev[0][1]
    for e in ev[1:]:
        if e[0] - end > 60:
            segs.append(cur)
            cur = []
        cur.append(e)
        end = max(end, e[0] + e[1])
    segs.append(cur)
    seen = set()
    for s in segs:
        if len(s) < mink or len(s) in seen:
            continue
        seen.add(len(s))
        t0, t1 = s[0][0], max(e[0] + e[1] for e in s)
        print(f"step with {len(s)} kernels: wall {t1 - t0:.1f} us, sum of kernel time {sum(e[1] for e in s):.1f} us")
        agg = collections.defaultdict(lambda: [0, 0.0])
        for e in s:
            agg[short(e[3])][0] += 1
            agg[short(e[3])][1] += e[1]
        for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:30]:
            print(f"  {n:66s} {c:4d} {t:9.1f} us  avg {t / c:7.2f}")


main()
