"""Timeline of the warp-specialised sweep kernel from a DMT_WS_TRACE build: per role warp the period per tile and the two phase
durations (a->b, b->c) in cycles.  usage: ws_trace_summary.py trace.txt [first_tile n_tiles]"""
import sys, statistics
rows = [l.split() for l in open(sys.argv[1]) if l.startswith('TR')]
d = {}
for r in rows:
    d[(r[1], int(r[2]), int(r[3]))] = tuple(int(x) for x in r[4:7])
roles = sorted(set((k[0], k[1]) for k in d), key=lambda x: x[1])
t0 = min(v[0] for v in d.values())
j0 = int(sys.argv[2]) if len(sys.argv) > 2 else 100
n = int(sys.argv[3]) if len(sys.argv) > 3 else 6
for j in range(j0, j0 + n):
    print('j=%d' % j, ' '.join('%s%d[%d +%d +%d]' % (ro[0], ro[1], d[(ro[0], ro[1], j)][0] - t0, d[(ro[0], ro[1], j)][1] - d[(ro[0], ro[1], j)][0],
                                                    d[(ro[0], ro[1], j)][2] - d[(ro[0], ro[1], j)][1]) for ro in roles if (ro[0], ro[1], j) in d))
for ro in roles:
    js = [j for j in range(100, 147) if (ro[0], ro[1], j) in d and (ro[0], ro[1], j + 1) in d]
    vs = [d[(ro[0], ro[1], j)] for j in js]
    per = [d[(ro[0], ro[1], j + 1)][0] - d[(ro[0], ro[1], j)][0] for j in js]
    print(ro, 'period %.0f  a->b %.0f  b->c %.0f' % (statistics.mean(per), statistics.mean(v[1] - v[0] for v in vs), statistics.mean(v[2] - v[1] for v in vs)))
