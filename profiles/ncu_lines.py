"""Per-source-line stall samples of one kernel from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur, hdr = None, None
agg = collections.defaultdict(lambda: [0, 0, 0, 0, 0, 0])
srcs = {}
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) > 4 and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) - 5 or r[0] == "":
        continue
    try:
        ln = int(r[0])
    except ValueError:
        continue
    d = dict(zip(hdr[4:], r[4:]))

    def f(k):
        try:
            return float(d.get(k, 0) or 0)
        except ValueError:
            return 0.0
    a = agg[(cur, ln)]
    for j, k in enumerate(("# Samples", "Instructions Executed", "stall_long_sb", "stall_barrier", "stall_wait", "stall_short_sb")):
        a[j] += f(k)
    srcs[(cur, ln)] = r[1]
tot = sum(a[0] for a in agg.values())
toti = sum(a[1] for a in agg.values())
byfile = collections.defaultdict(lambda: [0, 0])
for (f_, l), a in agg.items():
    byfile[f_][0] += a[0]
    byfile[f_][1] += a[1]
print("samples %d, warp instructions %d; share of samples / instructions per file:" % (tot, toti),
      {k: (round(v[0] / tot, 3), round(v[1] / toti, 3)) for k, v in byfile.items()})
print("file:line | % samples | % instructions | long_scoreboard barrier wait short_scoreboard samples | source")
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%-15s %4d %5.1f%% %5.1f%% %6.0f %6.0f %6.0f %6.0f | %s" % (key[0], key[1], 100 * a[0] / tot, 100 * a[1] / toti,
                                                                  a[2], a[3], a[4], a[5], srcs[key].strip()[:100]))
