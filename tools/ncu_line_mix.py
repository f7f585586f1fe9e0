"""Executed warp instructions and stall samples per CUDA source line, from an ncu capture taken with --import-source on:

    ncu -i gpurun_out/prof_full.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:<kernel> > /tmp/k.csv
    python tools/ncu_line_mix.py /tmp/k.csv [top]

Aggregates over every launch of the kernels the regex selects.  Needs no GPU."""
import collections, csv, sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 24
cur, hdr, line = None, None, None
agg = collections.defaultdict(lambda: [0.0, 0.0, ""])
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr, line = r, None
        continue
    if hdr is None or len(r) < 10:
        continue
    if r[0]:                                       # a source line's own row: remember it, count only the SASS rows under it
        try:
            line = int(r[0])
        except ValueError:
            line = None
        if line is not None and r[1] and not agg[(cur, line)][2]:
            agg[(cur, line)][2] = " ".join(r[1].split())[:100]
        continue
    if line is None:
        continue
    try:
        ex, sm = float(r[hdr.index("Instructions Executed")]), float(r[hdr.index("# Samples")])
    except ValueError:
        continue
    a = agg[(cur, line)]
    a[0] += ex
    a[1] += sm
    if r[1] and not a[2]:
        a[2] = " ".join(r[1].split())[:100]
tot, ts = sum(v[0] for v in agg.values()), sum(v[1] for v in agg.values())
print("| file:line | exec % | smp % | source |\n|---|---|---|---|")
for (f, l), v in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    print("| %s:%d | %.1f | %.1f | `%s` |" % (f, l, 100 * v[0] / tot, 100 * v[1] / ts, v[2].replace("|", "\\|").replace("`", "'")))
