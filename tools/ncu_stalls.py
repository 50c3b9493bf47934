"""Per-source-line stall breakdown from an ncu source page CSV (--print-source cuda,sass)."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
key = sys.argv[3] if len(sys.argv) > 3 else "stall_long_sb"
hdr = None; cur = None; agg = {}
tot = collections.Counter()
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if len(r) > 4 and r[0] == 'Line No': hdr = r; cols = {n: i for i, n in enumerate(hdr)}; continue
    if hdr is None or len(r) < len(hdr) or not r[0].isdigit(): continue
    k = (cur, int(r[0]))
    a = agg.setdefault(k, [collections.Counter(), r[1]])
    for name in ("stall_long_sb", "stall_short_sb", "stall_wait", "stall_no_inst", "stall_branch_resolving", "stall_barrier", "stall_math", "stall_selected", "stall_not_selected", "stall_lg", "stall_mio", "stall_dispatch"):
        if name in cols:
            try: v = int(r[cols[name]])
            except ValueError: v = 0
            a[0][name] += v; tot[name] += v
print("totals:", dict(tot))
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0][key])[:top]:
    print("%6d %s:%d  %s" % (a[0][key], f, ln, a[1][:100]))
