"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): per-kernel launches, total time, share."""
import csv, sys, collections, re
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    if r[hdr.index("Metric Name")] != "gpu__time_duration.sum":
        continue
    v = float(r[iv].replace(",", ""))
    u = r[iu]
    us = v / 1000.0 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1000.0)
    name = re.sub(r"\(.*", "", r[ik]).replace("void wg::", "")
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1; a[1] += us
tot = sum(a[1] for a in agg.values())
print("| kernel | launches | total us | share | avg us |\n|---|---|---|---|---|")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("| %s | %d | %.1f | %.1f%% | %.2f |" % (k, a[0], a[1], 100 * a[1] / tot, a[1] / a[0]))
print("| total | %d | %.1f | | |" % (sum(a[0] for a in agg.values()), tot))
