"""Aggregate an ncu source page (--print-source cuda,sass --csv) per CUDA source line: samples and instructions."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file = None
agg = collections.OrderedDict()
hdr = None
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path':
        cur_file = r[1].split('/')[-1]; continue
    if len(r) >= 2 and r[0] == 'Function Name':
        continue
    if len(r) > 4 and r[0] == 'Line No':
        hdr = r; continue
    if hdr is None or len(r) < 8: continue
    if r[0] != '' and r[0].isdigit():
        key = (cur_file, int(r[0]))
        try:
            samples = int(r[4]) if r[4] not in ('-', '') else 0
            inst = int(r[7]) if r[7] not in ('-', '') else 0
        except ValueError:
            continue
        a = agg.setdefault(key, [0, 0, r[1]])
        a[0] += samples; a[1] += inst
tot_s = sum(a[0] for a in agg.values()); tot_i = sum(a[1] for a in agg.values())
print("total samples", tot_s, "total inst", tot_i)
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% smp %5.1f%% inst  %s:%d  %s" % (100.0 * a[0] / max(tot_s, 1), 100.0 * a[1] / max(tot_i, 1), f, ln, a[2][:110]))
