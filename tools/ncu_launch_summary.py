"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: total time per kernel name.
usage: ncu_launch_summary.py launches.csv [skip_first_n]"""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hdr = None; recs = []
for r in rows:
    if "Kernel Name" in r: hdr = r; continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        try: recs.append((d["Kernel Name"], float(d["Metric Value"].replace(",", "")), d.get("Metric Unit", "")))
        except ValueError: pass
recs = recs[skip:]
tot = collections.OrderedDict()
for k, v, u in recs:
    k = re.sub(r"\(.*", "", k).replace("dcrf::<unnamed>::", "").replace("void ", "")
    v = v / 1e3 if u in ("ns", "nsecond") else v   # -> us
    a = tot.setdefault(k, [0.0, 0]); a[0] += v; a[1] += 1
s = sum(a[0] for a in tot.values())
print("launches %d, total %.1f us" % (len(recs), s))
for k, (v, n) in sorted(tot.items(), key=lambda kv: -kv[1][0]):
    print("%10.1f us %5d x %9.1f  %5.1f%%  %s" % (v, n, v / n, 100 * v / s, k))
