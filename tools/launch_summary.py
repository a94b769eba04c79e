"""Summarise an ncu --csv launch list (gpu__time_duration.sum) by kernel: count, total, share, average."""
import csv, re, sys
from collections import defaultdict
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
H = rows[hdr]
kn, mv, mn = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Name")
tot = defaultdict(float); cnt = defaultdict(int)
last = int(sys.argv[2]) if len(sys.argv) > 2 else 0
data = [r for r in rows[hdr + 1:] if len(r) > mv and r[mn] == "gpu__time_duration.sum"]
if last: data = data[-last:]
for r in data:
    name = re.sub(r"\(.*", "", r[kn]).replace("lb::", "")
    name = re.sub(r"^void ", "", name)
    tot[name] += float(r[mv].replace(",", "")) / 1e3; cnt[name] += 1
T = sum(tot.values())
print("%d launches, total GPU time %.3f ms (cold cache, serialised: compare SHARES)" % (sum(cnt.values()), T / 1e3))
for k in sorted(tot, key=lambda k: -tot[k]):
    print("%-44s n=%5d %9.3f ms %5.1f%% avg %8.1f us" % (k[:44], cnt[k], tot[k] / 1e3, 100 * tot[k] / T, tot[k] / cnt[k]))
