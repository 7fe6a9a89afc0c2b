"""Summarise an `ncu --metrics gpu__time_duration.sum --csv --log-file X` launch list: launches, total and share per kernel."""
import csv
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ik, im, iv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows:
    if len(r) > iv and r[im] == "gpu__time_duration.sum":
        tot[r[ik]] += float(r[iv].replace(",", "")) / 1e3
        cnt[r[ik]] += 1
total = sum(tot.values())
print("launches   total_us  share   avg_us  kernel")
for k in sorted(tot, key=tot.get, reverse=True):
    print(f"{cnt[k]:8d} {tot[k]:10.1f} {100 * tot[k] / total:5.1f}% {tot[k] / cnt[k]:8.2f}  {k[:110]}")
