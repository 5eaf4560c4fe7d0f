"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
  python scripts/launch_summary.py gpurun_out/x_launches.csv "header comment" > profiles/x_launches_summary.txt"""
import csv
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[12] == "gpu__time_duration.sum"]
acc = defaultdict(lambda: [0.0, 0])
for r in rows:
    acc[r[4]][0] += float(r[14].replace(",", "")) / 1e3
    acc[r[4]][1] += 1
tot = sum(v[0] for v in acc.values())
if len(sys.argv) > 2:
    print("# " + sys.argv[2])
print(f"# total {tot:.1f} us over {len(rows)} launches")
print("#   time_us  share   n   avg_us  kernel")
for k, v in sorted(acc.items(), key=lambda kv: -kv[1][0]):
    print(f"{v[0]:10.1f} {100 * v[0] / tot:5.1f}% {v[1]:4d} {v[0] / v[1]:8.2f}  {k[:150]}")
