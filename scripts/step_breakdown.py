"""Per-kernel durations of the last bench step from an ncu launch list (gpu__time_duration.sum csv):
    python scripts/step_breakdown.py gpurun_out/launches_x.csv"""
import csv, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
rows = list(csv.DictReader(lines))
names = [r["Kernel Name"].split("(")[0] for r in rows]
vals = []
for r in rows:
    v = float(r["Metric Value"].replace(",", "")); v *= {"us": 1e3, "ms": 1e6, "s": 1e9, "ns": 1}.get(r["Metric Unit"], 1.0); vals.append(v)
idx = [i for i, n in enumerate(names) if 'fuse_kernel' in n]
a, b = idx[-2] + 1, idx[-1] + 1
tot = 0
for i in range(a, b):
    print(f"{vals[i]/1e3:9.1f} us  {names[i][:90]}"); tot += vals[i]
print("step total us", round(tot / 1e3, 1))
