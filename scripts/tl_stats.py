import sys,re
v=[float(m.group(1)) for m in re.finditer(r"step_end=([0-9.]+)", sys.stdin.read())]
v=v[5:]
import statistics
print(len(v), "median", statistics.median(v), "mean", sum(v)/len(v), "outliers>1.8:", sum(1 for x in v if x>1.8), "max", max(v))
