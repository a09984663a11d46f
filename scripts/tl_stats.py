"""Median / mean / outliers of the hybrid step length from RSE_TIMELINE=1 output:
    RSE_TIMELINE=1 python bench.py --steps 300 2>&1 | grep "rse timeline" | python scripts/tl_stats.py
(the first five steps are dropped as warm-up)."""
import re
import statistics
import sys

v = [float(m.group(1)) for m in re.finditer(r"step_end=([0-9.]+)", sys.stdin.read())][5:]
print(len(v), "median", statistics.median(v), "mean", sum(v) / len(v), "outliers>1.8:", sum(1 for x in v if x > 1.8),
      "max", max(v))
