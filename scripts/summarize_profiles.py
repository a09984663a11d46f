"""Turn the raw ncu outputs in gpurun_out/ into the small text summaries kept under profiles/.

    python scripts/summarize_profiles.py <round-tag> <launches.csv> [<full.ncu-rep> ...]
"""
import collections
import csv
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
           "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
           "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"]


def launches(tag, path):
    lines = [l for l in open(path) if not l.startswith("==")]
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for row in csv.DictReader(lines):
        name = row["Kernel Name"].split("(")[0]
        if "rse::" not in name:
            continue          # torch kernels that generate the synthetic corpus are not part of the step
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        v *= {"us": 1e3, "ms": 1e6, "s": 1e9}.get(row["Metric Unit"], 1.0)
        tot[name] += v; cnt[name] += 1
    T = sum(tot.values())
    out = [f"# ncu launch list summary ({tag}) — per-launch times are cold-cache and serialised: compare SHARES",
           f"# source: {Path(path).name}; {sum(cnt.values())} launches, {T/1e6:.1f} ms total", ""]
    for k, v in sorted(tot.items(), key=lambda x: -x[1])[:25]:
        out.append(f"{v/T*100:6.2f}%  n={cnt[k]:5d}  avg={v/cnt[k]/1e3:10.1f} us  {k[:110]}")
    # the kernels of ONE step (between the last two fuse_kernel launches), in launch order
    rows = [r for r in csv.DictReader(lines)]
    names = [r["Kernel Name"].split("(")[0] for r in rows]
    fuse = [i for i, n in enumerate(names) if "fuse_kernel" in n]
    if len(fuse) >= 2:
        out += ["", "# one step (launch order; ncu serialises the launches — in a real step bm25_fx_kernel runs on a second",
                "# stream underneath knn_tc3_kernel<1>, so the step is shorter than this sum):"]
        tot = 0.0
        for i in range(fuse[-2] + 1, fuse[-1] + 1):
            v = float(rows[i]["Metric Value"].replace(",", "")) * {"us": 1e3, "ms": 1e6, "s": 1e9}.get(rows[i]["Metric Unit"], 1.0)
            tot += v
            out.append(f"{v/1e3:10.1f} us  {names[i][:100]}")
        out.append(f"{tot/1e3:10.1f} us  total")
    (ROOT / "profiles" / f"{tag}_launches_summary.txt").write_text("\n".join(out) + "\n")
    print("\n".join(out))


def full(tag, rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    out = [f"# ncu --set full capture ({tag}) of {Path(rep).name}", ""]
    for li, row in enumerate(data):
        out.append(f"## launch {li}: {row[hdr.index('Kernel Name')][:120]}")
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                out.append(f"  {m:90s} {row[i]:>18s} {units[i]}")
        out.append("")
    name = Path(rep).stem
    (ROOT / "profiles" / f"{name}_metrics.txt").write_text("\n".join(out) + "\n")
    # DRAM traffic per launch of the dominant kernel (bench.py quotes it as roofline.traffic)
    import json
    for row in data:
        kn = row[hdr.index("Kernel Name")]
        if "knn_tc3_kernel<1>" in kn or "knn_tc3_kernel<(int)1>" in kn:
            def val(m):
                i = hdr.index(m)
                return float(row[i].replace(",", "")) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(units[i], 1)
            grid = row[hdr.index("Grid Size")] if "Grid Size" in hdr else ""
            (ROOT / "profiles" / f"{tag}_dominant_kernel_traffic.json").write_text(json.dumps({
                "kernel": "knn_tc3_kernel<1> (filter)", "tc_kind": "f16-shadow", "rows": 4799462, "source": Path(rep).name,
                "dram_bytes_read": val("dram__bytes_read.sum"), "dram_bytes_write": val("dram__bytes_write.sum"),
                "gpu_time_ms_under_ncu": float(row[hdr.index("gpu__time_duration.sum")].replace(",", "")) *
                {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(units[hdr.index("gpu__time_duration.sum")], float("nan")),
                "grid": grid}, indent=1) + "\n")
            break
    print("\n".join(out[:40]))


if __name__ == "__main__":
    tag = sys.argv[1]
    launches(tag, sys.argv[2])
    for rep in sys.argv[3:]:
        full(tag, rep)
