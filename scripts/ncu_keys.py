"""Print the handful of ncu metrics we steer by from a .ncu-rep (raw page):  python scripts/ncu_keys.py file.ncu-rep"""
import csv, subprocess, sys
rep = sys.argv[1]
extra = sys.argv[2:]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct", "sm__throughput.avg.pct", "lts__t_bytes.sum",
        "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum", "smsp__average_warp", "smsp__warp_issue_stalled",
        "smsp__average_warps_issue_stalled", "sm__inst_executed.sum", "smsp__inst_executed.sum", "sm__cycles_active.avg",
        "smsp__issue_active.avg.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sector_hit_rate",
        "smsp__pcsamp_warps_issue_stalled"] + extra
for r in rows[2:]:
    print("KERNEL", r[hdr.index("Kernel Name")][:80], "grid", r[hdr.index("Grid Size")] if "Grid Size" in hdr else "")
    for h, u, v in zip(hdr, units, r):
        if any(k in h for k in keys) and v not in ("", "0", "n/a") and "peak_sustained" not in h.split(".")[-1]:
            try:
                if float(v.replace(",", "")) == 0: continue
            except ValueError:
                pass
            print(f"  {h} [{u}] = {v}")
