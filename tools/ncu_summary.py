"""Summarise an .ncu-rep: per-kernel duration, DRAM bytes/throughput, occupancy, issue stats.  Usage: ncu_summary.py rep [out.md]"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "launch__grid_size", "launch__block_size",
        "smsp__cycles_active.avg", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
        "smsp__warp_issue_stalled_wait_per_warp_active.pct", "smsp__warp_issue_stalled_not_selected_per_warp_active.pct"]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        out.append(f"### {d['Kernel Name'][:90]}  (id {d['ID']})")
        for k in KEYS:
            if k in d:
                out.append(f"- {k}: {d[k]} {u[k]}")
        try:
            t = float(d["gpu__time_duration.sum"].replace(",", ""))
            tu = u["gpu__time_duration.sum"]
            t_s = t * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1}.get(tu, 1e-9)
            def b(k):
                v = float(d[k].replace(",", ""))
                return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u[k], 1)
            tr = b("dram__bytes_read.sum") + b("dram__bytes_write.sum")
            out.append(f"- DRAM traffic: {tr / 1e6:.1f} MB  -> {tr / t_s / 1e9:.0f} GB/s over {t_s * 1e6:.1f} us")
        except Exception as e:
            out.append(f"- (traffic calc failed: {e})")
        out.append("")
    txt = "\n".join(out)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(txt)
    print(txt)


main()
