"""Summarise an .ncu-rep -- or the `ncu -i rep --page raw --csv` dump of one, made on the GPU box when the report itself is too large to
bring back -- into a small CSV for profiles/ (no GPU needed).
Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep|prof_raw.csv profiles/r01_prof.csv"""
import csv
import io
import subprocess
import sys

METRICS = [
    "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
    "sm__cycles_elapsed.avg", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
    "smsp__inst_executed.sum", "smsp__cycles_active.avg", "sm__inst_executed_pipe_uniform.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
]


def main(rep, out):
    if rep.endswith(".csv"):
        with open(rep) as f:
            raw = "".join(l for l in f if not l.startswith("=="))
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = [hdr.index(m) for m in METRICS if m in hdr]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([hdr[i] for i in idx])
        w.writerow([units[i] for i in idx])
        for r in rows[2:]:
            w.writerow([r[i] for i in idx])
    print(f"{out}: {len(rows) - 2} launches, {len(idx)} metrics")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
