"""ncu report -> small CSV under profiles/ (kernel name + the counters the rooflines are argued from).
Usage: python scripts/ncu_extract.py gpurun_out/x.ncu-rep profiles/y.csv [kernel-substring ...]"""
import csv
import subprocess
import sys

COLS = ["Kernel Name", "gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active"]
rep, out, filt = sys.argv[1], sys.argv[2], sys.argv[3:]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = [hdr.index(c) for c in COLS if c in hdr]
with open(out, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow([hdr[i] for i in idx])
    w.writerow([units[i] for i in idx])
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        if not filt or any(s in name for s in filt):
            w.writerow([r[i] for i in idx])
print(f"{out}: {len(rows) - 2} launches in the report")
