"""Summarises an .ncu-rep (ncu --set full) into the handful of numbers DESIGN.md / bench.py cite.
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x_summary.txt"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__inst_issued.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("kernel:", r[hdr.index("Kernel Name")][:100])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"  {k:82s} {r[i]:>20s} {units[i]}")
        for i, h in enumerate(hdr):
            if "average_warps_issue_stalled" in h and "per_issue_active" in h:
                try:
                    if float(r[i]) > 0.05:
                        print(f"  {h:82s} {r[i]:>20s}")
                except ValueError:
                    pass
        print()


if __name__ == "__main__":
    main()
