"""Dump the metrics that matter for the roofline discussion from an .ncu-rep (ncu -i ... --page raw --csv)."""
import csv
import io
import subprocess
import sys

KEYS = ("gpu__time_duration.sum", "sm__cycles_active.avg", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tensor",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size",
        "launch__cluster_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu", "smsp__inst_executed_pipe_xu", "sm__pipe_fma_cycles_active", "sm__pipe_alu_cycles_active",
        "smsp__average_warp_latency_issue_stalled", "smsp__average_warps_issue_stalled", "gpc__cycles_elapsed.max",
        "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__cycles_active.avg", "sm__mem_tensor")


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("Kernel:", r[hdr.index("Kernel Name")][:140])
        for h, u, v in zip(hdr, units, r):
            if any(k in h for k in KEYS):
                print(f"  {h} [{u}] = {v}")


if __name__ == "__main__":
    main(sys.argv[1])
