#!/usr/bin/env python
"""Summarise an .ncu-rep (read with `ncu -i ... --page raw --csv`) into a small CSV of the metrics
the roofline discussion uses.  Usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/x.csv"""
import csv
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__waves_per_multiprocessor",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg.per_second",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_bytes.sum", "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_tensor.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_fma.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__warps_eligible.avg.per_cycle_active",
]


TENSOR = [
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__ops_path_tensor_src_tf32_dst_fp32.avg.pct_of_peak_sustained_elapsed",
    "sm__ops_path_tensor_op_utchmma_src_tf32_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "smsp__mem_tensor_reads_op_ldt.sum.pct_of_peak_sustained_elapsed",
    "smsp__mem_tensor_reads_op_utcmma_matrix_c.sum.pct_of_peak_sustained_elapsed",
    "smsp__mem_tensor_writes_op_utcmma.sum.pct_of_peak_sustained_elapsed",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "metric", "unit", "value"])
        for vals in rows[2:]:
            name = vals[hdr.index("Kernel Name")][:100]
            for h, u, v in zip(hdr, units, vals):
                if h in KEEP or (h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")) \
                        or h in TENSOR:
                    w.writerow([name, h, u, v])
    print("wrote", out)


if __name__ == "__main__":
    main()
