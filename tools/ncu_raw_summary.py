"""Key counters per launch from `ncu -i <rep> --page raw --csv` (the `--set full` capture of tools/ncu_step_window.sh).

    python tools/ncu_raw_summary.py gpurun_out/r2_step_window_raw.csv > profiles/r02_ncu_full_step_window.txt
"""
import csv
import re
import sys

WANT = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__cycles_active.avg",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "launch__grid_size", "launch__cluster_size", "launch__registers_per_thread", "launch__shared_mem_per_block",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    for n, r in enumerate(rows[2:]):
        name = re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("void ", "")
        print(f"--- launch {n}: {name}")
        for w in WANT:
            if w in col and r[col[w]] != "":
                print(f"{w} [{units[col[w]]}] = {r[col[w]]}")


if __name__ == "__main__":
    main(sys.argv[1])
