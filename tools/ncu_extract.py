"""Selected raw metrics of every kernel in an `ncu --set full` report, as a text table (what
profiles/*_ncu_full.txt hold).   python tools/ncu_extract.py gpurun_out/x.ncu-rep > profiles/x.txt"""
import csv
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram rd"),
    ("dram__bytes_write.sum", "dram wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__cluster_size", "cluster"),
    ("smsp__inst_executed.sum", "warp instr"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_sb"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_sb"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math"),
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    cols = [(hdr.index(m), label) for m, label in WANT if m in hdr]
    for r in rows[2:]:
        name = r[ki].replace("hia::<unnamed>::", "").replace("void ", "")
        print(name[:70])
        for i, label in cols:
            print(f"    {label:16s} {r[i]:>16s} {units[i]}")
        print()


if __name__ == "__main__":
    main(sys.argv[1])
