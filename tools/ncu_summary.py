"""Summarise an .ncu-rep (ncu --set full) into a small text file for profiles/.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_xxx.txt"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_warps", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__cycles_elapsed.max",
]
STALL = "smsp__average_warps_issue_stalled_"


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write(f"# summary of {rep} (ncu --set full --clock-control none); one block per profiled launch\n")
        for vals in rows[2:]:
            d = dict(zip(hdr, zip(units, vals)))
            f.write(f"\n## {d.get('Kernel Name', ('', '?'))[1]}  grid {d.get('Grid Size', ('', ''))[1]} block {d.get('Block Size', ('', ''))[1]}\n")
            for k in KEYS:
                if k in d:
                    f.write(f"{k:75s} {d[k][1]:>16s} {d[k][0]}\n")
            stalls = sorted(((float(v[1] or 0), k[len(STALL):].replace('_per_issue_active.ratio', ''))
                             for k, v in d.items() if k.startswith(STALL) and k.endswith("per_issue_active.ratio")),
                            reverse=True)
            f.write("stall reasons (warps stalled per issue-active cycle): " +
                    ", ".join(f"{n}={x:.2f}" for x, n in stalls[:8]) + "\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
