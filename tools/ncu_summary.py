"""Turns an `ncu --set full` report into the text summary kept under profiles/ (one `metric [unit] = value` line per
raw-page metric of the first captured launch).  bench.py reads dram__bytes_{read,write}.sum from these files.
usage: python tools/ncu_summary.py report.ncu-rep profiles/r1_ncu_full_final_<kernel>.txt"""
import csv
import subprocess
import sys

KEEP = ("Kernel Name", "Block Size", "Grid Size", "dram__", "gpu__time_duration", "launch__", "sm__inst_executed_pipe", "sm__pipe_", "sm__issue_active",
        "smsp__issue_active", "smsp__inst_executed.sum", "smsp__average_warps_issue_stalled", "sm__warps_active", "sm__throughput", "lts__t_sector_hit_rate",
        "l1tex__t_sector_hit_rate", "smsp__thread_inst_executed_per_inst_executed", "sass__inst_executed_per_opcode", "TriageCompute", "sm__cycles_elapsed",
        "smsp__cycles_active", "lts__t_bytes", "l1tex__t_bytes", "smsp__warps_eligible", "sm__maximum_warps", "smsp__inst_executed_op_local")


def main():
    rep, out = sys.argv[1], sys.argv[2]
    rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    with open(out, "w") as f:
        for h, u, v in zip(hdr, units, vals):
            if any(h.startswith(k) or k in h for k in KEEP):
                f.write(f"{h} [{u}] = {v}\n")
    print("wrote", out)


if __name__ == "__main__":
    main()
