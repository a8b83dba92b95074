"""Summarise an .ncu-rep (raw page) into a few key metrics per kernel.  usage: ncu_summary.py file.ncu-rep"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "lts__t_bytes.sum", "l1tex__t_bytes.sum", "smsp__inst_executed.sum", "launch__occupancy_limit_registers",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active"]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]
    ki = h.index("Kernel Name")
    for r in rows[2:]:
        print("==", r[ki][:100])
        for k in KEYS:
            if k in h:
                i = h.index(k)
                print(f"   {k:75s} {r[i]:>16s} {units[i]}")


if __name__ == "__main__":
    main(sys.argv[1])
