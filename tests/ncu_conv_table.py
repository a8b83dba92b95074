"""Markdown / JSON table of an `ncu --set full ... --page raw --csv` export, one row per launch.
usage: ncu_conv_table.py raw.csv out.md out.json "title" "command"
"""
import csv
import json
import sys

COLS = [("gpu__time_duration.sum", "us", 1e-3), ("dram__bytes_read.sum", "DRAM read MB", None), ("dram__bytes_write.sum", "DRAM write MB", None),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %", 1), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %", 1),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %", 1),
        ("sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "tensor pipe (hmma) %", 1)]
TO_MB = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}
TO_US = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}


def main(path, out_md, out_json, title, command):
    rows = list(csv.reader(open(path, errors="replace")))
    h, units = rows[0], rows[1]

    def find(name):
        for i, c in enumerate(h):
            if c == name or c.endswith("." + name):
                return i
        return None

    ki, gi = h.index("Kernel Name"), h.index("Grid Size")
    idx = [(find(n), lab) for n, lab, _ in COLS]
    recs = []
    for r in rows[2:]:
        if len(r) < len(h):
            continue
        name = r[ki].split("(")[0].replace("void ", "").replace("shineon::", "")
        rec = {"kernel": name, "grid": r[gi]}
        for (i, lab) in idx:
            if i is None or r[i] == "":
                continue
            v = float(r[i].replace(",", ""))
            u = units[i]
            if lab == "us":
                v *= TO_US.get(u, 1.0)
            elif "MB" in lab:
                v *= TO_MB.get(u, 1e-6)
            rec[lab] = v
        recs.append(rec)
    labs = [lab for i, lab in idx if i is not None]
    with open(out_md, "w") as f:
        f.write(f"# {title}\n\nCommand: `{command}`.\nPer-launch values; ncu serialises and replays, so durations are cold-cache.\n\n")
        f.write("| # | kernel | grid | " + " | ".join(labs) + " |\n|---:|---|---:|" + "---:|" * len(labs) + "\n")
        for n, rec in enumerate(recs):
            f.write(f"| {n} | `{rec['kernel']}` | {rec['grid']} | " + " | ".join(f"{rec.get(l, float('nan')):.1f}" for l in labs) + " |\n")
        tot = sum(r.get("us", 0) for r in recs)
        conv = [r for r in recs if "conv_igemm" in r["kernel"]]
        traffic = sum(r.get("DRAM read MB", 0) + r.get("DRAM write MB", 0) for r in conv)
        f.write(f"\nTotals: {len(recs)} launches, {tot:.0f} us under ncu; conv_igemm: {len(conv)} launches, "
                f"{traffic:.0f} MB DRAM traffic = {traffic / max(1, len(conv)):.1f} MB per launch (read + write).\n")
    json.dump({"launches": recs, "conv_launches_captured": len(conv), "conv_launches_per_step": len(conv),
               "conv_dram_bytes_per_launch": traffic * 1e6 / max(1, len(conv)), "conv_dram_bytes_per_step": traffic * 1e6,
               "command": command}, open(out_json, "w"), indent=1)
    print(open(out_md).read())


if __name__ == "__main__":
    main(*sys.argv[1:6])
