"""Condense an ncu --set full report into the per-kernel numbers DESIGN.md / profiles/ quote."""
import csv
import subprocess
import sys

KEEP = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "sm__inst_executed.sum.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "lts__t_bytes.sum", "sm__cycles_elapsed.max"]


def main(rep, out=None):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = [hdr.index(k) for k in KEEP if k in hdr]
    lines = [",".join(hdr[i] for i in idx), ",".join(units[i] for i in idx)]
    for r in rows[2:]:
        lines.append(",".join('"' + r[i].replace('"', "'") + '"' if "," in r[i] else r[i] for i in idx))
    text = "\n".join(lines) + "\n"
    if out:
        open(out, "w").write(text)
    for r in rows[2:]:
        print("----", r[hdr.index("Kernel Name")][:70])
        for i in idx[3:]:
            print(f"    {hdr[i]:72s} {r[i]:>14s} {units[i]}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
