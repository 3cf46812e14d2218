"""gpurun_out/ncu/*.csv (raw pages written by tools/ncu_hbm.sh) -> one markdown table: duration, DRAM bytes, DRAM
throughput, L2 / L1 hit rates, achieved occupancy per captured launch.  Usage: python tools/ncu_hbm_table.py > profiles/x.md"""
import csv
import glob
import os
import sys

WANT = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"), ("lts__t_bytes.sum", "L2 bytes"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"), ("launch__registers_per_thread", "regs"),
        ("launch__grid_size", "grid"), ("smsp__inst_executed_op_global_red.sum", "global REDs")]


def main(d):
    print("| capture | kernel | " + " | ".join(n for _, n in WANT) + " |")
    print("|---|---|" + "---:|" * len(WANT))
    for path in sorted(glob.glob(os.path.join(d, "*.csv"))):
        rows = list(csv.reader(open(path)))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        idx = {h: i for i, h in enumerate(hdr)}
        for r in rows[2:]:
            name = r[idx["Kernel Name"]].replace("void ", "").replace("(anonymous namespace)::", "")[:44]
            cells = []
            for m, _ in WANT:
                cells.append(("%s %s" % (r[idx[m]], units[idx[m]])).strip() if m in idx else "-")
            print("| %s | `%s` | %s |" % (os.path.basename(path)[:-4], name, " | ".join(cells)))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/ncu")
