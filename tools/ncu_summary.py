"""Dumps the key metrics of every launch in an .ncu-rep (or a --csv raw page) to markdown.
Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/xxx.md"""
import csv
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
]


def main(path):
    if path.endswith(".csv"):
        text = open(path).read()
    else:
        text = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(text.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print("| # | kernel | " + " | ".join(n for _, n in WANT) + " |")
    print("|---|---|" + "---:|" * len(WANT))
    for k, r in enumerate(rows[2:]):
        name = r[idx["Kernel Name"]].replace("void ", "").replace("<unnamed>::", "")[:40]
        cells = []
        for m, _ in WANT:
            if m in idx:
                cells.append("%s %s" % (r[idx[m]], units[idx[m]]))
            else:
                cells.append("-")
        print("| %d | `%s` | %s |" % (k, name, " | ".join(cells)))


if __name__ == "__main__":
    main(sys.argv[1])
