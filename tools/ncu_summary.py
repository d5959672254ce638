"""Extract the judged metrics from an .ncu-rep (read offline with `ncu -i ... --page raw --csv`)."""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "dram_active_pct"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("lts__t_bytes.sum", "l2_bytes"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__cycles_active.avg", "cycles"),
]


def main(path, limit=8):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    cols = [(hdr.index(k), n) for k, n in WANT if k in hdr]
    ki = hdr.index("Kernel Name")
    print("| kernel | " + " | ".join(n for _, n in cols) + " |")
    print("|---|" + "---|" * len(cols))
    for r in rows[2:2 + limit]:
        name = r[ki].split("(")[0].replace("void ", "").replace("b200::", "")
        print(f"| {name} | " + " | ".join(f"{r[i]} {units[i]}".strip() for i, _ in cols) + " |")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 8)
