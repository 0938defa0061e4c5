"""One-paragraph summary per kernel of an .ncu-rep (read on the CPU box).  usage: python tools/ncu_summary.py <rep> [...]
Prints the metrics DESIGN.md / profiles/README.md quote: duration, DRAM bytes (-> roofline `traffic`), tensor-pipe and
XU (MUFU) utilisation, L2 hit rate, registers, achieved occupancy."""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_pct"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu_pipe_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("launch__registers_per_thread", "regs"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct"),
    ("sm__cycles_elapsed.avg.per_second", "sm_clock"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
]


def main():
    for rep in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        if len(rows) < 3:
            print(f"{rep}: no kernels")
            continue
        hdr, units = rows[0], rows[1]
        col = {h: i for i, h in enumerate(hdr)}
        print(f"== {rep}")
        for r in rows[2:]:
            name = r[col["Kernel Name"]]
            parts = []
            for k, short in KEYS:
                if k in col:
                    parts.append(f"{short}={r[col[k]]}{units[col[k]] if units[col[k]] not in ('', '%') else ('%' if units[col[k]] == '%' else '')}")
            print(f"  {name[:90]}\n    " + "  ".join(parts))


if __name__ == "__main__":
    main()
