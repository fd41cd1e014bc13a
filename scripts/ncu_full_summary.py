"""Key counters of an `ncu --set full` report (read here with `ncu -i ... --page raw --csv`), one line per captured launch."""
import csv
import subprocess
import sys

WANT = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_%"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu_%"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_%"), ("launch__registers_per_thread", "regs"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("lts__t_sector_hit_rate.pct", "l2_hit_%")]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "").replace("t4s::", "")
        parts = []
        for key, label in WANT:
            if key in hdr:
                i = hdr.index(key)
                parts.append(f"{label}={r[i]}{units[i] if units[i] not in ('', '%') else ''}")
        print(f"| `{name}` | " + " | ".join(parts) + " |")


if __name__ == "__main__":
    for p in sys.argv[1:]:
        print(f"\n{p}")
        main(p)
