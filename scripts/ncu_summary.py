"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel share table (markdown)."""
import csv
import re
import sys
from collections import defaultdict


def main(path, title=""):
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        name = name.replace("t4s::", "").replace("void ", "")
        rows.append((name, ms))
    agg = defaultdict(lambda: [0, 0.0])
    for n, ms in rows:
        agg[n][0] += 1
        agg[n][1] += ms
    tot = sum(v[1] for v in agg.values())
    print(title)
    print("| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
    for n, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if ms / tot < 0.002:
            continue
        print(f"| `{n}` | {c} | {ms:.3f} | {100 * ms / tot:.1f}% |")
    print(f"\nTotal {tot:.1f} ms over {len(rows)} launches.")


if __name__ == "__main__":
    main(sys.argv[1], " ".join(sys.argv[2:]))
