"""Top stall sites of one kernel in an `ncu --set full --import-source on` report (SASS view): `python scripts/ncu_source_top.py rep regex [n [skip]]` (skip = matching launches to pass over)."""
import csv
import subprocess
import sys


def main(path, regex, n=40, skip=0):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--kernel-name", f"regex:{regex}", "--launch-skip", str(skip), "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    ia, isrc, isamp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    stalls = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    body = []
    for r in rows[2:]:
        if len(r) != len(hdr) or not r[isamp].isdigit():
            if body:
                break           # the report repeats the table per view: keep the first (SASS) one
            continue
        body.append(r)
    total = sum(int(r[isamp] or 0) for r in body)
    agg = {}
    for r in body:
        for i, h in stalls:
            agg[h] = agg.get(h, 0) + int(r[i] or 0)
    print(f"total samples {total}; by reason: " + ", ".join(f"{h[6:]}={v * 100 // max(total, 1)}%" for h, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    ranked = sorted(range(len(body)), key=lambda k: -int(body[k][isamp] or 0))[:n]
    for k in sorted(ranked):
        r = body[k]
        top = sorted(((int(r[i] or 0), h[6:]) for i, h in stalls), reverse=True)[:2]
        print(f"{k:5d} {int(r[isamp]) * 100.0 / total:5.1f}% ex={r[iex]:>8} {r[isrc].strip()[:90]:90s} {top}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40, int(sys.argv[4]) if len(sys.argv) > 4 else 0)


def buckets(path, regex, step=40):
    """Samples per bucket of `step` SASS instructions, with the notable opcodes seen in the bucket."""
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--kernel-name", f"regex:{regex}"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    body = []
    for r in rows[2:]:
        if len(r) != len(hdr) or not r[isamp].isdigit():
            if body:
                break
            continue
        body.append(r)
    total = sum(int(r[isamp]) for r in body)
    keys = ("LDTM", "STTM", "UTCHMMA", "UTMALDG", "SYNCS", "MUFU", "STS", "LDS", "BAR", "STG", "UTMAREDG", "FMNMX", "EXIT", "UTCBAR")
    for b0 in range(0, len(body), step):
        chunk = body[b0:b0 + step]
        smp = sum(int(r[isamp]) for r in chunk)
        ops = {}
        for r in chunk:
            for k in keys:
                if k in r[isrc]:
                    ops[k] = ops.get(k, 0) + 1
        ex = max(int(r[iex] or 0) for r in chunk)
        print(f"{b0:5d} {smp * 100.0 / total:5.1f}% exmax={ex:>9} {ops}")
