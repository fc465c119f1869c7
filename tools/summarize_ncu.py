"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals and shares."""
import csv
import re
import sys
from collections import defaultdict


def main(path, out=None):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    r = csv.DictReader(lines)
    for row in r:
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = row["Kernel Name"]
        name = re.sub(r"\(.*", "", name)
        val = float(row["Metric Value"].replace(",", ""))
        unit = row.get("Metric Unit", "ns")
        if unit in ("us", "usecond"):
            val *= 1e3
        elif unit in ("ms", "msecond"):
            val *= 1e6
        rows.append((name, val))
    tot = sum(v for _, v in rows)
    agg = defaultdict(lambda: [0, 0.0])
    for n, v in rows:
        agg[n][0] += 1
        agg[n][1] += v
    lines = [f"launches: {len(rows)}   total device time: {tot / 1e6:.3f} ms (cold-cache, serialised: compare shares)", "",
             "| kernel | launches | total ms | avg us | share |", "|---|---:|---:|---:|---:|"]
    for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"| `{n}` | {c} | {v / 1e6:.3f} | {v / c / 1e3:.1f} | {100 * v / tot:.1f} % |")
    txt = "\n".join(lines)
    print(txt)
    if out:
        with open(out, "w") as f:
            f.write(txt + "\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
