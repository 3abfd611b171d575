"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel: python tools/launch_summary.py csv [out.md]"""
import collections
import csv
import re
import sys


def main():
    with open(sys.argv[1]) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(u, 1)
        name = row["Kernel Name"]
        m = re.match(r"(?:void )?([\w:]+(?:<[^>]*>)?)", name)
        key = m.group(1) if m else name[:60]
        agg[key][0] += 1
        agg[key][1] += ns
        tot += ns
    out = [f"total {tot / 1e6:.2f} ms over {sum(a[0] for a in agg.values())} launches (per-launch times are cold-cache, "
           "serialised; compare SHARES)", "", "| kernel | launches | total ms | share | avg us |", "|---|---|---|---|---|"]
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        out.append(f"| {k} | {n} | {t / 1e6:.2f} | {100 * t / tot:.1f}% | {t / n / 1e3:.1f} |")
    print("\n".join(out))
    if len(sys.argv) > 2:
        with open(sys.argv[2], "w") as f:
            f.write("\n".join(out) + "\n")


if __name__ == "__main__":
    main()
