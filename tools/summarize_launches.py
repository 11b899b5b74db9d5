"""Aggregates an ncu `--metrics gpu__time_duration.sum --csv` launch list into a per-kernel table (markdown)."""
import collections
import csv
import io
import sys


def main(path, title):
    lines = open(path).read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        name = r["Kernel Name"].split("(")[0].replace("void ", "")
        agg[name][0] += 1
        agg[name][1] += float(r["Metric Value"].replace(",", "")) / 1e3
    tot = sum(v[1] for v in agg.values())
    print(f"### {title}\n")
    print(f"{len(rows)} launches of one SD-2.1-base UNet forward (B=2), serialized by ncu; total {tot:.0f} us.\n")
    print("| kernel | launches | total us | share | avg us |")
    print("|---|---:|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k[:70]}` | {v[0]} | {v[1]:.0f} | {100 * v[1] / tot:.1f} % | {v[1] / v[0]:.1f} |")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else sys.argv[1])
