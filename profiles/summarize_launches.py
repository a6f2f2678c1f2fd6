"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and,
with --grid, per (kernel, grid) totals.   python profiles/summarize_launches.py <csv> [--grid] [--top N]"""
import collections
import csv
import re
import sys


def load(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)      # -> microseconds
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        name = re.sub(r"void |msig::|at::native::", "", name)[:64]
        yield name, row["Grid Size"], v


def main():
    path = sys.argv[1]
    by_grid = "--grid" in sys.argv
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
    tot = collections.defaultdict(list)
    for name, grid, us in load(path):
        tot[(name, grid) if by_grid else name].append(us)
    total = sum(sum(v) for v in tot.values())
    print(f"total {total / 1e3:.2f} ms over {sum(len(v) for v in tot.values())} launches")
    for k, v in sorted(tot.items(), key=lambda kv: -sum(kv[1]))[:top]:
        print(f"{sum(v) / 1e3:8.2f} ms {100 * sum(v) / total:5.1f}%  n={len(v):4d}  avg={sum(v) / len(v):8.1f} us  "
              f"min={min(v):8.1f}  max={max(v):8.1f}  {k}")


if __name__ == "__main__":
    main()
