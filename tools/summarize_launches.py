"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals and shares."""
import collections
import csv
import sys


def main(path, steps):
    lines = [l for l in open(path) if not l.startswith("==")]
    per = collections.OrderedDict()
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e3 if unit in ("ns", "nsecond") else (v * 1e3 if unit in ("ms", "msecond") else v)
        per.setdefault(row["Kernel Name"], []).append(v)
    total = sum(sum(v) for v in per.values())
    print(f"| kernel | launches | total us | avg us | share |")
    print(f"|---|---:|---:|---:|---:|")
    for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
        name = k.split("(")[0].replace("void ", "")[:70]
        print(f"| `{name}` | {len(v)} | {sum(v):.1f} | {sum(v) / len(v):.1f} | {100 * sum(v) / total:.1f}% |")
    print(f"\ntotal {total:.1f} us over the captured launches ({steps} steps incl. warm-up)")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "?")
