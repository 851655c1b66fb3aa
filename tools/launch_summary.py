"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (shares, not absolutes)."""
import collections
import csv
import re
import sys


def summarise(path, top=30):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg, n = collections.OrderedDict(), 0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = re.sub(r"<.*", "", row["Kernel Name"])
        k = re.sub(r"\(.*", "", k)[:64]
        v = float(row["Metric Value"].replace(",", ""))
        v = v / 1e3 if row["Metric Unit"] == "ns" else (v * 1e3 if row["Metric Unit"] == "ms" else v)
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
        n += 1
    tot = sum(a[1] for a in agg.values())
    out = [f"launches {n}, total {tot:.1f} us", "", "| share | total us | n | avg us | kernel |", "|---|---|---|---|---|"]
    for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
        out.append(f"| {t / tot * 100:.2f}% | {t:.1f} | {c} | {t / c:.1f} | `{k}` |")
    return "\n".join(out)


if __name__ == "__main__":
    print(summarise(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30))
