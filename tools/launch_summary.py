"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, mean and share."""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path, errors="replace")))
    hdr, agg, order = None, collections.defaultdict(list), []
    for r in rows:
        if len(r) > 5 and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            try:
                v = float(d["Metric Value"].replace(",", ""))
            except ValueError:
                continue
            if d.get("Metric Unit", "ns") in ("us", "usecond"):
                v *= 1e3
            name = d["Kernel Name"].split("(")[0][-60:]
            agg[name].append(v)
    tot = sum(sum(v) for v in agg.values())
    print(f"| kernel | launches | mean us | total us | share |\n|---|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"| {k} | {len(v)} | {sum(v) / len(v) / 1e3:.2f} | {sum(v) / 1e3:.1f} | {100 * sum(v) / tot:.1f}% |")


if __name__ == "__main__":
    main(sys.argv[1])
