"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (count, total, share)."""
import collections
import csv
import sys


def main(path):
    hdr, agg = None, collections.defaultdict(lambda: [0, 0.0])
    for r in csv.reader(open(path)):
        if len(r) < 6:
            continue
        if r[0] == "ID":
            hdr = r
            continue
        if hdr is None:
            continue
        d = dict(zip(hdr, r))
        try:
            v = float(d["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        v *= {"us": 1e-3, "ns": 1e-6, "s": 1e3, "ms": 1.0}.get(d["Metric Unit"], 1.0)
        k = d["Kernel Name"].split("(")[0][-48:]
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    print("%-50s %6s %12s %10s %7s" % ("kernel", "n", "total_ms", "avg_ms", "share"))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-50s %6d %12.3f %10.4f %6.1f%%" % (k, v[0], v[1], v[1] / v[0], 100 * v[1] / tot))


if __name__ == "__main__":
    main(sys.argv[1])
