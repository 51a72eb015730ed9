"""Per-kernel summary of an ncu launch list (--metrics gpu__time_duration.sum --csv)."""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hdr = None
    agg = collections.OrderedDict()
    for r in rows:
        if len(r) > 5 and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            if d.get("Metric Name") == "gpu__time_duration.sum":
                k = d["Kernel Name"][:64]
                v = float(d["Metric Value"].replace(",", ""))
                v *= {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "s": 1e6}.get(d["Metric Unit"], 1.0)
                a = agg.setdefault(k, [0, 0.0, 0.0])
                a[0] += 1; a[1] += v; a[2] = max(a[2], v)
    tot = sum(a[1] for a in agg.values())
    for k, (n, t, mx) in agg.items():
        print("%-66s n=%4d total=%10.1f us avg=%9.1f max=%9.1f share=%5.1f%%" % (k, n, t, t / n, mx, 100 * t / tot))


if __name__ == "__main__":
    main(sys.argv[1])
