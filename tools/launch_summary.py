"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total, share."""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    idx = {n: i for i, n in enumerate(hdr)}
    agg, tot, n = collections.OrderedDict(), 0.0, 0
    for r in rows[hi + 1:]:
        if len(r) < len(hdr):
            continue
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("fpb::<unnamed>::", "").replace("<unnamed>::", "")
        v = float(r[idx["Metric Value"]].replace(",", ""))
        unit = r[idx["Metric Unit"]]
        v = v / 1e3 if unit == "ns" else (v * 1e3 if unit == "ms" else v)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
        n += 1
    print(f"{n} launches, {tot:.1f} us total (cold-cache, serialised: compare shares)")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:36s} n={c:4d} total={t:9.1f} us {t / tot * 100:5.1f}%  avg={t / c:8.1f} us")


if __name__ == "__main__":
    main(sys.argv[1])
