"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel name
launches, total time, share.  python tools/launch_summary.py file.csv [skip_first_n]"""
import collections
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    hdr, data = None, []
    for r in rows:
        if r and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            data.append(dict(zip(hdr, r)))
    data = data[skip:]
    agg = collections.defaultdict(lambda: [0, 0.0, collections.Counter()])
    for d in data:
        n = d["Kernel Name"]
        t = float(d["Metric Value"].replace(",", ""))
        if d["Metric Unit"] in ("us", "usecond"):
            t *= 1e3
        elif d["Metric Unit"] in ("ms", "msecond"):
            t *= 1e6
        agg[n][0] += 1
        agg[n][1] += t
        agg[n][2][d["Grid Size"]] += 1
    tot = sum(v[1] for v in agg.values())
    print(f"{len(data)} launches, {tot / 1e6:.3f} ms total")
    for n, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:22]:
        print(f"{n[:70]:70s} {v[0]:5d} {v[1] / 1e3:10.1f} us {100 * v[1] / tot:5.1f}%  "
              f"avg {v[1] / v[0] / 1e3:7.1f} us")


if __name__ == "__main__":
    main()
