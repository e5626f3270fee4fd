"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, average and share."""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    ix = {h: i for i, h in enumerate(hdr)}
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for r in rows[hi + 1:]:
        if len(r) < len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = r[ix["Kernel Name"]].split("(")[0].split("::")[-1]
        v, u = float(r[ix["Metric Value"]].replace(",", "")), r[ix["Metric Unit"]]
        v = v / 1000 if u in ("ns", "nsecond") else (v * 1000 if u in ("ms", "msecond") else v)
        tot[name] += v
        cnt[name] += 1
    total = sum(tot.values())
    print(f"{'kernel':44s} {'n':>5s} {'avg us':>9s} {'share':>7s}")
    for k, v in sorted(tot.items(), key=lambda x: -x[1]):
        print(f"{k:44s} {cnt[k]:5d} {v / cnt[k]:9.2f} {100 * v / total:6.1f}%")
    print(f"{'total':44s} {sum(cnt.values()):5d} {total:9.1f} us")


if __name__ == "__main__":
    main(sys.argv[1])
