"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list of tools/profile_step.py
into per-kernel shares for the prefill step and the decode step."""
import collections
import csv
import re
import sys


def us(row):
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    return v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v * 1e6 if u == "s" else v


def main(path, detail=False):
    with open(path) as f:
        rows = list(csv.DictReader([l for l in f if not l.startswith("==")]))
    idx = [i for i, r in enumerate(rows) if "embedding_kernel" in r["Kernel Name"]]
    if len(idx) >= 2:
        regions = {"prefill step (16 x 512 = 8192 tokens)": rows[idx[0]:idx[1]],
                   "decode step (B=256, ctx~516)": rows[idx[1]:]}
    else:
        regions = {"decode step (B=256, ctx~516)": rows[idx[0]:]}
    for title, rs in regions.items():
        agg, tot = collections.OrderedDict(), 0.0
        for r in rs:
            n = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("b200::", "")
            a = agg.setdefault(n, [0, 0.0])
            a[0] += 1; a[1] += us(r); tot += us(r)
        print(f"## {title}: {len(rs)} launches, {tot / 1e3:.3f} ms of kernel time")
        print("| kernel | launches | total ms | share | avg us |\n|---|---|---|---|---|")
        for n, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
            print(f"| {n} | {c} | {v / 1e3:.3f} | {100 * v / tot:.1f}% | {v / c:.1f} |")
        if detail:
            print("\nfirst layer, launch by launch:")
            for r in rs[:14]:
                print(f"  {us(r):8.1f} us  grid {r['Grid Size']:>14}  {re.sub(r'[(].*', '', r['Kernel Name'])[:70]}")
        print()


if __name__ == "__main__":
    main(sys.argv[1], detail=len(sys.argv) > 2)
