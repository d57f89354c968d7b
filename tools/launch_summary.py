"""Aggregate an ncu launch list (`--metrics gpu__time_duration.sum --csv`) of tools/profile_step.py by kernel for the
LAST complete training step (the launches between the last two adam_kernel launches; earlier launches are engine
construction, autotuning and the first step).
    python tools/launch_summary.py gpurun_out/x_launches.csv > profiles/x_launches_by_kernel.txt"""
import collections
import csv
import re
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    ci = {h: i for i, h in enumerate(rows[start])}
    L = [(r[ci["Kernel Name"]], float(r[ci["Metric Value"]].replace(",", ""))) for r in rows[start + 1:]
         if len(r) > ci["Metric Value"] and r[ci["Metric Name"]] == "gpu__time_duration.sum"]
    ends = [i for i, (n, _) in enumerate(L) if "adam_kernel" in n]
    if len(ends) >= 2:
        L = L[ends[-2] + 1: ends[-1] + 1]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for n, ns in L:
        n = re.sub(r"^void ", "", n).split("(")[0][:60]
        agg[n][0] += 1
        agg[n][1] += ns
    tot = sum(v[1] for v in agg.values())
    print("# last complete eager step (GDN_GRAPH=0), ncu gpu__time_duration per launch, aggregated by kernel")
    print("launches %d  total %.3f ms" % (len(L), tot / 1e6))
    for n, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-62s %4d  %8.3f ms  %4.1f%%" % (n, c, ns / 1e6, 100 * ns / tot))


if __name__ == "__main__":
    main(sys.argv[1])
