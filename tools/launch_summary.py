"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launch counts, total time and
share of the LAST training step in the capture (the launches after the last `k_gather_mask` pair that opens a step).

    python tools/launch_summary.py gpurun_out/launches.csv [--all] [--out-csv profiles/x.csv]
"""
import csv
import re
import sys
from collections import OrderedDict


def short(name):
    name = re.sub(r"\(.*", "", name)
    m = re.search(r"gemm_kernel<([^>]*)>", name)
    if m:
        args = [a.strip().split(")")[-1] for a in m.group(1).split(",")]
        return "gemm<" + ",".join(args) + ">"
    m = re.search(r"(k_[a-z0-9_]+)", name)
    if m:
        return m.group(1)
    name = name.replace("void ", "").replace("at::native::", "torch:")
    return "torch:" + re.sub(r"[^A-Za-z_:]", "", name)[:44] if not name.startswith("torch:") else name[:50]


def read(path):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1.0)
        rows.append((int(r["ID"]), short(r["Kernel Name"]), v))
    return rows


def main():
    path = sys.argv[1]
    rows = read(path)
    if "--all" not in sys.argv:
        packs = [i for i, (_, k, _) in enumerate(rows) if k == "k_pack_weight_multi"]
        # a step opens with the weight repack (two pack launches a few kernels apart)
        starts = [i for n, i in enumerate(packs) if n == 0 or i - packs[n - 1] > 8]
        for n, i in enumerate(starts):     # a step's zero-fills precede its first weight-pack launch
            while i > 0 and rows[i - 1][1].startswith("torch:") and "Fill" in rows[i - 1][1]:
                i -= 1
            starts[n] = i
        if len(starts) >= 2:
            # keep the last COMPLETE step (a capture cut by `-c N` ends inside one)
            full = starts[-1] - starts[-2]
            if len(rows) - starts[-1] >= full:
                rows = rows[starts[-1]:starts[-1] + full]
            else:
                rows = rows[starts[-2]:starts[-1]]
    if "--out-csv" in sys.argv:
        with open(sys.argv[sys.argv.index("--out-csv") + 1], "w") as f:
            f.write("id,kernel,duration_ns\n")
            for i, k, v in rows:
                f.write("%d,%s,%d\n" % (i, k, v))
    agg = OrderedDict()
    for _, k, v in rows:
        n, t = agg.get(k, (0, 0.0))
        agg[k] = (n + 1, t + v)
    total = sum(t for _, t in agg.values())
    print("total %.1f us over %d launches" % (total / 1e3, len(rows)))
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-48s %4d %10.1f us %5.1f%%" % (k[:48], n, t / 1e3, 100 * t / total))


if __name__ == "__main__":
    main()
