"""Print the headline numbers and the per-kernel table of bench.py JSON lines side by side.

    python tools/bench_cmp.py gpurun_out/a.json gpurun_out/b.json ...
"""
import json
import sys


def load(path):
    with open(path) as f:
        lines = [l for l in f if l.startswith("{")]
    return json.loads(lines[-1])


def main():
    runs = [(p, load(p)) for p in sys.argv[1:]]
    for p, d in runs:
        r = d.get("roofline", {})
        print("%-40s ms %.3f (host issue %.3f) value %.0f e2e %.0f step_frac %.3f gemm_share %.3f launches %s clocks %s" % (
            p.split("/")[-1], d["ms_per_step"], d.get("host_issue_ms_per_step", float("nan")), d["value"],
            d["e2e"]["value"], r.get("step_frac", 0), r.get("gemm_share_of_step", 0), d.get("gpu_launches"),
            (d.get("clocks") or {}).get("sm_mhz")))
    names = []
    for _, d in runs:
        for k in d.get("roofline", {}).get("kernels", []):
            if k["kernel"] not in names:
                names.append(k["kernel"])
    for n in names:
        row = "%-24s" % n
        for _, d in runs:
            k = [x for x in d["roofline"]["kernels"] if x["kernel"] == n]
            row += "  %7.3f ms %6.0f TF" % (k[0]["ms_per_step"], k[0]["tflops"]) if k else "  %20s" % "-"
        print(row)


if __name__ == "__main__":
    main()
