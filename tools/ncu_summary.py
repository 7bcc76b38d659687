"""Summarise `ncu --set full` reports of the GEMM kernels: one line per captured launch with the metrics DESIGN.md cites.
Usage: python tools/ncu_summary.py report.ncu-rep [...] > profiles/<name>.md"""
import csv
import io
import subprocess
import sys

METRICS = [
    ("time_us", "gpu__time_duration.sum"),
    ("tensor_pipe_%", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
    ("dram_rd_MB", "dram__bytes_read.sum"),
    ("dram_wr_MB", "dram__bytes_write.sum"),
    ("dram_%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("l2_hit_%", "lts__t_sector_hit_rate.pct"),
    ("l2_to_sm_TB/s", "derived__lts__lts2xbar_bytes.sum.per_second"),
    ("sm_to_l2_TB/s", "l1tex__m_l1tex2xbar_write_bytes.sum.per_second"),
    ("sm_GHz", "sm__cycles_elapsed.avg.per_second"),
    ("regs", "launch__registers_per_thread"),
]
SCALE = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6,
         "Tbyte/s": 1.0, "Gbyte/s": 1e-3, "Mbyte/s": 1e-6, "Tbyte": 1.0, "Gbyte": 1e3}


def rows_of(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def main():
    print("| report | kernel <CG,BN,A_MN,B_MN,EPI,F32> | " + " | ".join(m for m, _ in METRICS) + " |")
    print("|---|---|" + "---|" * len(METRICS))
    for path in sys.argv[1:]:
        hdr, units, rows = rows_of(path)
        idx = {h: i for i, h in enumerate(hdr)}
        for r in rows:
            name = r[idx["Kernel Name"]]
            name = name[name.index("<"):name.index(">") + 1] if "<" in name else name
            cells = []
            for label, key in METRICS:
                i = idx.get(key)
                if i is None:
                    cells.append("-")
                    continue
                v, u = float(r[i]), units[i]
                if label.endswith("MB"):
                    v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
                elif label == "time_us":
                    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
                elif label.endswith("TB/s"):
                    v *= {"Tbyte/s": 1.0, "Gbyte/s": 1e-3, "Tbyte": 1.0, "Gbyte": 1e-3}.get(u, 1.0)
                elif label == "sm_GHz":
                    v *= {"Ghz": 1.0, "Mhz": 1e-3, "GHz": 1.0, "MHz": 1e-3}.get(u, 1.0)
                cells.append("%.1f" % v if abs(v) >= 10 else "%.2f" % v)
            print("| %s | `%s` | %s |" % (path.split("/")[-1].replace(".ncu-rep", ""), name, " | ".join(cells)))


if __name__ == "__main__":
    main()
