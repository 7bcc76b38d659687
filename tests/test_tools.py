"""The profiling helpers under tools/ keep working on the artefacts committed under profiles/."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    return subprocess.run([sys.executable, *args], cwd=ROOT, capture_output=True, text=True, timeout=120)


def test_bench_cmp_reads_committed_bench_lines():
    files = sorted(f for f in os.listdir(os.path.join(ROOT, "profiles")) if f.startswith("r01_bench_n1_final"))
    assert files
    res = _run("tools/bench_cmp.py", *[os.path.join("profiles", f) for f in files[-2:]])
    assert res.returncode == 0, res.stderr
    assert "gemm_gru_zr_ab" in res.stdout and "value" in res.stdout


def test_committed_bench_line_has_the_contract_keys():
    path = os.path.join(ROOT, "profiles", "r01_bench_n1_final5.json")
    with open(path) as f:
        line = json.loads([l for l in f if l.startswith("{")][-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert key in line, key
    assert line["config"]["workload"].startswith("ggnn_stage_fwd_bwd") and line["gpu_launches"] > 0
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(line["roofline"])
    assert {"value", "unit", "cores", "kind", "sample"} <= set(line["cpu_baseline"])
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(line["e2e"])
