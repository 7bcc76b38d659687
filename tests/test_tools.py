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


def test_round2_bench_lines_have_the_contract_keys():
    """The bench lines committed for round 2 (1 / 2 / 4 / 8 GPUs + the reference arm)."""
    prof = os.path.join(ROOT, "profiles")
    for n in (1, 2, 4, 8):
        with open(os.path.join(prof, "r02_bench_n%d_final.json" % n)) as f:
            line = json.loads([l for l in f if l.startswith("{")][-1])
        for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                    "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline"):
            assert key in line, (n, key)
        assert line["n_gpus"] == n and line["scaling"] == "strong" and line["gpu_launches"] > 100
        assert line["config"]["preheat_s"] >= 2.5                         # timed in the sustained clock state
        assert line["config"]["launch"] == ("eager" if n == 1 else "cuda_graph_replay")
        assert {"bound", "achieved", "peak", "unit", "frac", "traffic", "frac_burst", "step_frac",
                "step_frac_executed"} <= set(line["roofline"])
        assert 0.5 < line["roofline"]["step_frac_executed"] < 1.0           # executed FLOPs never exceed the peak
        assert not ({"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(line["clocks"]["reasons"]))
        if n == 1:
            assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["value"] > 0
            assert line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["d2h_bytes_per_step"] > 0
    with open(os.path.join(prof, "r02_bench_reference_arm_final.json")) as f:
        ref = json.loads([l for l in f if l.startswith("{")][-1])
    assert ref["impl"] == "reference" and ref["cpu_baseline"]["kind"] == "reference" and ref["gpu_launches"] == 0
    assert ref["config"]["sample_batch"] == 64 and ref["metric"] == "ggnn_images_per_sec_fwd_bwd"


def test_launch_summary_reads_the_committed_launch_list():
    res = _run("tools/launch_summary.py", os.path.join("profiles", "r02_ncu_launches_final.csv"))
    assert res.returncode == 0, res.stderr
    assert "gemm<2,256,0,0,2,0>" in res.stdout and "k_aggregate_rows" in res.stdout and "k_cast_pad" not in res.stdout
