"""GPU bring-up check of the raw tcgen05 GEMM (srg_gemm_bf16) against torch.matmul.

Each case runs in its own subprocess with a timeout so a faulting variant cannot take the others down.
Usage:  python tools/gemm_check.py            # all cases
        python tools/gemm_check.py --case N   # one case (internal)
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "situation_recognition_b200", "libsrggnn.so")

# (name, cg, a_mn, b_mn, c_dtype, M, N, K, bias, alpha, k_splits, reduce)
CASES = [
    ("cg1_kk_small", 1, 0, 0, 1, 128, 128, 64, 0, 1.0, 1, 0),
    ("cg1_kk_k256", 1, 0, 0, 1, 128, 128, 256, 0, 1.0, 1, 0),
    ("cg1_kk_multi", 1, 0, 0, 1, 1536, 2048, 2048, 1, 0.5, 1, 0),
    ("cg1_kk_tail", 1, 0, 0, 1, 6, 256, 128, 1, 1.0, 1, 0),
    ("cg1_kk_bf16", 1, 0, 0, 2, 300, 256, 512, 1, 1.0, 1, 0),
    ("cg2_kk_small", 2, 0, 0, 1, 256, 256, 64, 0, 1.0, 1, 0),
    ("cg2_kk_multi", 2, 0, 0, 1, 1536, 2048, 2048, 1, 0.5, 1, 0),
    ("cg2_kk_tail", 2, 0, 0, 1, 6, 512, 128, 1, 1.0, 1, 0),
    ("cg2_kk_bf16", 2, 0, 0, 2, 300, 512, 512, 1, 1.0, 1, 0),
    ("cg1_kn", 1, 0, 1, 1, 384, 256, 512, 0, 1.0, 1, 0),
    ("cg2_kn", 2, 0, 1, 1, 640, 512, 512, 0, 1.0, 1, 0),
    ("cg1_nn", 1, 1, 1, 1, 256, 256, 512, 0, 1.0, 1, 0),
    ("cg2_nn", 2, 1, 1, 1, 512, 512, 1024, 0, 1.0, 1, 0),
    ("cg2_nn_splitk", 2, 1, 1, 1, 2048, 2048, 4608, 0, 1.0, 8, 1),
    ("cg1_nn_splitk", 1, 1, 1, 1, 256, 256, 4608, 0, 1.0, 4, 1),
    ("cg2_kk_big", 2, 0, 0, 2, 36864, 2048, 2048, 1, 1.0, 1, 0),
    ("cg2_kk_big4k", 2, 0, 0, 2, 36864, 4096, 4096, 1, 1.0, 1, 0),
    ("cg1_kk_big", 1, 0, 0, 2, 36864, 2048, 2048, 1, 1.0, 1, 0),
    ("cg2_kn_big", 2, 0, 1, 2, 36864, 2048, 2048, 0, 1.0, 1, 0),
    ("cg2_nn_big", 2, 1, 1, 1, 2048, 2048, 36864, 0, 1.0, 8, 1),
]


def check_case(idx, time_it=True):
    """Run CASES[idx] on the current CUDA device and return {"case", "max_abs_err", "ref_max", "ok"[, "ms", "tflops"]}.
    tests/test_gemm_cases.py runs every case through this in the -m gpu suite."""
    import torch
    name, cg, a_mn, b_mn, c_dt, M, N, K, use_bias, alpha, k_splits, reduce = CASES[idx]
    lib = ctypes.CDLL(LIB)
    lib.srg_last_error.restype = ctypes.c_char_p
    fn = lib.srg_gemm_bf16
    vp, i64, i, f = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_float
    fn.argtypes = [vp, i64, i, vp, i64, i, vp, i64, i, i, i, i, vp, f, i, i, i, vp]
    fn.restype = i
    torch.manual_seed(idx)
    dev = "cuda"
    A = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    B = (torch.randn(N, K, device=dev) * 0.5).bfloat16()
    bias = torch.randn(N, device=dev) if use_bias else None
    A_st = A.t().contiguous() if a_mn else A
    B_st = B.t().contiguous() if b_mn else B
    cdtype = torch.float32 if c_dt == 1 else torch.bfloat16
    C = torch.full((M, N), 7.0, device=dev, dtype=cdtype) if not reduce else torch.ones(M, N, device=dev)
    ref = alpha * (A.float() @ B.float().t())
    if use_bias:
        ref = ref + bias
    if reduce:
        ref = ref + 1.0
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def call():
        rc = fn(A_st.data_ptr(), A_st.stride(0), a_mn, B_st.data_ptr(), B_st.stride(0), b_mn, C.data_ptr(),
                C.stride(0), c_dt, M, N, K, bias.data_ptr() if use_bias else None, alpha, cg, k_splits, reduce,
                stream)
        if rc != 0:
            raise RuntimeError(lib.srg_last_error().decode())

    call()
    torch.cuda.synchronize()
    err = (C.float() - ref).abs().max().item()
    scale = ref.abs().max().item()
    tol = (2e-2 if c_dt == 2 else 2e-3) * max(scale, 1.0)
    out = {"case": name, "max_abs_err": err, "ref_max": scale, "ok": bool(err <= tol)}
    if not time_it:
        return out
    if M * N * K >= 10 ** 10 and not reduce:
        for _ in range(3):
            call()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 10
        e0.record()
        for _ in range(iters):
            call()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        out["ms"] = ms
        out["tflops"] = 2.0 * M * N * K / ms / 1e9
    elif M * N * K >= 10 ** 10:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            call()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        out["ms"] = ms
        out["tflops"] = 2.0 * M * N * K / ms / 1e9
    return out


def run_case(idx):
    out = check_case(idx)
    print("RESULT " + json.dumps(out), flush=True)
    return 0 if out["ok"] else 1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", type=int, default=-1)
    ap.add_argument("--only", type=str, default="")
    args = ap.parse_args()
    if args.case >= 0:
        sys.exit(run_case(args.case))
    summary = []
    for idx, c in enumerate(CASES):
        if args.only and args.only not in c[0]:
            continue
        t0 = time.time()
        try:
            res = subprocess.run([sys.executable, os.path.abspath(__file__), "--case", str(idx)], capture_output=True,
                                 text=True, timeout=180)
            lines = [l for l in res.stdout.splitlines() if l.startswith("RESULT ")]
            if lines:
                r = json.loads(lines[-1][7:])
            else:
                r = {"case": c[0], "ok": False, "error": (res.stdout + res.stderr)[-1500:]}
            r["rc"] = res.returncode
        except subprocess.TimeoutExpired:
            r = {"case": c[0], "ok": False, "error": "timeout"}
        r["wall_s"] = round(time.time() - t0, 1)
        print(json.dumps(r), flush=True)
        summary.append(r)
    bad = [r["case"] for r in summary if not r.get("ok")]
    print("SUMMARY ok=%d bad=%d %s" % (len(summary) - len(bad), len(bad), bad))


if __name__ == "__main__":
    main()
