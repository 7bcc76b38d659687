"""-m gpu: the raw tcgen05 GEMM behind `srg_gemm_bf16` (include/srggnn.h) against torch.matmul on every operand
layout the GGNN stage uses: K-major and MN-major A / B (forward, dgrad, wgrad), 1- and 2-CTA tiles, M tails,
bf16 and fp32 outputs, bias + alpha, split-K with the TMA reduce-add epilogue.  The cases are tools/gemm_check.py's
(the bring-up tool of round 1), now part of the suite."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gemm_check  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("idx", range(len(gemm_check.CASES)), ids=[c[0] for c in gemm_check.CASES])
def test_raw_gemm_case(idx):
    out = gemm_check.check_case(idx, time_it=False)
    assert out["ok"], out


def test_raw_gemm_rejects_bad_arguments():
    """Errors are return codes + srg_last_error(), never a launch."""
    import ctypes
    import torch
    from situation_recognition_b200 import _lib
    lib = _lib.load()
    a = torch.zeros(128, 64, device="cuda", dtype=torch.bfloat16)
    c = torch.zeros(128, 100, device="cuda")
    s = _lib.stream_ptr()
    rc = lib.srg_gemm_bf16(_lib.ptr(a), 64, 0, _lib.ptr(a), 64, 0, _lib.ptr(c), 100, _lib.SRG_DT_F32, 128, 100, 64, None,
                           1.0, 2, 1, 0, s)                       # N not a multiple of the column tile
    assert rc != 0 and b"multiple" in lib.srg_last_error()
    rc = lib.srg_gemm_bf16(_lib.ptr(a), 64, 0, _lib.ptr(a), 64, 0, _lib.ptr(c), 128, 7, 128, 128, 64, None, 1.0, 1, 1, 0, s)
    assert rc != 0 and b"c_dtype" in lib.srg_last_error()
