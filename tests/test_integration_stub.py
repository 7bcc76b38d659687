"""-m gpu: INTEGRATION.md section 2 shows the raw `ctypes` stub a maintainer of the reference would write against
include/srggnn.h.  This test executes that code block VERBATIM (so the document cannot drift from the ABI) and checks
its `predict_nouns` against the packaged host side."""
import os
import re

import numpy
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def stub_source():
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```python\n(.*?)```", text, flags=re.S)
    stubs = [b for b in blocks if "ctypes.CDLL" in b]
    assert len(stubs) == 1
    return stubs[0]


def test_stub_is_present_and_names_only_exported_symbols():
    from situation_recognition_b200 import _lib
    used = set(re.findall(r"lib\.(srg_[a-z0-9_]+)", stub_source()))
    assert used and used <= set(_lib.SIGNATURES)


@pytest.mark.gpu
def test_stub_runs_verbatim_and_matches_the_package(monkeypatch):
    import situation_recognition_b200 as S
    from situation_recognition_b200.synthetic import make_batch, make_train_json
    monkeypatch.chdir(ROOT)                                   # the stub loads the library by its repo-relative path
    encoder = S.imsitu_encoder(make_train_json(seed=0), verbose=False)
    torch.manual_seed(0)
    model = S.FCGGNN(encoder, 2048, backbone=None, precision="bf16").cuda().eval()
    ns = {"encoder": encoder, "model": model, "numpy": numpy}
    exec(compile(stub_source(), "INTEGRATION.md", "exec"), ns)
    _, fn, gt_verb, _ = make_batch(encoder, 37, 2048, seed=4)
    fn, gt_verb = fn.cuda(), gt_verb.cuda()
    got = ns["predict_nouns"](fn, gt_verb)
    with torch.no_grad():
        want = model.predict_nouns(fn, gt_verb, 37)
    torch.cuda.synchronize()
    assert got.shape == want.shape == (37, 6, 2001)
    assert torch.equal(got, want)                             # the same kernels on the same operands
