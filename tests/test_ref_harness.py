"""CPU: the reference arm of bench.py.  `oracle/ref_harness.py` runs the UNMODIFIED reference from the git-ignored copy
`baseline/_ref` (made by __graft_entry__.build() where /root/reference exists).  When the copy is present, one
training step of the reference itself is checked against the oracle restatement on the same weights and inputs --
the same pinning the committed golden fixtures provide, live."""
import pytest
import torch

from oracle import ggnn_oracle as O
from oracle import ref_harness
from situation_recognition_b200.synthetic import make_batch, make_train_json


def test_reference_copy_is_never_tracked_by_git():
    import os
    import subprocess
    root = ref_harness.ROOT
    if not os.path.isdir(os.path.join(root, ".git")):
        pytest.skip("not a git checkout")
    out = subprocess.run(["git", "ls-files", "baseline"], cwd=root, capture_output=True, text=True).stdout
    assert out.strip() == ""
    ignored = subprocess.run(["git", "check-ignore", "baseline/_ref/model.py"], cwd=root, capture_output=True, text=True)
    assert ignored.returncode == 0


@pytest.mark.skipif(not ref_harness.available(), reason="no baseline/_ref copy of the reference on this machine")
def test_reference_step_matches_oracle():
    D, B = 64, 6
    st = ref_harness.ReferenceStep(make_train_json(seed=0, images_per_verb=1), D=D, seed=3)
    st.model.eval()                                    # no dropout: the oracle then needs no mask
    enc = st.encoder
    params = {k: v.detach().clone() for k, v in st.model.state_dict().items()}
    fv, fn, gt_verb, gt_nouns = make_batch(enc, B, D, seed=8)
    t, c = O.build_tables(enc.roles_per_verb, enc.verb_list, enc.role_list)
    pv, pn, gpn = st.model(fn, gt_verb)
    vl, nl, gl = st.losses(pv, gt_verb, pn, gpn, gt_nouns)
    (vl + nl).backward()
    (ovl, onl, ogl), grads, (opv, opn, ogpn) = O.train_step_grads(params, fn, fn, gt_verb, gt_nouns, t, c,
                                                                  enc.get_num_labels())
    for mine, ref in ((opv, pv), (opn, pn), (ogpn, gpn)):
        assert torch.allclose(mine, ref.detach(), rtol=1e-5, atol=1e-6)
    assert abs(float(ovl) - vl.item()) < 1e-5 and abs(float(onl) - nl.item()) < 1e-4 and abs(float(ogl) - gl.item()) < 1e-4
    for k, p in st.model.named_parameters():
        g = p.grad if p.grad is not None else torch.zeros_like(p)
        scale = max(g.abs().max().item(), 1e-12)
        assert (grads[k] - g).abs().max().item() <= 1e-4 * scale, k
