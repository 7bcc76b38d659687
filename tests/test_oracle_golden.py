"""Pins oracle/ggnn_oracle.py and the package's encoder against fixtures produced by the UNMODIFIED reference
(oracle/make_golden.py, run in the build container).  CPU only."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import ggnn_oracle as O
from situation_recognition_b200.imsitu_encoder import imsitu_encoder
from situation_recognition_b200.synthetic import make_train_json

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")


def overfit_annotations():
    """Annotations of the reference's tiny fixture set, carried inside the golden file (written by make_golden.py)."""
    return json.loads(str(np.load(os.path.join(GOLDEN, "encoder_overfitting.npz"))["annotations_json"]))


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name)))


def _params(g, dtype=torch.float32):
    return {k[len("param."):]: torch.from_numpy(v).to(dtype) for k, v in g.items() if k.startswith("param.")}


@pytest.fixture(scope="module")
def enc_syn():
    return imsitu_encoder(make_train_json(seed=0), verbose=False)


@pytest.fixture(scope="module")
def enc_over():
    return imsitu_encoder(overfit_annotations(), verbose=False)


def test_hand_checkable_kat(enc_over):
    # the encoder built from overfitting.json, verified by hand
    assert enc_over.roles_to_verb_tensor_list.tolist() == [[0, 1, 2, 7], [3, 4, 0, 2], [5, 0, 2, 7], [6, 0, 2, 7],
                                                           [3, 4, 0, 2]]
    assert (enc_over.get_num_verbs(), enc_over.get_num_roles(), enc_over.get_num_labels(),
            enc_over.get_max_role_count()) == (5, 7, 30, 4)


@pytest.mark.parametrize("which", ["overfitting", "synthetic504"])
def test_encoder_tables_match_reference(which, enc_over, enc_syn):
    g = load("encoder_%s.npz" % which)
    enc = enc_over if which == "overfitting" else enc_syn
    V, nr, nl, R = g["num"].tolist()
    assert (enc.get_num_verbs(), enc.get_num_roles(), enc.get_num_labels(), enc.get_max_role_count()) == (V, nr, nl, R)
    assert np.array_equal(enc.roles_to_verb_tensor_list.numpy(), g["verb2roles"])
    assert np.array_equal(np.array([enc.get_role_count(v) for v in range(V)]), g["role_count"])
    verbs = torch.arange(V)
    ids = enc.get_role_ids_batch(verbs)
    adj = enc.get_adj_matrix_noself(verbs)
    assert ids.dtype == torch.int64 and adj.dtype == torch.float32
    assert np.array_equal(ids.numpy(), g["role_ids_batch_all"])
    assert np.array_equal(adj.numpy(), g["adj_all"])
    # oracle restatement of the same two functions
    t, c = O.build_tables(enc.roles_per_verb, enc.verb_list, enc.role_list)
    assert np.array_equal(t, g["verb2roles"]) and np.array_equal(c, g["role_count"])
    assert np.array_equal(O.get_role_ids_batch(t, np.arange(V)), g["role_ids_batch_all"])
    assert np.array_equal(O.get_adj_matrix_noself(c, np.arange(V), R), g["adj_all"])
    # verb2role_encoding (imsitu_encoder.py:93-112)
    for v in range(V):
        assert enc.verb2role_encoding[v].tolist() == [1] * int(g["role_count"][v]) + [0] * (R - int(g["role_count"][v]))


def test_encoder_random_batch_and_labels(enc_syn, enc_over):
    g = load("encoder_synthetic504.npz")
    rnd = torch.from_numpy(g["rand_verbs"])
    assert np.array_equal(enc_syn.get_role_ids_batch(rnd).numpy(), g["rand_role_ids"])
    assert np.array_equal(enc_syn.get_adj_matrix_noself(rnd).numpy(), g["rand_adj"])
    train = make_train_json(seed=0)
    keys = list(train)[:40]
    items = [enc_syn.encode(train[k]) for k in keys]
    assert [v for v, _ in items] == g["encode_verbs"].tolist()
    assert np.array_equal(torch.stack([l for _, l in items]).numpy(), g["encode_labels"])
    go = load("encoder_overfitting.npz")
    over = overfit_annotations()
    items = [enc_over.encode(over[k]) for k in over]
    assert [v for v, _ in items] == go["encode_verbs"].tolist()
    assert np.array_equal(torch.stack([l for _, l in items]).numpy(), go["encode_labels"])


def test_empty_batch(enc_syn):
    assert enc_syn.get_role_ids_batch(torch.zeros(0, dtype=torch.int64)).shape == (0, 6)
    assert enc_syn.get_adj_matrix_noself(torch.zeros(0, dtype=torch.int64)).shape == (0, 6, 6)


@pytest.mark.parametrize("name,enc_name", [("model_overfitting_D256.npz", "over"), ("model_synthetic504_D64.npz", "syn")])
def test_oracle_forward_losses_grads_match_reference(name, enc_name, enc_over, enc_syn):
    g = load(name)
    enc = enc_over if enc_name == "over" else enc_syn
    params = _params(g)
    t, c = O.build_tables(enc.roles_per_verb, enc.verb_list, enc.role_list)
    feat = torch.from_numpy(g["feat"])
    gt_verb = torch.from_numpy(g["gt_verb"])
    gt_nouns = torch.from_numpy(g["gt_nouns"])
    L = enc.get_num_labels()
    (vl, nl, gl), grads, (pv, pn, gpn) = O.train_step_grads(params, feat, feat, gt_verb, gt_nouns, t, c, L)
    # same torch ops in the same order => bit-exact or within a few ulp
    for mine, key in [(pv, "pred_verb"), (pn, "pred_nouns"), (gpn, "gt_pred_nouns")]:
        ref = torch.from_numpy(g[key])
        assert mine.shape == ref.shape
        assert torch.allclose(mine, ref, rtol=1e-5, atol=1e-6), key
        assert torch.equal(mine.argmax(-1), ref.argmax(-1)), key
    assert abs(float(vl) - float(g["verb_loss"])) < 1e-5
    assert abs(float(nl) - float(g["nouns_loss"])) < 1e-5
    assert abs(float(gl) - float(g["gt_nouns_loss"])) < 1e-5
    for k, v in grads.items():
        if "convnet" in k:
            continue
        ref = torch.from_numpy(g["grad." + k])
        scale = max(ref.abs().max().item(), 1e-12)
        assert (v - ref).abs().max().item() <= 1e-5 * scale + 1e-8, k


@pytest.mark.parametrize("name", ["model_overfitting_D256.npz", "model_synthetic504_D64.npz"])
def test_oracle_ggsnn_matches_reference(name):
    g = load(name)
    p = O._sub(_params(g), "ggsnn.")
    out = O.ggsnn_forward(p, torch.from_numpy(g["ggsnn_in_noun"]), mask=torch.from_numpy(g["ggsnn_mask"]), verb=False)
    assert torch.allclose(out, torch.from_numpy(g["ggsnn_out_noun"]), rtol=1e-5, atol=1e-6)
    out = O.ggsnn_forward(p, torch.from_numpy(g["ggsnn_in_verb"]), mask=None, verb=True)
    assert torch.allclose(out, torch.from_numpy(g["ggsnn_out_verb"]), rtol=1e-5, atol=1e-6)


def test_bias_counted_six_times_quirk():
    """model.py:73-75: W_p's bias is applied to all R neighbours (masked ones too) before the sum."""
    torch.manual_seed(0)
    D, B, R = 32, 3, 6
    p = {k: v for k, v in O._sub(O.init_params(5, 7, 11, D, seed=3), "ggsnn.").items()}
    h = torch.rand(B * R, D)
    rc = np.array([2, 6, 1])
    mask = torch.from_numpy(O.get_adj_matrix_noself(rc, np.arange(3), R))
    tr = []
    O.ggsnn_forward(p, h, mask=mask, steps=1, trace=tr)
    agg = torch.einsum("bij,bjd->bid", mask, h.view(B, R, D)).reshape(B * R, D)
    closed = agg @ p["W_p.weight"].t() + R * p["W_p.bias"]
    assert torch.allclose(tr[0]["m"], closed, rtol=1e-5, atol=1e-5)


def test_tables_from_any_reference_style_encoder(enc_syn):
    """FCGGNN only needs the reference encoder API: a duck-typed object without `device_tables` yields the same
    flat tables (this is how the reference's own imsitu_encoder instance is consumed)."""
    from situation_recognition_b200.imsitu_encoder import tables_from_encoder

    class RefStyle:
        def __init__(self, e):
            self.roles_to_verb_tensor_list = e.roles_to_verb_tensor_list
            self._e = e

        def get_num_verbs(self):
            return self._e.get_num_verbs()

        def get_max_role_count(self):
            return self._e.get_max_role_count()

        def get_role_count(self, v):
            return self._e.get_role_count(v)

    t0, c0 = tables_from_encoder(enc_syn)
    t1, c1 = tables_from_encoder(RefStyle(enc_syn))
    assert t0.dtype == np.int32 and c0.dtype == np.int32
    assert np.array_equal(t0, t1) and np.array_equal(c0, c1)
    assert t0.shape == (504 * 6,) and int(t0.max()) == 190 and c0.min() >= 1 and c0.max() == 6


def test_encoder_pickles_without_transforms(enc_syn, tmp_path):
    """sr.py caches the encoder with torch.save (sr.py:442-447)."""
    p = tmp_path / "encoder"
    torch.save(enc_syn, p)
    e2 = torch.load(p, weights_only=False)
    assert e2.verb_list == enc_syn.verb_list and torch.equal(e2.roles_to_verb_tensor_list, enc_syn.roles_to_verb_tensor_list)
    assert e2.dev_transform is not None
