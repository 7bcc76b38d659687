"""The vectorised scorer against the UNMODIFIED reference scorer's outputs (tests/golden/scorer_synthetic504.npz,
written by oracle/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from situation_recognition_b200.imsitu_encoder import imsitu_encoder
from situation_recognition_b200.imsitu_scorer import imsitu_scorer
from situation_recognition_b200.synthetic import make_train_json

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _run(device):
    g = dict(np.load(os.path.join(GOLDEN, "scorer_synthetic504.npz")))
    enc = imsitu_encoder(make_train_json(seed=0), verbose=False)
    t = lambda k: torch.from_numpy(g[k]).to(device)
    pv, v, gn = t("pred_verbs"), t("verbs"), t("gt_nouns")
    pn, gpn = t("pred_nouns").float(), t("gt_pred_nouns").float()
    for k in (1, 5):
        s = imsitu_scorer(enc, k, 3)
        s.add_point_both(pv[:40], v[:40], pn[:40], gn[:40], gpn[:40])
        s.add_point_both(pv[40:], v[40:], pn[40:], gn[40:], gpn[40:])
        avg = s.get_average_results_both()
        keys = [str(x) for x in g["top%d_keys" % k]]
        assert sorted(avg) == keys
        assert np.allclose([avg[x] for x in keys], g["top%d_avg" % k], rtol=0, atol=1e-12)
        cards = np.array([[float(c[x]) for x in keys] for c in s.score_cards])
        assert np.array_equal(cards, g["top%d_cards" % k])
        assert 0.0 < g["top%d_avg" % k].min() and g["top%d_avg" % k].max() < 1.0   # every metric is exercised


def test_scorer_matches_reference_cpu():
    _run("cpu")


@pytest.mark.gpu
def test_scorer_matches_reference_cuda():
    _run("cuda")


def test_scorer_empty_batch_and_empty_average():
    enc = imsitu_encoder(make_train_json(seed=0, images_per_verb=1), verbose=False)
    s = imsitu_scorer(enc, 1, 3)
    s.add_point_both(torch.zeros(0, 504), torch.zeros(0, dtype=torch.long), torch.zeros(0, 6, 2001),
                     torch.zeros(0, 3, 6, dtype=torch.long), torch.zeros(0, 6, 2001))
    assert s.score_cards == []
    with pytest.raises(ZeroDivisionError):
        s.get_average_results_both()
