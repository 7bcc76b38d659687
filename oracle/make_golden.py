"""Generate tests/golden/*.npz by EXECUTING the unmodified reference (/root/reference) in the build container.

Run here (CPU, no GPU needed):   python oracle/make_golden.py
The reference cannot travel to the GPU box, so its outputs are committed as fixtures; tests/test_oracle_golden.py
pins oracle/ggnn_oracle.py and the package's encoder against them.

Harness: `model.resnet` is replaced by nn.Identity so that predict_*(img=features) runs the
untouched reference code on [B, D] features; nothing else is patched (CPU => autocast is a no-op => fp32).
"""
import io
import contextlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("SRG_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)


def load_reference():
    sys.path.insert(0, REF)
    import warnings
    warnings.filterwarnings("ignore")
    import model as ref_model                      # /root/reference/model.py
    from utils import imsitu_encoder as ref_enc    # /root/reference/utils/imsitu_encoder.py
    ref_model.resnet = lambda out_layers: torch.nn.Identity()
    return ref_model, ref_enc


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def encoder_tables(enc):
    V, R = enc.get_num_verbs(), enc.get_max_role_count()
    verbs = torch.arange(V)
    return {
        "verb2roles": enc.roles_to_verb_tensor_list.numpy().astype(np.int64),
        "role_count": np.array([enc.get_role_count(v) for v in range(V)], dtype=np.int64),
        "role_ids_batch_all": enc.get_role_ids_batch(verbs).numpy(),
        "adj_all": enc.get_adj_matrix_noself(verbs).numpy(),
        "num": np.array([V, enc.get_num_roles(), enc.get_num_labels(), R], dtype=np.int64),
    }


def run_model_case(ref_model, enc, D, B, seed, train_json):
    torch.manual_seed(seed)
    model = ref_model.FCGGNN(enc, D)
    model.eval()  # dropout = identity; parity of the dropout mask itself is tested CUDA-vs-oracle with an explicit mask
    g = torch.Generator().manual_seed(seed + 1)
    feat = torch.randn(B, D, generator=g).abs() * 0.5
    V, R, L = enc.get_num_verbs(), enc.get_max_role_count(), enc.get_num_labels()
    gt_verb = torch.randint(0, V, (B,), generator=g)
    # labels: encode real annotations when the verbs come from the json, else random with the reference's padding
    gt_nouns = torch.full((B, 3, R), L, dtype=torch.int64)
    for b in range(B):
        n = enc.get_role_count(int(gt_verb[b]))
        gt_nouns[b, :, :n] = torch.randint(0, L, (3, n), generator=g)
    params = {k: v.detach().clone() for k, v in model.state_dict().items()}
    for p in model.parameters():
        p.requires_grad_(True)
    pred_verb, pred_nouns, gt_pred_nouns = model(feat, gt_verb)
    # model.verb_loss / nouns_loss call .cuda() on the loss module (model.py:184,191), which needs a GPU;
    # the loss arithmetic itself is reproduced with the same torch calls (model.py:185,195-199).
    vl = torch.nn.CrossEntropyLoss()(pred_verb, gt_verb)
    nl_fn = torch.nn.CrossEntropyLoss(ignore_index=L)
    pn_t = pred_nouns.transpose(1, 2)
    gpn_t = gt_pred_nouns.transpose(1, 2)
    nl = sum(nl_fn(pn_t, gt_nouns[torch.arange(B), i]) for i in range(3))
    gl = sum(nl_fn(gpn_t, gt_nouns[torch.arange(B), i]) for i in range(3))
    (vl + nl).backward()
    out = {"feat": feat.numpy(), "gt_verb": gt_verb.numpy(), "gt_nouns": gt_nouns.numpy(),
           "pred_verb": pred_verb.detach().numpy(), "pred_nouns": pred_nouns.detach().numpy(),
           "gt_pred_nouns": gt_pred_nouns.detach().numpy(),
           "verb_loss": vl.detach().numpy(), "nouns_loss": nl.detach().numpy(), "gt_nouns_loss": gl.detach().numpy()}
    # GGSNN.forward alone (model.py:59-86), both modes
    with torch.no_grad():
        h0 = torch.randn(B * R, D, generator=g).abs() * 0.3
        mask = enc.get_adj_matrix_noself(gt_verb)
        out["ggsnn_in_noun"] = h0.numpy()
        out["ggsnn_mask"] = mask.numpy()
        out["ggsnn_out_noun"] = model.ggsnn(h0, mask=mask, verb=False).numpy()
        hv = torch.randn(B, D, generator=g).abs() * 0.3
        out["ggsnn_in_verb"] = hv.numpy()
        out["ggsnn_out_verb"] = model.ggsnn(hv, mask=None, verb=True).numpy()
    for k, v in params.items():
        out["param." + k] = v.numpy()
    for k, p in model.named_parameters():
        out["grad." + k] = (p.grad if p.grad is not None else torch.zeros_like(p)).numpy()
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    ref_model, ref_enc = load_reference()
    from situation_recognition_b200.synthetic import make_train_json

    # (1) the reference's own fixture imSitu/overfitting.json: 5 verbs / 7 roles / 30 labels / max 4 roles
    with open(os.path.join(REF, "imSitu", "overfitting.json")) as f:
        over = json.load(f)
    enc_over = quiet(ref_enc.imsitu_encoder, over)
    tabs = encoder_tables(enc_over)
    # label encoding of every fixture image (imsitu_encoder.py:161-166,182-207)
    enc_items = [enc_over.encode(over[k]) for k in over]
    tabs["encode_verbs"] = np.array([v for v, _ in enc_items], dtype=np.int64)
    tabs["encode_labels"] = torch.stack([l for _, l in enc_items]).numpy()
    tabs["annotations_json"] = np.array(json.dumps(over))   # the encoder's input, so the fixture is self-contained
    np.savez_compressed(os.path.join(OUT, "encoder_overfitting.npz"), **tabs)

    # (2) synthetic imSitu-shaped vocabulary: 504 / 190 / 2001 / 6
    train = make_train_json(seed=0)
    enc_syn = quiet(ref_enc.imsitu_encoder, train)
    tabs = encoder_tables(enc_syn)
    g = torch.Generator().manual_seed(7)
    rnd = torch.randint(0, 504, (97,), generator=g)
    tabs["rand_verbs"] = rnd.numpy()
    tabs["rand_role_ids"] = enc_syn.get_role_ids_batch(rnd).numpy()
    tabs["rand_adj"] = enc_syn.get_adj_matrix_noself(rnd).numpy()
    keys = list(train)[:40]
    items = [enc_syn.encode(train[k]) for k in keys]
    tabs["encode_verbs"] = np.array([v for v, _ in items], dtype=np.int64)
    tabs["encode_labels"] = torch.stack([l for _, l in items]).numpy()
    np.savez_compressed(os.path.join(OUT, "encoder_synthetic504.npz"), **tabs)

    # (3) model cases: forward logits, losses and autograd gradients of the unmodified FCGGNN
    case_a = run_model_case(ref_model, enc_over, D=256, B=5, seed=11, train_json=over)
    np.savez_compressed(os.path.join(OUT, "model_overfitting_D256.npz"), **case_a)
    case_b = run_model_case(ref_model, enc_syn, D=64, B=8, seed=12, train_json=train)
    # keep this fixture small: drop the big classifier gradient/param duplicates that case A already pins
    np.savez_compressed(os.path.join(OUT, "model_synthetic504_D64.npz"), **case_b)
    # (4) the reference scorer (utils/imsitu_scorer.py) on seeded random logits with planted hits
    from utils import imsitu_scorer as ref_scorer
    g = torch.Generator().manual_seed(21)
    B, V, R, L = 64, 504, 6, 200        # the scorer only ranks the last axis; 200 labels keep the fixture small
    verbs = torch.randint(0, V, (B,), generator=g)
    gt_nouns = torch.full((B, 3, R), L, dtype=torch.int64)
    for b in range(B):
        n = enc_syn.get_role_count(int(verbs[b]))
        gt_nouns[b, :, :n] = torch.randint(0, L, (3, n), generator=g)
    pred_verbs = torch.randn(B, V, generator=g)
    pred_nouns = torch.randn(B, R, L, generator=g)
    gt_pred_nouns = torch.randn(B, R, L, generator=g)
    for b in range(B):          # plant correct answers so that every metric is exercised
        if b % 3 == 0:
            pred_verbs[b, verbs[b]] += 10
        n = enc_syn.get_role_count(int(verbs[b]))
        for r in range(n):
            if (b + r) % 2 == 0 and b % 5 != 0:
                pred_nouns[b, r, gt_nouns[b, (b + r) % 3, r]] += 10
            if (b + r) % 4 != 1 and b % 7 != 0:
                gt_pred_nouns[b, r, gt_nouns[b, r % 3, r]] += 10
    sc = {"verbs": verbs.numpy(), "gt_nouns": gt_nouns.numpy(), "pred_verbs": pred_verbs.numpy(),
          "pred_nouns": pred_nouns.numpy().astype(np.float16), "gt_pred_nouns": gt_pred_nouns.numpy().astype(np.float16)}
    pn16, gpn16 = torch.from_numpy(sc["pred_nouns"]).float(), torch.from_numpy(sc["gt_pred_nouns"]).float()
    for k in (1, 5):
        s = ref_scorer.imsitu_scorer(enc_syn, k, 3)
        s.add_point_both(pred_verbs[:40], verbs[:40], pn16[:40], gt_nouns[:40], gpn16[:40])
        s.add_point_both(pred_verbs[40:], verbs[40:], pn16[40:], gt_nouns[40:], gpn16[40:])
        avg = s.get_average_results_both()
        keys = sorted(avg)
        sc["top%d_keys" % k] = np.array(keys)
        sc["top%d_avg" % k] = np.array([avg[x] for x in keys], dtype=np.float64)
        sc["top%d_cards" % k] = np.array([[float(c[x]) for x in keys] for c in s.score_cards], dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "scorer_synthetic504.npz"), **sc)
    for n in sorted(os.listdir(OUT)):
        print(n, os.path.getsize(os.path.join(OUT, n)))


if __name__ == "__main__":
    main()
