"""Runs the UNMODIFIED reference (vFones/situation-recognition: model.py + utils/) on the host CPU.

TEST / MEASUREMENT INFRASTRUCTURE ONLY (see the header of ggnn_oracle.py): used by `bench.py --impl reference`, by the
`cpu_baseline` leg of `bench.py` and by `oracle/make_golden.py`-style checks.  Never imported by the product.

The reference sources are NOT part of this repository.  `__graft_entry__.build()` copies them byte for byte from
`/root/reference` into the git-ignored `baseline/_ref/` when that directory exists (the build container), and the
snapshot sent to the GPU box carries that copy, exactly like the built `libsrggnn.so`.  Harness:
`model.resnet` is replaced by `nn.Identity` before `FCGGNN` is constructed, so `model(img := features[B, 2048],
gt_verb)` runs the reference's own `forward` -> `predict_verb` / `predict_nouns` x2 -> `GGSNN.forward` code on
backbone FEATURES (the ResNet-152 backbones are stock torchvision and are timed separately); nothing else is touched.
On CPU tensors the `@autocast()` decorators (torch.cuda.amp.autocast) do not apply, so the arithmetic is fp32.
"""
import importlib
import os
import shutil
import sys
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_COPY = os.path.join(ROOT, "baseline", "_ref")
REF_SRC = "/root/reference"
_FILES = ["model.py", os.path.join("utils", "imsitu_encoder.py"), os.path.join("utils", "imsitu_scorer.py"),
          os.path.join("utils", "imsitu_loader.py"), os.path.join("utils", "utils.py"), "sr.py", "LICENSE"]


def install(src=REF_SRC, dst=REF_COPY):
    """Byte-for-byte copy of the reference's Python files into baseline/_ref (git-ignored).  No-op without `src`."""
    if not os.path.isdir(src):
        return None
    for rel in _FILES:
        s = os.path.join(src, rel)
        if not os.path.exists(s):
            continue
        d = os.path.join(dst, rel)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(s, d)
    return dst


def available(path=REF_COPY):
    return os.path.exists(os.path.join(path, "model.py")) and os.path.exists(os.path.join(path, "utils",
                                                                                           "imsitu_encoder.py"))


def load(path=REF_COPY, stub_backbones=True):
    """(reference `model` module, reference `utils.imsitu_encoder` module) imported from `path`.
    stub_backbones=False keeps the reference's own `resnet` class (end-to-end runs: the caller makes
    torchvision.models.resnet152 constructible offline)."""
    import torch
    if not available(path):
        raise FileNotFoundError("no reference copy under %s (run __graft_entry__.build() where /root/reference exists)"
                                % path)
    for name in ("model", "utils", "utils.imsitu_encoder"):
        sys.modules.pop(name, None)
    sys.path.insert(0, path)
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ref_model = importlib.import_module("model")
            ref_enc = importlib.import_module("utils.imsitu_encoder")
    finally:
        sys.path.remove(path)
    if stub_backbones:
        ref_model.resnet = lambda out_layers: torch.nn.Identity()  # the only patch: backbones -> features pass through
    return ref_model, ref_enc


class ReferenceStep:
    """The body of the reference's training loop (sr.py:63-83) on CPU, on backbone features:
         model(img, verb); verb_loss, nouns_loss, gt_nouns_loss; (verb_loss + nouns_loss).backward();
         clip_grad_norm_(model.parameters(), 1); Adamax.step()
    GradScaler / autocast are CUDA-only in the reference (sr.py:44,64) and are no-ops here."""

    def __init__(self, train_json, D=2048, seed=0, lr=0.002):
        import contextlib
        import io
        import torch
        self.torch = torch
        ref_model, ref_enc = load()
        with contextlib.redirect_stdout(io.StringIO()):
            self.encoder = ref_enc.imsitu_encoder(train_json)
        torch.manual_seed(seed)
        self.model = ref_model.FCGGNN(self.encoder, D)
        self.model.train()
        self.opt = torch.optim.Adamax(filter(lambda p: p.requires_grad, self.model.parameters()), lr=lr)   # sr.py:472-473

    def losses(self, pred_verb, gt_verb, pred_nouns, gt_pred_nouns, gt_nouns):
        m = self.model
        return m.verb_loss(pred_verb, gt_verb), m.nouns_loss(pred_nouns, gt_nouns), m.nouns_loss(gt_pred_nouns, gt_nouns)

    def __call__(self, feat, gt_verb, gt_nouns):
        torch = self.torch
        self.opt.zero_grad()                                                       # sr.py:62
        pred_verb, pred_nouns, pred_gt_nouns = self.model(feat, gt_verb)           # sr.py:65
        verb_loss, nouns_loss, gt_nouns_loss = self.losses(pred_verb, gt_verb, pred_nouns, pred_gt_nouns, gt_nouns)
        loss = verb_loss + nouns_loss                                              # sr.py:76
        loss.backward()                                                            # sr.py:79 (no scaler on CPU)
        torch.nn.utils.clip_grad_norm_(self.model.parameters(), 1)                 # sr.py:81
        self.opt.step()                                                            # sr.py:82
        return verb_loss.item(), nouns_loss.item(), gt_nouns_loss.item()
