"""GPU bring-up of the whole GGNN stage: prints error metrics of every output against the golden fixtures
(reference-generated) and against the CPU oracle.  Diagnostic tool; the pass/fail versions live in tests/."""
import json
import os
import sys
import time
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import situation_recognition_b200 as S  # noqa: E402
from situation_recognition_b200.synthetic import make_batch, make_train_json  # noqa: E402
from oracle import ggnn_oracle as O  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def relmax(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def load_model_from_golden(name, enc, D, precision):
    g = dict(np.load(os.path.join(GOLDEN, name)))
    m = S.FCGGNN(enc, D, backbone=None, precision=precision)
    sd = {k[len("param."):]: torch.from_numpy(v) for k, v in g.items() if k.startswith("param.")}
    missing = m.load_state_dict(sd, strict=False)
    assert not missing.unexpected_keys, missing
    return m.cuda(), g


def report(tag, **kw):
    print("CHECK " + json.dumps({"tag": tag, **kw}), flush=True)


def golden_case(cg):
    ann = json.loads(str(np.load(os.path.join(GOLDEN, "encoder_overfitting.npz"))["annotations_json"]))
    enc = S.imsitu_encoder(ann, verbose=False)
    for prec in ("fp32", "bf16"):
        m, g = load_model_from_golden("model_overfitting_D256.npz", enc, 256, prec)
        m.eval()
        feat = torch.from_numpy(g["feat"]).cuda()
        gt_verb = torch.from_numpy(g["gt_verb"]).cuda()
        gt_nouns = torch.from_numpy(g["gt_nouns"]).cuda()
        m._engine_for(feat.device).set_cta_group(cg)
        with torch.no_grad():
            h = m.ggsnn(torch.from_numpy(g["ggsnn_in_noun"]).cuda(), mask=torch.from_numpy(g["ggsnn_mask"]).cuda())
            report("golden_ggsnn_noun", cg=cg, prec=prec, err=relmax(h, torch.from_numpy(g["ggsnn_out_noun"])))
            h = m.ggsnn(torch.from_numpy(g["ggsnn_in_verb"]).cuda(), verb=True)
            report("golden_ggsnn_verb", cg=cg, prec=prec, err=relmax(h, torch.from_numpy(g["ggsnn_out_verb"])))
            pv, pn, gpn = m(feat, gt_verb)
            for mine, key in ((pv, "pred_verb"), (pn, "pred_nouns"), (gpn, "gt_pred_nouns")):
                ref = torch.from_numpy(g[key])
                agree = (mine.argmax(-1).cpu() == ref.argmax(-1)).float().mean().item()
                report("golden_" + key, cg=cg, prec=prec, err=relmax(mine, ref), argmax_agree=agree)
            vl = m.verb_loss(pv, gt_verb).item()
            nl = m.nouns_loss(pn, gt_nouns).item()
            gl = m.nouns_loss(gpn, gt_nouns).item()
            report("golden_losses", cg=cg, prec=prec, verb=[vl, float(g["verb_loss"])], nouns=[nl, float(g["nouns_loss"])],
                   gt=[gl, float(g["gt_nouns_loss"])])
        if prec == "bf16":
            m.zero_grad()
            pv, pn, gpn = m(feat, gt_verb)
            loss = m.verb_loss(pv, gt_verb) + m.nouns_loss(pn, gt_nouns)
            loss.backward()
            torch.cuda.synchronize()
            for k, p in m.named_parameters():
                ref = torch.from_numpy(g["grad." + k])
                got = p.grad if p.grad is not None else torch.zeros_like(p)
                report("golden_grad", cg=cg, name=k, err=relmax(got, ref), ref_max=ref.abs().max().item())


def oracle_case(B, D, prec, train, cg=2):
    enc = S.imsitu_encoder(make_train_json(seed=0), verbose=False)
    params = O.init_params(enc.get_num_verbs(), enc.get_num_roles(), enc.get_num_labels(), D, seed=0)
    m = S.FCGGNN(enc, D, backbone=None, precision=prec)
    m.load_state_dict(params, strict=False)
    m = m.cuda()
    fv, fn, gt_verb, gt_nouns = make_batch(enc, B, D, seed=1234)
    t, c = O.build_tables(enc.roles_per_verb, enc.verb_list, enc.role_list)
    m._engine_for(torch.device("cuda", 0)).set_cta_group(cg)
    t0 = time.time()
    if train:
        (vl, nl, gl), grads, (pv, pn, gpn) = O.train_step_grads(params, fv, fn, gt_verb, gt_nouns, t, c,
                                                                enc.get_num_labels())
    else:
        with torch.no_grad():
            pv, pn, gpn = O.forward(params, fv, fn, gt_verb, t, c)
            vl, nl, gl = O.verb_loss(pv, gt_verb), O.nouns_loss(pn, gt_nouns, 2001), O.nouns_loss(gpn, gt_nouns, 2001)
    report("oracle_time", B=B, D=D, train=train, seconds=time.time() - t0)
    m.eval()
    ctx = torch.enable_grad() if train else torch.no_grad()
    with ctx:
        mpv, mpn, mgpn = m(fv.cuda(), gt_verb.cuda(), img_nouns=fn.cuda())
        mvl = m.verb_loss(mpv, gt_verb.cuda())
        mnl = m.nouns_loss(mpn, gt_nouns.cuda())
        mgl = m.nouns_loss(mgpn, gt_nouns.cuda())
        if train:
            (mvl + mnl).backward()
    torch.cuda.synchronize()
    # the predicted-verb path is only comparable where the argmax verbs agree
    same_verb = (mpv.argmax(-1).cpu() == pv.argmax(-1))
    report("oracle_pred_verb", B=B, prec=prec, err=relmax(mpv, pv), argmax_agree=same_verb.float().mean().item())
    report("oracle_gt_pred_nouns", B=B, prec=prec, err=relmax(mgpn, gpn),
           argmax_agree=(mgpn.argmax(-1).cpu() == gpn.argmax(-1)).float().mean().item())
    if same_verb.any():
        report("oracle_pred_nouns", B=B, prec=prec, err=relmax(mpn.cpu()[same_verb], pn[same_verb]),
               argmax_agree=(mpn.argmax(-1).cpu()[same_verb] == pn.argmax(-1)[same_verb]).float().mean().item())
    report("oracle_losses", B=B, prec=prec, verb=[mvl.item(), float(vl)], nouns=[mnl.item(), float(nl)],
           gt=[mgl.item(), float(gl)])
    if train:
        for k, p in m.named_parameters():
            got = p.grad if p.grad is not None else torch.zeros_like(p)
            report("oracle_grad", B=B, name=k, err=relmax(got, grads[k]), ref_max=grads[k].abs().max().item())


def bf16_agreement(B=256, D=2048):
    """Top-1 agreement of the bf16 CUDA path with (a) the fp32 oracle and (b) the oracle run under CPU bf16 autocast
    (what the reference's @autocast decorators would do with bf16), on random-init weights (nearly tied logits)."""
    enc = S.imsitu_encoder(make_train_json(seed=0), verbose=False)
    params = O.init_params(enc.get_num_verbs(), enc.get_num_roles(), enc.get_num_labels(), D, seed=0)
    m = S.FCGGNN(enc, D, backbone=None, precision="bf16")
    m.load_state_dict(params, strict=False)
    m = m.cuda().eval()
    fv, fn, gt_verb, gt_nouns = make_batch(enc, B, D, seed=1234)
    t, c = O.build_tables(enc.roles_per_verb, enc.verb_list, enc.role_list)
    with torch.no_grad():
        pv32 = O.predict_verb(params, fv)
        gpn32 = O.predict_nouns(params, fn, gt_verb, t, c)
        with torch.autocast("cpu", dtype=torch.bfloat16):
            pv16 = O.predict_verb(params, fv).float()
            gpn16 = O.predict_nouns(params, fn, gt_verb, t, c).float()
        mpv = m.predict_verb(fv.cuda(), B).cpu()
        mgpn = m.predict_nouns(fn.cuda(), gt_verb.cuda(), B).cpu()
    ag = lambda a, b: (a.argmax(-1) == b.argmax(-1)).float().mean().item()
    report("bf16_agreement", B=B,
           cuda_bf16_vs_fp32_oracle={"verb": ag(mpv, pv32), "nouns": ag(mgpn, gpn32),
                                     "err_verb": relmax(mpv, pv32), "err_nouns": relmax(mgpn, gpn32)},
           cuda_bf16_vs_bf16_autocast_oracle={"verb": ag(mpv, pv16), "nouns": ag(mgpn, gpn16)},
           bf16_autocast_oracle_vs_fp32_oracle={"verb": ag(pv16, pv32), "nouns": ag(gpn16, gpn32),
                                                "err_verb": relmax(pv16, pv32), "err_nouns": relmax(gpn16, gpn32)})


def timing(B, D=2048, cg=2, iters=5):
    enc = S.imsitu_encoder(make_train_json(seed=0), verbose=False)
    m = S.FCGGNN(enc, D, backbone=None, precision="bf16").cuda()
    m._engine_for(torch.device("cuda", 0)).set_cta_group(cg)
    fv, fn, gt_verb, gt_nouns = [x.cuda() for x in make_batch(enc, B, D, seed=1234)]
    m.train()

    def step():
        m.zero_grad(set_to_none=True)
        pv, pn, gpn = m(fv, gt_verb, img_nouns=fn)
        vl = m.verb_loss(pv, gt_verb)
        nl = m.nouns_loss(pn, gt_nouns)
        with torch.no_grad():
            gl = m.nouns_loss(gpn, gt_nouns)
        (vl + nl).backward()
        return vl, nl, gl

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    report("timing_train_step", B=B, cg=cg, ms=ms, img_per_s=B / ms * 1e3, tflops=B * 6.547e9 / ms / 1e9,
           frac_sustained=B * 6.547e9 / ms / 1e9 / 1396.2)
    m.eval()
    with torch.no_grad():
        for _ in range(2):
            m(fv, gt_verb, img_nouns=fn)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(iters):
            m(fv, gt_verb, img_nouns=fn)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    report("timing_forward", B=B, cg=cg, ms=ms, img_per_s=B / ms * 1e3, tflops=B * 3.155e9 / ms / 1e9)


def main():
    stages = sys.argv[1:] or ["golden2", "golden1", "oracle_fp32", "oracle_bf16", "oracle_train", "timing"]
    for st in stages:
        try:
            if st == "golden2":
                golden_case(2)
            elif st == "golden1":
                golden_case(1)
            elif st == "oracle_fp32":
                oracle_case(256, 2048, "fp32", False)
            elif st == "oracle_bf16":
                oracle_case(256, 2048, "bf16", False)
            elif st == "oracle_train":
                oracle_case(48, 2048, "bf16", True)
            elif st == "bf16_agreement":
                bf16_agreement()
            elif st == "timing":
                timing(6144)
            elif st == "timing_small":
                timing(768)
        except Exception:
            print("STAGE_FAILED", st)
            traceback.print_exc()
        sys.stdout.flush()


if __name__ == "__main__":
    main()
