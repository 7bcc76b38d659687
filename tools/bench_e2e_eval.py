"""End-to-end eval throughput (BASELINE.json configs[4]): stock torchvision ResNet-152 backbones + the B200-native GGNN
stage + the vectorised scorer over a synthetic batch stream, with the three parts timed SEPARATELY (the backbone is a
library module and not the product; the north star asks for it to be reported on its own).

    python tools/bench_e2e_eval.py [--batch 256] [--batches 4]            # 1 GPU
    torchrun --nproc-per-node N tools/bench_e2e_eval.py --batch 256       # each rank evaluates its own shard stream
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import situation_recognition_b200 as S  # noqa: E402
from situation_recognition_b200.imsitu_scorer import imsitu_scorer  # noqa: E402
from situation_recognition_b200.synthetic import make_train_json  # noqa: E402


def timed(fn, iters):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256, help="images per GPU per step")
    ap.add_argument("--batches", type=int, default=4)
    ap.add_argument("--autocast", default="bf16", choices=["bf16", "fp32"], help="backbone precision")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    if world > 1:
        torch.distributed.init_process_group("nccl")
    dev = torch.device("cuda", torch.cuda.current_device())

    enc = S.imsitu_encoder(make_train_json(seed=0), verbose=False)
    model = S.FCGGNN(enc, 2048, backbone="resnet152", pretrained=False).to(dev).eval()
    model = model.to(memory_format=torch.channels_last)
    B = args.batch
    g = torch.Generator().manual_seed(rank)
    img = torch.randn(B, 3, 224, 224, generator=g).to(dev).contiguous(memory_format=torch.channels_last)
    gt_verb = torch.randint(0, 504, (B,), generator=g).to(dev)
    counts = torch.tensor([enc.get_role_count(int(v)) for v in gt_verb.cpu()])
    gt_nouns = torch.randint(0, 2001, (B, 3, 6), generator=g)
    gt_nouns[(torch.arange(6)[None, None, :] >= counts[:, None, None]).expand(B, 3, 6)] = 2001
    gt_nouns = gt_nouns.to(dev)
    dt = torch.bfloat16 if args.autocast == "bf16" else torch.float32

    def backbones():
        with torch.no_grad(), torch.autocast("cuda", dtype=dt, enabled=(dt != torch.float32)):
            return model.convnet_verbs(img).float(), model.convnet_nouns(img).float()

    ident = torch.nn.Identity()

    def ggnn(fv, fn):
        # the stage alone: features in, three logits tensors + the three losses out (sr.py:183-197)
        cv, cn = model.convnet_verbs, model.convnet_nouns
        model.convnet_verbs = model.convnet_nouns = ident
        try:
            with torch.no_grad():
                pv, pn, gpn = model(fv, gt_verb, img_nouns=fn)
                losses = torch.stack([model.verb_loss(pv, gt_verb), model.nouns_loss(pn, gt_nouns),
                                      model.nouns_loss(gpn, gt_nouns)])
        finally:
            model.convnet_verbs, model.convnet_nouns = cv, cn
        return pv, pn, gpn, losses

    def score(pv, pn, gpn):
        t1, t5 = imsitu_scorer(enc, 1, 3), imsitu_scorer(enc, 5, 3)
        t1.add_point_both(pv, gt_verb, pn, gt_nouns, gpn)
        t5.add_point_both(pv, gt_verb, pn, gt_nouns, gpn)
        return t1.get_average_results_both(), t5.get_average_results_both()

    fv, fn = backbones()                      # warm-up (cudnn autotune, weight packing)
    pv, pn, gpn, _ = ggnn(fv, fn)
    score(pv, pn, gpn)
    ms_backbone, (fv, fn) = timed(backbones, args.batches)
    ms_ggnn, (pv, pn, gpn, losses) = timed(lambda: ggnn(fv, fn), args.batches)
    t0 = time.time()
    for _ in range(args.batches):
        score(pv, pn, gpn)
    torch.cuda.synchronize()
    ms_score = (time.time() - t0) / args.batches * 1e3
    total = ms_backbone + ms_ggnn + ms_score
    line = {"metric": "e2e_eval_images_per_sec", "n_gpus": world, "per_gpu_batch": B, "data": "synthetic 224x224",
            "backbone": "torchvision resnet152 x2 (random init, %s, channels_last) -- stock library module" % args.autocast,
            "ms_backbone_x2": ms_backbone, "ms_ggnn_stage": ms_ggnn, "ms_scorer": ms_score,
            "images_per_sec_total": world * B / total * 1e3, "images_per_sec_ggnn_stage_only": world * B / ms_ggnn * 1e3,
            "share_backbone": ms_backbone / total}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.barrier()
        os._exit(0)


if __name__ == "__main__":
    main()
