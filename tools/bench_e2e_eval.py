"""End-to-end eval throughput (BASELINE.json configs[4]): a synthetic dev-set-sized image stream through the stock
torchvision ResNet-152 backbones, the B200-native GGNN stage and the vectorised scorer -- the loop of the reference's
`eval` (sr.py:165-232) -- on N GPUs (one process per GPU), next to the reference's own CPU path on the host's cores.

    python tools/bench_e2e_eval.py [--images 25200] [--batch 256] [--epochs 3]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_e2e_eval.py

The reference evaluates the dev set after EVERY training epoch (sr.py:123) and recomputes three ResNet-152 passes per
image each time although both backbones are frozen (model.py:17-18,116,159,175-178).  Here epoch 1 runs the backbones
(twice per image: the noun backbone is evaluated once for both noun passes) and stores their features in
`features.FeatureCache`; later epochs read the cache and run only the GGNN stage + scorer.  Reported per epoch and in
parts: the backbone is a library module, not the product, and the north star asks for it to be timed separately.

Images: the imSitu files are not available offline, so the stream is synthetic: every rank cycles a small pool of
pinned host batches of 224x224 images (the H2D copy is inside the timed region), each stream position with its own
image name (= cache key), verb and labels.  Weights are random-init.
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def cpu_reference(sample, steps):
    """The unmodified reference's eval step (sr.py:182-197) with its own ResNet-152 backbones on the host CPU."""
    os.environ["CUDA_VISIBLE_DEVICES"] = ""
    import torch
    import torchvision as tv
    from oracle import ref_harness
    from situation_recognition_b200.synthetic import make_batch, make_train_json
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    if not ref_harness.available():
        print(json.dumps({"cpu_reference": None, "why": "no baseline/_ref copy of the reference on this box"}))
        return
    orig = tv.models.resnet152
    tv.models.resnet152 = lambda pretrained=True, progress=False: orig(weights=None)     # no network: random init
    ref_model, ref_enc = ref_harness.load(stub_backbones=False)   # the reference's own resnet wrapper
    sys.path.insert(0, ref_harness.REF_COPY)
    from utils import imsitu_scorer as ref_scorer
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        enc = ref_enc.imsitu_encoder(make_train_json(seed=0))
    torch.manual_seed(0)
    model = ref_model.FCGGNN(enc, 2048).eval()
    _, _, gt_verb, gt_nouns = make_batch(enc, sample, 2048, seed=3)
    img = torch.randn(sample, 3, 224, 224)
    times = []
    for _ in range(steps + 1):
        t0 = time.time()
        with torch.no_grad():
            top1, top5 = ref_scorer.imsitu_scorer(enc, 1, 3), ref_scorer.imsitu_scorer(enc, 5, 3)
            pv, pn, gpn = model(img, gt_verb)                                       # sr.py:183
            top1.add_point_both(pv, gt_verb, pn, gt_nouns, gpn)                     # sr.py:185-188
            top5.add_point_both(pv, gt_verb, pn, gt_nouns, gpn)
            model.verb_loss(pv, gt_verb), model.nouns_loss(pn, gt_nouns), model.nouns_loss(gpn, gt_nouns)
        times.append(time.time() - t0)
    dt = sum(times[1:]) / steps
    print(json.dumps({"cpu_reference": {"images_per_sec": sample / dt, "cores": torch.get_num_threads(),
                                        "kind": "reference",
                                        "sample": "%d eval steps of %d images after 1 warm-up: unmodified reference "
                                                  "(model + scorer + losses, sr.py:182-197), torchvision ResNet-152 x3 "
                                                  "per image, fp32, random init" % (steps, sample)}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=25200, help="images of the stream (imSitu dev set: 25 200)")
    ap.add_argument("--batch", type=int, default=256, help="images per GPU per step")
    ap.add_argument("--epochs", type=int, default=3)
    ap.add_argument("--autocast", default="bf16", choices=["bf16", "fp32"], help="backbone precision")
    ap.add_argument("--pool", type=int, default=3, help="distinct pinned host batches cycled as the image stream")
    ap.add_argument("--cpu-sample", type=int, default=16)
    ap.add_argument("--cpu-steps", type=int, default=2)
    ap.add_argument("--no-cpu-reference", action="store_true")
    ap.add_argument("--cpu-reference-only", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.cpu_reference_only:
        return cpu_reference(args.cpu_sample, args.cpu_steps)

    import torch
    import torch.distributed as dist
    import situation_recognition_b200 as S
    from situation_recognition_b200.features import FeatureCache
    from situation_recognition_b200.imsitu_scorer import imsitu_scorer
    from situation_recognition_b200.parallel import shard_range
    from situation_recognition_b200.synthetic import make_train_json

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    cpu_line = {}
    if rank == 0 and not args.no_cpu_reference:
        # the CPU leg runs FIRST, in a child process that sees no GPU, while the other ranks wait in the rendezvous
        # (later they would spin in NCCL and compete for the host cores)
        env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
        for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "OMP_NUM_THREADS"):
            env.pop(k, None)
        try:
            res = subprocess.run([sys.executable, os.path.abspath(__file__), "--cpu-reference-only", "--cpu-sample",
                                  str(args.cpu_sample), "--cpu-steps", str(args.cpu_steps)], env=env, capture_output=True,
                                 text=True, timeout=480)
            cpu_line = json.loads([l for l in res.stdout.splitlines() if l.startswith("{")][-1])
        except Exception as ex:
            cpu_line = {"cpu_reference": None, "why": str(ex)[:200]}
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    dev = torch.device("cuda", torch.cuda.current_device())

    enc = S.imsitu_encoder(make_train_json(seed=0), verbose=False)
    torch.manual_seed(0)                                      # same random-init weights on every rank
    model = S.FCGGNN(enc, 2048, backbone="resnet152", pretrained=False).to(dev).eval()
    model = model.to(memory_format=torch.channels_last)
    B = args.batch
    lo, hi = shard_range(args.images, rank, world)            # this rank's slice of the stream
    n_local = hi - lo
    steps = (n_local + B - 1) // B
    g = torch.Generator().manual_seed(1000 + rank)
    pool = [torch.randn(B, 3, 224, 224, generator=g).contiguous(memory_format=torch.channels_last).pin_memory()
            for _ in range(args.pool)]
    gt_verb_all = torch.randint(0, 504, (n_local,), generator=g)
    counts = torch.tensor([enc.get_role_count(int(v)) for v in gt_verb_all])
    gt_nouns_all = torch.randint(0, 2001, (n_local, 3, 6), generator=g)
    gt_nouns_all[(torch.arange(6)[None, None, :] >= counts[:, None, None]).expand(n_local, 3, 6)] = 2001
    gt_verb_all, gt_nouns_all = gt_verb_all.to(dev), gt_nouns_all.to(dev)
    names_all = ["img_%08d" % i for i in range(lo, hi)]
    dt = torch.bfloat16 if args.autocast == "bf16" else torch.float32
    cache = FeatureCache(n_local, 2048, dev)

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def epoch():
        """One pass of the reference's eval loop (sr.py:176-201) over this rank's slice; returns timing parts (ms)."""
        top1, top5 = imsitu_scorer(enc, 1, 3), imsitu_scorer(enc, 5, 3)
        sums = torch.zeros(3, device=dev)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        part = [0.0, 0.0, 0.0]
        for s in range(steps):
            i0, i1 = s * B, min(n_local, (s + 1) * B)
            names, verb, nouns = names_all[i0:i1], gt_verb_all[i0:i1], gt_nouns_all[i0:i1]
            ev[0].record()
            with torch.no_grad():
                feats = cache.lookup(names)
                if feats is None:
                    img = pool[s % len(pool)][:i1 - i0].to(dev, non_blocking=True)           # H2D of the image batch
                    with torch.autocast("cuda", dtype=dt, enabled=(dt != torch.float32)):
                        feats = model.extract_features(img)                                  # ResNet-152 x2 (library)
                    cache.store(names, *feats)
                ev[1].record()
                pv, pn, gpn = model.forward_features(feats[0], feats[1], verb)               # the GGNN stage
                sums += torch.stack([model.verb_loss(pv, verb), model.nouns_loss(pn, nouns), model.nouns_loss(gpn, nouns)])
                ev[2].record()
                top1.add_point_both(pv, verb, pn, nouns, gpn)
                top5.add_point_both(pv, verb, pn, nouns, gpn)
                ev[3].record()
            ev[3].synchronize()
            for k in range(3):
                part[k] += ev[k].elapsed_time(ev[k + 1])
        a1, a5 = top1.get_average_results_both(), top5.get_average_results_both()
        return part, sums, (a1, a5)

    # warm-up on one batch (cudnn autotune, weight packing), then forget it
    w = pool[0].to(dev)
    with torch.no_grad(), torch.autocast("cuda", dtype=dt, enabled=(dt != torch.float32)):
        fv, fn = model.extract_features(w)
    with torch.no_grad():
        model.forward_features(fv, fn, gt_verb_all[:B] if n_local >= B else gt_verb_all.repeat(B)[:B])
    del w, fv, fn

    per_epoch = []
    for e in range(args.epochs):
        sync()
        t0 = time.perf_counter()
        part, sums, _ = epoch()
        sync()
        wall = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        parts = torch.tensor(part, device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(wall, op=dist.ReduceOp.MAX)
            dist.all_reduce(parts, op=dist.ReduceOp.MAX)
        per_epoch.append({"epoch": e, "seconds": wall.item(), "images_per_sec": args.images / wall.item(),
                          "ms_backbones_or_cache": parts[0].item(), "ms_ggnn_stage_and_losses": parts[1].item(),
                          "ms_scorer": parts[2].item(), "cache_hits": cache.hits, "cache_misses": cache.misses})

    line = {"metric": "e2e_eval_images_per_sec", "n_gpus": world, "images": args.images, "per_gpu_batch": B,
            "epochs": per_epoch, "data": "synthetic 224x224 stream from pinned host batches (H2D timed), random-init",
            "backbone": "torchvision resnet152 x2 (%s, channels_last) -- stock library module" % args.autocast,
            "first_epoch_images_per_sec": per_epoch[0]["images_per_sec"],
            "cached_epoch_images_per_sec": per_epoch[-1]["images_per_sec"] if len(per_epoch) > 1 else None,
            "feature_cache_mb_per_gpu": 2 * n_local * 2048 * 4 / 2 ** 20}
    line.update(cpu_line)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        os._exit(0)


if __name__ == "__main__":
    main()
