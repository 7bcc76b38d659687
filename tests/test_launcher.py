"""sr.py-compatible launcher: CLI surface on CPU; all five modes end to end on the GPU with a tiny synthetic dataset."""
import json
import os
import re

import pytest
import torch


def test_cli_flags_match_reference():
    from situation_recognition_b200.sr import build_parser
    flags = {a.option_strings[0] for a in build_parser()._actions if a.option_strings}
    reference = {'--resume_model', '--evaluate_dev', '--evaluate_test', '--test_img', '--verb', '--subset',
                 '--model_saving_name', '--saving_folder', '--imgset_dir', '--dataset_folder', '--train_file',
                 '--dev_file', '--test_file', '--batch_size', '--num_workers', '--epochs', '--lr'}     # sr.py:384-420
    assert reference <= flags
    d = build_parser().parse_args([])
    assert (d.batch_size, d.num_workers, d.epochs, d.lr) == (6144, 10, 1000, 0.002)


def test_sharded_batch_sampler_covers_every_sample_once():
    from situation_recognition_b200.imsitu_loader import ShardedBatchSampler
    for n, gb, world in ((10, 4, 2), (7, 7, 1), (25, 8, 4), (5, 4, 4)):
        per_rank = [list(ShardedBatchSampler(n, gb, r, world, shuffle=True, seed=3)) for r in range(world)]
        assert len({len(x) for x in per_rank}) == 1                      # same number of steps on every rank
        seen = sorted(i for x in per_rank for b in x for i in b)
        assert set(seen) == set(range(n)) and len(seen) >= n


def test_format_dict_matches_reference_format():
    from situation_recognition_b200.utils import format_dict
    assert format_dict({'verb': 0.3237, 'value': 0.7468}, '{:.2f}', '1-') == '1-verb: 32.37, 1-value: 74.68'


def _make_dataset(root):
    from PIL import Image
    from situation_recognition_b200.synthetic import make_train_json
    ann = make_train_json(seed=0, images_per_verb=1)
    keys = list(ann)
    train = {k: ann[k] for k in keys}
    small = {k: ann[k] for k in keys[:6]}
    os.makedirs(os.path.join(root, "imSitu"))
    os.makedirs(os.path.join(root, "resized_256"))
    for name, d in (("train.json", train), ("tiny.json", small), ("dev.json", small), ("test.json", small)):
        with open(os.path.join(root, "imSitu", name), "w") as f:
            json.dump(d, f)
    g = torch.Generator().manual_seed(0)
    for k in small:
        arr = (torch.rand(256, 256, 3, generator=g) * 255).to(torch.uint8).numpy()
        Image.fromarray(arr).save(os.path.join(root, "resized_256", k))
    nouns = {lab: {"gloss": ["gloss_" + lab]} for fr in ann.values() for f in fr["frames"] for lab in f.values() if lab}
    verbs = {v["verb"]: {"roles": {r: {} for r in v["frames"][0]}} for v in ann.values()}
    with open(os.path.join(root, "imSitu", "imsitu_space.json"), "w") as f:
        json.dump({"nouns": nouns, "verbs": verbs}, f)
    return keys[0]


@pytest.mark.gpu
def test_all_modes_end_to_end(tmp_path, capsys, monkeypatch):
    from situation_recognition_b200 import sr
    first = _make_dataset(str(tmp_path))
    monkeypatch.chdir(tmp_path)
    common = ["--no_pretrained", "--num_workers", "0", "--batch_size", "4"]
    sr.main(common + ["--train_file", "tiny.json", "--epochs", "2", "--model_saving_name", "tiny"])
    out = capsys.readouterr().out
    assert "Model training started!" in out and "Epoch-0, lr: 0.0020" in out and "Epoch-1, lr: 0.0020" in out
    # the frozen backbones' dev-set features are cached: the second epoch's validation reports the same kind of line and
    # (same weights for the backbones, updated GGNN) runs without touching the images again
    assert len(re.findall(r"val losses = \[v:", out)) == 2
    assert re.search(r"training losses = \[v: \d+\.\d\d, n: \d+\.\d\d, gt: \d+\.\d\d\]", out)
    assert re.search(r"1-verb: \d+\.\d\d, 1-value: \d+\.\d\d, 1-value-all: \d+\.\d\d", out)
    ckpt = torch.load(tmp_path / "checkpoints" / "tiny", weights_only=False)
    assert set(ckpt) == {'epoch', 'avg_scores', 'verb_losses', 'nouns_losses', 'val_avg_scores', 'val_verb_losses',
                         'val_nouns_losses', 'model_state_dict', 'optimizer_state_dict'}            # sr.py:145-157
    assert ckpt['epoch'] == 2 and 'ggsnn.W_p.weight' in ckpt['model_state_dict']
    assert 'convnet_verbs.model.conv1.weight' in ckpt['model_state_dict']
    # the fused optimizer's state has torch.optim.Adamax's format: 20 trainable tensors, and torch's Adamax loads it
    osd = ckpt['optimizer_state_dict']
    assert len(osd['state']) == 20 and set(osd['state'][0]) == {'step', 'exp_avg', 'exp_inf'}
    assert float(osd['state'][0]['step']) == 4.0                      # 6 images / batch 4 = 2 steps per epoch
    ps = [torch.nn.Parameter(torch.zeros_like(osd['state'][i]['exp_avg'])) for i in range(20)]
    torch.optim.Adamax(ps, lr=0.002).load_state_dict(osd)
    sr.main(common + ["--resume_model", "tiny", "--evaluate_dev"])
    out = capsys.readouterr().out
    assert "Loading encoder file" in out and "=> evaluating model with dev-set..." in out and "val losses = [v:" in out
    sr.main(common + ["--resume_model", "tiny", "--test_img", os.path.join("resized_256", first)])
    out = capsys.readouterr().out
    assert "No ground truth verb found, calculating by myself..." in out and re.search(r"action \(\d+\.\d\d%\): verb\d+", out)
    sr.main(common + ["--resume_model", "tiny", "--test_img", os.path.join("resized_256", first), "--verb", "verb003"])
    assert "action (100.00%): verb003" in capsys.readouterr().out
    sr.main(common + ["--resume_model", "tiny", "--subset", "2"])
    out = capsys.readouterr().out
    assert out.count("Analizing: ") == 2 and "---- Ground truth ----" in out


@pytest.mark.gpu
def test_two_rank_launcher(tmp_path):
    """`torchrun --nproc-per-node 2 -m situation_recognition_b200.sr`: sharded loader (unequal shards: 6 images over
    global batches of 4 -> 2+2, then 1+1), global loss denominators, sharded fused optimizer, rank-0 checkpoint.  On a
    single-GPU box the two ranks share the GPU and exchange through gloo (SRG_DIST_BACKEND); with two or more GPUs the
    same command runs on NCCL.  The resulting weights must equal a single-process run of the same two epochs."""
    import subprocess
    import sys
    from situation_recognition_b200 import sr
    _make_dataset(str(tmp_path))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    args = ["--no_pretrained", "--num_workers", "0", "--batch_size", "4", "--train_file", "tiny.json", "--epochs", "2",
            "--no_feature_cache"]
    env = dict(os.environ, PYTHONPATH=root + os.pathsep + os.environ.get("PYTHONPATH", ""), SRG_SEED="7")
    if torch.cuda.device_count() < 2:
        env["SRG_DIST_BACKEND"] = "gloo"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", "-m", "situation_recognition_b200.sr"] + args + \
          ["--model_saving_name", "two"]
    res = subprocess.run(cmd, cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "Using 2 GPUs!" in res.stdout and res.stdout.count("training losses = [v:") == 2      # rank 0 prints once
    res1 = subprocess.run([sys.executable, "-m", "situation_recognition_b200.sr"] + args + ["--model_saving_name", "one"],
                          cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=900)
    assert res1.returncode == 0, res1.stdout[-3000:] + res1.stderr[-3000:]
    two = torch.load(tmp_path / "checkpoints" / "two", weights_only=False)
    one = torch.load(tmp_path / "checkpoints" / "one", weights_only=False)
    assert two["epoch"] == one["epoch"] == 2
    # same data order and initial weights (SRG_SEED + the rank-0 broadcast).  The runs agree statistically, not bit for
    # bit: the dropout masks differ (Philox is keyed by the local row) and the backbones' BatchNorm layers are in training
    # mode (sr.py:24), so a rank normalises with the statistics of its 2-image shard -- exactly as a DataParallel replica
    # does.  Exact equality of sharded and full-batch gradients is tests/test_parallel_cuda.py's job.
    import math
    for k in ("verb_losses", "nouns_losses", "val_verb_losses", "val_nouns_losses"):
        assert len(two[k]) == len(one[k]) == 2
        for a, b in zip(two[k], one[k]):
            assert math.isfinite(a) and math.isfinite(b) and 0.5 * b <= a <= 2.0 * b, (k, two[k], one[k])
    init = {k: v for k, v in one["model_state_dict"].items() if k.startswith("ggsnn.")}
    w2, w1 = two["model_state_dict"]["ggsnn.U_z.weight"], one["model_state_dict"]["ggsnn.U_z.weight"]
    assert (w2 - w1).abs().max().item() <= 2 * 4 * 0.002 + 1e-4            # 4 steps of at most lr each, both ways
    assert torch.equal(two["model_state_dict"]["convnet_nouns.model.conv1.weight"],
                       one["model_state_dict"]["convnet_nouns.model.conv1.weight"])      # frozen, broadcast from rank 0


@pytest.mark.gpu
def test_e2e_eval_stream_tool(tmp_path):
    """tools/bench_e2e_eval.py (BASELINE configs[4]) on a tiny stream: epoch 1 runs the backbones and fills the feature
    cache, epoch 2 is served from it; the JSON line carries the per-epoch parts."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, os.path.join(root, "tools", "bench_e2e_eval.py"), "--images", "48", "--batch",
                          "32", "--epochs", "2", "--pool", "1", "--no-cpu-reference"], cwd=str(tmp_path),
                         capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    line = json.loads([l for l in res.stdout.splitlines() if l.startswith("{")][-1])
    assert line["metric"] == "e2e_eval_images_per_sec" and line["images"] == 48 and len(line["epochs"]) == 2
    first, second = line["epochs"]
    assert first["cache_misses"] == 48 and first["cache_hits"] == 0
    assert second["cache_hits"] == 48 and second["cache_misses"] == 48            # cumulative counters
    assert second["ms_backbones_or_cache"] < first["ms_backbones_or_cache"]
    assert line["cached_epoch_images_per_sec"] > line["first_epoch_images_per_sec"]
