"""-m gpu: the one-process-per-GPU path of the PRODUCT code (FCGGNN with `loss_group`, the flat gradient buffer and
its all-reduce) against single-process full-batch gradients, on ONE GPU.

Two ranks (two processes sharing cuda:0) exchange through the `gloo` backend, which accepts CUDA tensors: the ranks'
kernels never wait on each other on the device (the exchange is a host-side rendezvous), so this is safe on a single
GPU, unlike two NCCL ranks.  Each rank runs `FCGGNN.forward`, `verb_loss`, `nouns_loss` and `backward` on its
contiguous shard of the batch -- the shards are UNEQUAL (11 and 10 images) -- then `FlatGrads.all_reduce()`.
Reference semantics: the reference's DataParallel computes every cross-entropy mean over the GLOBAL batch on GPU 0
(sr.py:66-76, 467-470), so the summed shard losses / gradients must equal the full-batch ones.
"""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu

B, D = 21, 256


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _setup():
    import situation_recognition_b200 as S
    from oracle import ggnn_oracle as O
    from situation_recognition_b200.synthetic import make_batch, make_train_json
    enc = S.imsitu_encoder(make_train_json(seed=0, images_per_verb=1), verbose=False)
    params = O.init_params(enc.get_num_verbs(), enc.get_num_roles(), enc.get_num_labels(), D, seed=0)
    batch = make_batch(enc, B, D, seed=3)
    return S, O, enc, params, batch


def _model(S, enc, params):
    m = S.FCGGNN(enc, D, backbone=None, precision="bf16")
    m.load_state_dict(params, strict=False)
    return m.cuda().eval()          # eval: no dropout, so the shards and the full batch see the same arithmetic


def _step(m, flat, batch, lo, hi):
    fv, fn, gv, gn = [x[lo:hi].cuda() for x in batch]
    flat.zero()
    pv, pn, gpn = m(fv, gv, img_nouns=fn)
    vl, nl, gl = m.verb_loss(pv, gv), m.nouns_loss(pn, gn), m.nouns_loss(gpn, gn)
    (vl + nl).backward()
    return torch.stack([vl.detach(), nl.detach(), gl.detach()])


def _worker(rank, world, port, out, know_global_batch):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(0)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from situation_recognition_b200 import parallel
    S, O, enc, params, batch = _setup()
    m = _model(S, enc, params)
    flat = parallel.attach(m)                     # loss_group = WORLD, flat gradient buffer
    assert m.loss_group is not None
    m.global_batch = B if know_global_batch else None
    lo, hi = parallel.shard_range(B, rank, world)
    assert len({h - l for l, h in (parallel.shard_range(B, r, world) for r in range(world))}) > 1   # unequal shards
    losses = _step(m, flat, batch, lo, hi)
    flat.all_reduce()
    dist.all_reduce(losses)                       # per-rank losses are partial sums of the global-batch means
    torch.cuda.synchronize()
    if rank == 0:
        torch.save({"losses": losses.cpu(), "grads": {k: p.grad.detach().cpu().clone() for k, p in m.named_parameters()}},
                   out)
    dist.barrier()
    dist.destroy_process_group()


def _worker_sharded_opt(rank, world, port, out):
    """Two training steps with the SHARDED fused optimizer (reduce-scatter, 1/G clip + Adamax, all-gather)."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(0)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from situation_recognition_b200 import parallel
    S, O, enc, params, batch = _setup()
    m = _model(S, enc, params)
    flat = parallel.attach(m, flat_params=True)
    opt = parallel.FlatAdamax(flat, lr=0.01, max_norm=1.0, group=dist.group.WORLD)
    assert opt.world == 2 and opt.exp_avg.numel() * 2 == flat.flat.numel()
    m.global_batch = B
    lo, hi = parallel.shard_range(B, rank, world)
    first = None
    for it in range(2):
        _step(m, flat, batch, lo, hi)
        opt.step()                                # no flat.all_reduce(): the reduce-scatter is part of the step
        if it == 0:
            first = opt.state_dict()["state"][0]["exp_inf"].cpu()     # collective: gathers the sharded state
    sd = opt.state_dict()
    torch.cuda.synchronize()
    if rank == 0:
        torch.save({"params": {k: v.detach().cpu().clone() for k, v in m.state_dict().items()},
                    "norm": opt.total_norm().item(), "exp_inf0": first,
                    "step": float(sd["state"][0]["step"])}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_optimizer_equals_replicated(tmp_path):
    out = str(tmp_path / "opt.pt")
    mp.spawn(_worker_sharded_opt, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    from situation_recognition_b200 import parallel
    S, O, enc, params, batch = _setup()
    m = _model(S, enc, params)
    flat = parallel.attach(m, flat_params=True)
    opt = parallel.FlatAdamax(flat, lr=0.01, max_norm=1.0)
    first = None
    for it in range(2):
        _step(m, flat, batch, 0, B)
        opt.step()
        if it == 0:
            first = opt.state_dict()["state"][0]["exp_inf"].cpu()
    torch.cuda.synchronize()
    assert got["step"] == 2.0
    # after ONE step the optimizer state only differs by the order of the gradient sums (the second step already sees
    # slightly different weights: see the sign-flip note below)
    assert (got["exp_inf0"] - first).abs().max().item() <= 1e-3 * first.abs().max().item()
    assert abs(got["norm"] - opt.total_norm().item()) <= 5e-3 * opt.total_norm().item()
    moved = 0.0
    for k, v in m.state_dict().items():
        diff = (got["params"][k] - v.cpu()).abs()
        # the first Adamax step moves every weight by lr * sign(g): elements whose gradient is within the rounding noise of
        # zero may go either way (see test_graphed_step_matches_eager; W_p / W_x see ~2e-3 of max|g| of bf16 noise from
        # the per-rank rounding of d/dP); everything else agrees far below the update size
        assert (diff > 1e-3).float().mean().item() <= 1e-2 and diff.max().item() <= 4 * 0.01 + 1e-3, (k, diff.max().item())
        moved = max(moved, (v.cpu() - params[k]).abs().max().item())
    assert moved > 0.01
    sd = opt.state_dict()
    # the saved state loads into torch.optim.Adamax (checkpoint compatibility, sr.py:28-41)
    ref = torch.optim.Adamax([p for p in m.parameters() if p.requires_grad], lr=0.01)
    ref.load_state_dict(sd)
    for p in m.parameters():
        p.grad = torch.zeros_like(p)
    ref.step()


@pytest.mark.parametrize("know_global_batch", [False, True])
def test_two_cuda_ranks_equal_full_batch(tmp_path, know_global_batch):
    out = str(tmp_path / "ranks.pt")
    mp.spawn(_worker, args=(2, _free_port(), out, know_global_batch), nprocs=2, join=True)
    got = torch.load(out)

    from situation_recognition_b200 import parallel
    S, O, enc, params, batch = _setup()
    m = _model(S, enc, params)
    flat = parallel.attach(m)                     # no process group here: plain full-batch denominators
    assert m.loss_group is None
    losses = _step(m, flat, batch, 0, B).cpu()
    full = {k: p.grad.detach().cpu().clone() for k, p in m.named_parameters()}
    pred = None
    with torch.no_grad():
        pred = m.predict_verb(batch[0].cuda(), B).argmax(-1).cpu()

    # (a) against the same CUDA arithmetic on the whole batch: the order of the gradient sums differs, and the chain rule
    # through the folded message weights rounds d/dP to bf16 per rank instead of once (W_p, W_z, W_r, W_h: ~2e-3)
    assert torch.allclose(got["losses"], losses, rtol=2e-5, atol=1e-6), (got["losses"], losses)
    worst = 0.0
    for k, v in full.items():
        scale = max(v.abs().max().item(), 1e-12)
        err = (got["grads"][k] - v).abs().max().item() / scale
        worst = max(worst, err)
        assert err <= 5e-3, (k, err)
    # (b) against the oracle (reference arithmetic, fp32) under the CUDA path's predicted verbs
    t, c = O.build_tables(enc.roles_per_verb, enc.verb_list, enc.role_list)
    (vl, nl, gl), ref, _ = O.train_step_grads(params, *batch, t, c, enc.get_num_labels(), pred_verbs=pred)
    for mine, r in zip(got["losses"].tolist(), (vl, nl, gl)):
        assert abs(mine - float(r)) <= 5e-3 * float(r)
    for k, v in ref.items():
        scale = max(v.abs().max().item(), 1e-12)
        assert (got["grads"][k] - v).abs().max().item() / scale <= 3e-2, k
    print("PARITY two_cuda_ranks: worst shard-sum vs full-batch grad err %.2e" % worst)
