"""world_size-2 gloo test (CPU) of the one-process-per-GPU plumbing: contiguous sharding, global loss
denominators and the single flat-buffer gradient all-reduce reproduce the full-batch gradients.
The arithmetic here is the CPU oracle (the CUDA path is covered by the -m gpu tests)."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _setup(D=64, B=10):
    from oracle import ggnn_oracle as O
    from situation_recognition_b200.imsitu_encoder import imsitu_encoder
    from situation_recognition_b200.synthetic import make_batch, make_train_json
    enc = imsitu_encoder(make_train_json(seed=0, images_per_verb=1), verbose=False)
    params = O.init_params(enc.get_num_verbs(), enc.get_num_roles(), enc.get_num_labels(), D, seed=0)
    batch = make_batch(enc, B, D, seed=3)
    tables = O.build_tables(enc.roles_per_verb, enc.verb_list, enc.role_list)
    return O, enc, params, batch, tables


def _shard_loss(O, ps, batch, tables, L, lo, hi, counts, B_global):
    """verb_loss + nouns_loss of one shard with GLOBAL denominators (what FCGGNN.loss_group implements on CUDA)."""
    fv, fn, gv, gn = [x[lo:hi] for x in batch]
    t, c = tables
    pv, pn, _ = O.forward(ps, fv, fn, gv, t, c)
    loss = F.cross_entropy(pv, gv, reduction="sum") / B_global
    pn_t = pn.transpose(1, 2)
    for a in range(3):
        loss = loss + F.cross_entropy(pn_t, gn[:, a], ignore_index=L, reduction="sum") / counts[a]
    return loss


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from situation_recognition_b200 import parallel
    O, enc, params, batch, tables = _setup()
    B, L = batch[0].shape[0], enc.get_num_labels()
    lo, hi = parallel.shard_range(B, rank, world)
    ps = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    flat = parallel.FlatGrads(list(ps.values()))
    counts = torch.tensor([(batch[3][lo:hi, a] != L).sum() for a in range(3)], dtype=torch.float32)
    dist.all_reduce(counts)                                   # global non-ignored counts (3 floats)
    flat.zero()
    _shard_loss(O, ps, batch, tables, L, lo, hi, counts, B).backward()
    assert all(p.grad.data_ptr() >= flat.flat.data_ptr() for p in flat.params)   # grads are views of the flat buffer
    flat.all_reduce()
    if rank == 0:
        torch.save({k: v.grad.clone() for k, v in ps.items()}, out)
    dist.destroy_process_group()


def test_two_rank_gradients_equal_full_batch(tmp_path):
    out = str(tmp_path / "grads.pt")
    port = _free_port()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out)
    O, enc, params, batch, tables = _setup()
    fv, fn, gv, gn = batch
    t, c = tables
    _, ref, _ = O.train_step_grads(params, fv, fn, gv, gn, t, c, enc.get_num_labels())
    for k, v in ref.items():
        scale = max(v.abs().max().item(), 1e-12)
        assert (got[k] - v).abs().max().item() <= 2e-5 * scale + 1e-9, k


def test_shard_range_covers_batch():
    from situation_recognition_b200.parallel import shard_range
    for n in (0, 1, 7, 6144):
        for w in (1, 2, 4, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
