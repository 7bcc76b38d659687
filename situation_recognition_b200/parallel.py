"""One process per GPU: batch sharding, global loss denominators and a single flat-buffer gradient all-reduce.

Replaces the reference's `torch.nn.DataParallel` (sr.py:467-470), which replicates the parameters, scatters the batch,
gathers the logits to GPU 0 and reduces the gradients to GPU 0 every step from ONE Python process.  Images are
independent in the GGNN stage, so the only exchange per step is the gradient sum (35.9 M fp32 values = 143.7 MB)
plus 3 floats of loss denominators.  The collective is NCCL over NVLink 5 / NVSwitch (`backend="nccl"`); the same
code runs on `gloo` for the CPU tests.
"""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous shard [lo, hi) of `n` items for `rank` of `world` (DataParallel's scatter order: dim 0 chunks)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class FlatGrads:
    """All trainable gradients live in ONE flat fp32 buffer (per-parameter `.grad` tensors are views of it), so a
    step needs a single all-reduce launch and clip/optimizer see ordinary `.grad` tensors."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        ref = self.params[0]
        self.flat = torch.zeros(n, dtype=ref.dtype, device=ref.device)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero(self):
        self.flat.zero_()

    def all_reduce(self, group=None, async_op=False):
        """Sum over ranks.  Loss denominators are already global (see `FCGGNN.loss_group`), so the sum of the
        per-shard gradients IS the full-batch gradient -- no division by the world size."""
        if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
            return None
        return dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)

    def nbytes(self):
        return self.flat.numel() * self.flat.element_size()


def attach(model, group=None):
    """Make `model`'s losses use global denominators over `group` and return the FlatGrads of its trainable
    parameters (call `.zero()` instead of `optimizer.zero_grad()`, `.all_reduce()` after `backward()`)."""
    if dist.is_available() and dist.is_initialized():
        model.loss_group = group if group is not None else dist.group.WORLD
    else:
        model.loss_group = None
    return FlatGrads(model.parameters())
