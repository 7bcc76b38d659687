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


_ALIGN = 64   # elements: every tensor starts on a 256-byte boundary of the flat buffer (TMA needs 16 bytes)


def flat_layout(params, align=_ALIGN):
    """(offsets, padded total) of `params` laid out back to back with each start rounded up to `align` elements."""
    offsets, off = [], 0
    for p in params:
        offsets.append(off)
        off += (p.numel() + align - 1) // align * align
    return offsets, off


class FlatGrads:
    """All trainable gradients live in ONE flat fp32 buffer (per-parameter `.grad` tensors are views of it), so a
    step needs a single all-reduce launch and clip/optimizer see ordinary `.grad` tensors.  The gaps that align the
    tensors stay zero."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        self.offsets, n = flat_layout(self.params)
        ref = self.params[0]
        self.flat = torch.zeros(n, dtype=ref.dtype, device=ref.device)
        for p, off in zip(self.params, self.offsets):
            p.grad = self.flat[off:off + p.numel()].view_as(p)

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


def attach(model, group=None, flat_params=False):
    """Make `model`'s losses use global denominators over `group` and return the FlatGrads (or, with
    `flat_params=True`, the FlatParams needed by FlatAdamax) of its trainable parameters: call `.zero()` instead of
    `optimizer.zero_grad()` and `.all_reduce()` after `backward()`."""
    if dist.is_available() and dist.is_initialized():
        model.loss_group = group if group is not None else dist.group.WORLD
    else:
        model.loss_group = None
    flat = FlatParams(model.parameters()) if flat_params else FlatGrads(model.parameters())
    model._flat = flat
    return flat


class FlatParams(FlatGrads):
    """Parameters AND gradients of the trainable tensors as two flat fp32 buffers (each tensor keeps its shape as a
    view), so clip + optimizer run as one fused kernel over contiguous memory (`FlatAdamax`)."""

    def __init__(self, params):
        params = [p for p in params if p.requires_grad]
        self.offsets, n_pad = flat_layout(params)
        ref = params[0]
        self.flat_param = torch.zeros(n_pad, dtype=ref.dtype, device=ref.device)
        with torch.no_grad():
            for p, off in zip(params, self.offsets):
                view = self.flat_param[off:off + p.numel()].view_as(p)
                view.copy_(p.data)
                p.data = view
        self.params = params
        self.version = 0   # bumped by FlatAdamax.step(): raw kernels do not touch torch's tensor version counters
        self.flat = torch.zeros(n_pad, dtype=ref.dtype, device=ref.device)
        for p, off in zip(params, self.offsets):
            p.grad = self.flat[off:off + p.numel()].view_as(p)


class FlatAdamax(torch.optim.Optimizer):
    """`clip_grad_norm_(params, max_norm)` + `torch.optim.Adamax` (sr.py:80-83,472-473) as ONE fused CUDA kernel over the
    flat buffers of `FlatParams` (srg_clip_adamax).  The per-parameter state (`step`, `exp_avg`, `exp_inf`) is exposed
    through the usual `state_dict()` so checkpoints stay interchangeable with torch.optim.Adamax."""

    fused_clip = True   # the gradient clipping is part of step()

    def __init__(self, flat, lr=0.002, betas=(0.9, 0.999), eps=1e-8, max_norm=1.0):
        if not isinstance(flat, FlatParams):
            raise TypeError("FlatAdamax needs parallel.FlatParams")
        super().__init__(flat.params, dict(lr=lr, betas=betas, eps=eps, max_norm=max_norm))
        self.flat = flat
        dev = flat.flat.device
        self.exp_avg = torch.zeros_like(flat.flat)
        self.exp_inf = torch.zeros_like(flat.flat)
        self.scratch = torch.zeros(2, dtype=torch.float32, device=dev)   # {||g||^2, steps taken}
        for p, off in zip(flat.params, flat.offsets):
            n = p.numel()
            self.state[p] = {"step": self.scratch[1], "exp_avg": self.exp_avg[off:off + n].view_as(p),
                             "exp_inf": self.exp_inf[off:off + n].view_as(p)}

    @torch.no_grad()
    def step(self, closure=None):
        from . import _lib
        g = self.param_groups[0]
        lib = _lib.load()
        _lib.check(lib.srg_clip_adamax(_lib.ptr(self.flat.flat_param), _lib.ptr(self.flat.flat), _lib.ptr(self.exp_avg),
                                       _lib.ptr(self.exp_inf), self.flat.flat.numel(), g["lr"], g["betas"][0],
                                       g["betas"][1], g["eps"], g["max_norm"], _lib.ptr(self.scratch),
                                       _lib.stream_ptr()))
        self.flat.version += 1     # the packed bf16 weights of the model are stale now

    def total_norm(self):
        """Gradient norm seen by the last step (before clipping), like the return value of clip_grad_norm_."""
        return self.scratch[0].sqrt()

    def state_dict(self):
        sd = super().state_dict()
        for st in sd["state"].values():       # materialise views so the checkpoint does not alias the live buffers
            for k in list(st):
                st[k] = st[k].detach().clone()
        return sd

    def load_state_dict(self, state_dict):
        groups = state_dict["param_groups"]
        ids = [i for grp in groups for i in grp["params"]]
        for pid, p in zip(ids, self.flat.params):
            st = state_dict["state"].get(pid)
            if st is None:
                continue
            self.state[p]["exp_avg"].copy_(st["exp_avg"])
            self.state[p]["exp_inf"].copy_(st["exp_inf"])
            self.scratch[1] = float(st["step"])
        for k in ("lr", "betas", "eps"):
            if k in groups[0]:
                self.param_groups[0][k] = groups[0][k]
