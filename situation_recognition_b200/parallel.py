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


def flat_layout(params, align=_ALIGN, total_multiple=1):
    """(offsets, padded total) of `params` laid out back to back with each start rounded up to `align` elements; the
    total is rounded up to a multiple of `total_multiple` (equal shards per rank for the sharded optimizer)."""
    offsets, off = [], 0
    for p in params:
        offsets.append(off)
        off += (p.numel() + align - 1) // align * align
    m = max(1, total_multiple)
    return offsets, (off + m - 1) // m * m


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


def _world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group)
    return 1


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


def broadcast_model(model, src=0, group=None):
    """Every rank constructs the model itself (random init of the GGNN stage, possibly of the backbones): make rank
    `src`'s parameters and buffers everyone's.  The reference has ONE process whose `nn.DataParallel` re-broadcasts the
    parameters at every step (sr.py:467-470); here they are sent once, and stay identical because every rank applies
    the same update to the same all-reduced gradient."""
    if _world(group) <= 1:
        return
    with torch.no_grad():
        for t in model.state_dict().values():
            if torch.is_tensor(t):
                dist.broadcast(t, src=src, group=group)


class FlatParams(FlatGrads):
    """Parameters AND gradients of the trainable tensors as two flat fp32 buffers (each tensor keeps its shape as a
    view), so clip + optimizer run as one fused kernel over contiguous memory (`FlatAdamax`)."""

    def __init__(self, params):
        params = [p for p in params if p.requires_grad]
        # the padded length divides into equal, 256-byte aligned shards for every world size up to 8 * k
        self.offsets, n_pad = flat_layout(params, total_multiple=_ALIGN * max(8, _world()))
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
    through the usual `state_dict()` so checkpoints stay interchangeable with torch.optim.Adamax.

    With `group` (a process group of G > 1 ranks) the weight-sized work of a step is SHARDED instead of replicated:
    `step()` reduce-scatters the flat gradient (each rank receives the sum of its 1/G slice), all-reduces one scalar
    (the squared gradient norm), clips + applies Adamax to its slice only -- the optimizer state is 1/G per rank --
    and all-gathers the updated parameters.  The two collectives move the bytes of the one all-reduce they replace;
    clip + Adamax + the norm run on 1/G of the 35.9 M parameters.  Do NOT call `FlatGrads.all_reduce()` as well."""

    fused_clip = True   # the gradient clipping is part of step()

    def __init__(self, flat, lr=0.002, betas=(0.9, 0.999), eps=1e-8, max_norm=1.0, group=None, shard=None):
        if not isinstance(flat, FlatParams):
            raise TypeError("FlatAdamax needs parallel.FlatParams")
        # the standard Adamax keys are carried along so that a saved `optimizer_state_dict` loads into torch.optim.Adamax
        super().__init__(flat.params, dict(lr=lr, betas=betas, eps=eps, weight_decay=0, foreach=None, maximize=False,
                                           differentiable=False, capturable=False, max_norm=max_norm))
        self.flat = flat
        dev = flat.flat.device
        self.group = group
        self.world = _world(group) if (shard or (shard is None and group is not None)) else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        n = flat.flat.numel()
        if n % (self.world * 4) != 0:
            raise ValueError("flat buffer of %d elements does not divide into %d aligned shards" % (n, self.world))
        self.shard_len = n // self.world
        self.lo = self.rank * self.shard_len
        # optimizer state of this rank's slice only
        self.exp_avg = torch.zeros(self.shard_len, dtype=torch.float32, device=dev)
        self.exp_inf = torch.zeros(self.shard_len, dtype=torch.float32, device=dev)
        self.grad_shard = torch.zeros(self.shard_len, dtype=torch.float32, device=dev) if self.world > 1 else None
        self.scratch = torch.zeros(2, dtype=torch.float32, device=dev)   # {||g||^2, steps taken}
        if self.world == 1:
            for p, off in zip(flat.params, flat.offsets):
                k = p.numel()
                self.state[p] = {"step": self.scratch[1], "exp_avg": self.exp_avg[off:off + k].view_as(p),
                                 "exp_inf": self.exp_inf[off:off + k].view_as(p)}

    @torch.no_grad()
    def step(self, closure=None):
        from . import _lib
        g = self.param_groups[0]
        lib = _lib.load()
        f = self.flat
        s = _lib.stream_ptr(f.flat.device)
        if self.world == 1:
            _lib.check(lib.srg_clip_adamax(_lib.ptr(f.flat_param), _lib.ptr(f.flat), _lib.ptr(self.exp_avg),
                                           _lib.ptr(self.exp_inf), f.flat.numel(), g["lr"], g["betas"][0],
                                           g["betas"][1], g["eps"], g["max_norm"], _lib.ptr(self.scratch), s))
        else:
            # sum over ranks of this rank's slice of the gradient (loss denominators are global: no division)
            self._reduce_scatter(self.grad_shard, f.flat)
            _lib.check(lib.srg_sumsq(_lib.ptr(self.grad_shard), self.shard_len, _lib.ptr(self.scratch), s))
            dist.all_reduce(self.scratch[0:1], group=self.group)          # ||g||^2 of the whole gradient
            p_shard = f.flat_param[self.lo:self.lo + self.shard_len]
            _lib.check(lib.srg_adamax_step(_lib.ptr(p_shard), _lib.ptr(self.grad_shard), _lib.ptr(self.exp_avg),
                                           _lib.ptr(self.exp_inf), self.shard_len, g["lr"], g["betas"][0], g["betas"][1],
                                           g["eps"], g["max_norm"], _lib.ptr(self.scratch), _lib.ptr(self.scratch[1:2]),
                                           s))
            self._all_gather(f.flat_param, p_shard)
        f.version += 1     # the packed bf16 weights of the model are stale now

    def _native(self):
        return dist.get_backend(self.group) == "nccl"

    def _reduce_scatter(self, out, full):
        if self._native():
            dist.reduce_scatter_tensor(out, full, op=dist.ReduceOp.SUM, group=self.group)
        else:   # gloo (tests): no reduce-scatter -- all-reduce, keep the own slice
            dist.all_reduce(full, op=dist.ReduceOp.SUM, group=self.group)
            out.copy_(full[self.lo:self.lo + self.shard_len])

    def _all_gather(self, full, part):
        if self._native():
            dist.all_gather_into_tensor(full, part, group=self.group)
        else:
            dist.all_gather(list(full.view(self.world, -1).unbind(0)), part.clone(), group=self.group)

    def total_norm(self):
        """Gradient norm seen by the last step (before clipping), like the return value of clip_grad_norm_."""
        return self.scratch[0].sqrt()

    def _full_state(self):
        """(exp_avg, exp_inf) over the whole flat buffer (gathered from the ranks when the state is sharded)."""
        if self.world == 1:
            return self.exp_avg, self.exp_inf
        out = []
        for t in (self.exp_avg, self.exp_inf):
            full = torch.empty(self.shard_len * self.world, dtype=t.dtype, device=t.device)
            self._all_gather(full, t)
            out.append(full)
        return out

    def state_dict(self):
        """torch.optim.Adamax's format.  Sharded state: a collective -- every rank must call it (each gets the full
        dict, rank 0 writes the checkpoint, sr.py:145-162)."""
        avg, inf = self._full_state()
        f = self.flat
        state = {}
        for i, (p, off) in enumerate(zip(f.params, f.offsets)):
            k = p.numel()
            state[i] = {"step": self.scratch[1].detach().clone(),
                        "exp_avg": avg[off:off + k].view_as(p).detach().clone(),
                        "exp_inf": inf[off:off + k].view_as(p).detach().clone()}
        grp = {k: v for k, v in self.param_groups[0].items() if k != "params"}
        grp["params"] = list(range(len(f.params)))
        return {"state": state, "param_groups": [grp]}

    def load_state_dict(self, state_dict):
        groups = state_dict["param_groups"]
        ids = [i for grp in groups for i in grp["params"]]
        f = self.flat
        n = f.flat.numel()
        avg = torch.zeros(n, dtype=torch.float32, device=f.flat.device)
        inf = torch.zeros(n, dtype=torch.float32, device=f.flat.device)
        for pid, p, off in zip(ids, f.params, f.offsets):
            st = state_dict["state"].get(pid)
            if st is None:
                continue
            k = p.numel()
            avg[off:off + k].copy_(st["exp_avg"].reshape(-1))
            inf[off:off + k].copy_(st["exp_inf"].reshape(-1))
            self.scratch[1] = float(st["step"])
        self.exp_avg.copy_(avg[self.lo:self.lo + self.shard_len])
        self.exp_inf.copy_(inf[self.lo:self.lo + self.shard_len])
        for k in ("lr", "betas", "eps"):
            if k in groups[0]:
                self.param_groups[0][k] = groups[0][k]
