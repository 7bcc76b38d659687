"""Flat gradient / parameter buffers (parallel.FlatGrads, FlatParams): every tensor view starts on a 256-byte boundary
whatever the vocabulary sizes are (TMA descriptors and 16-byte vector accesses are built on these views), the views
alias the flat buffer, and the alignment gaps stay zero."""
import torch

from situation_recognition_b200 import parallel


def _params(shapes):
    g = torch.Generator().manual_seed(0)
    return [torch.nn.Parameter(torch.randn(*s, generator=g)) for s in shapes]


def test_flat_layout_aligns_every_tensor():
    shapes = [(191, 64), (503, 64), (64, 64), (64,), (2001, 64), (2001,), (503,), (7,)]     # odd sizes on purpose
    ps = _params(shapes)
    offsets, total = parallel.flat_layout(ps)
    assert all(o % 64 == 0 for o in offsets) and total % 64 == 0
    assert all(offsets[i] + ps[i].numel() <= offsets[i + 1] for i in range(len(ps) - 1))
    assert total >= offsets[-1] + ps[-1].numel()


def test_flat_grads_views_alias_the_buffer():
    ps = _params([(5, 3), (7,), (2, 2)])
    fg = parallel.FlatGrads(ps)
    for p, off in zip(ps, fg.offsets):
        assert p.grad.data_ptr() == fg.flat.data_ptr() + 4 * off and p.grad.shape == p.shape
        p.grad.fill_(1.0)
    assert float(fg.flat.sum()) == sum(p.numel() for p in ps)          # the gaps were not touched
    fg.zero()
    assert all(float(p.grad.abs().sum()) == 0.0 for p in ps)


def test_flat_params_keep_values_and_shapes():
    ps = _params([(5, 3), (7,), (2, 2)])
    before = [p.detach().clone() for p in ps]
    fp = parallel.FlatParams(ps)
    for p, b, off in zip(ps, before, fp.offsets):
        assert torch.equal(p.detach(), b)
        assert p.data_ptr() == fp.flat_param.data_ptr() + 4 * off
    assert fp.flat.numel() == fp.flat_param.numel() and fp.flat.numel() % 64 == 0
