"""CPU: host-side logic added in round 2 -- flat buffer layout for equal shards, the sampler's global batch sizes,
the cross-epoch feature cache, the ncu summary tool's unit handling."""
import torch

from situation_recognition_b200 import parallel
from situation_recognition_b200.features import FeatureCache
from situation_recognition_b200.imsitu_loader import ShardedBatchSampler


def test_flat_layout_divides_into_aligned_shards():
    ps = [torch.nn.Parameter(torch.zeros(5, 7)), torch.nn.Parameter(torch.zeros(100)), torch.nn.Parameter(torch.zeros(3))]
    offsets, total = parallel.flat_layout(ps, total_multiple=64 * 8)
    assert offsets == [0, 64, 192] and total % (64 * 8) == 0 and total >= 256
    for world in (1, 2, 4, 8):
        assert (total // world) % 64 == 0                     # every shard starts on a 256-byte boundary


def test_sampler_reports_the_global_batch_sizes():
    for n, gb, world in ((10, 4, 2), (25, 8, 4), (1974 + 6144, 6144, 8), (5, 4, 4)):
        samplers = [ShardedBatchSampler(n, gb, r, world, shuffle=True, seed=1) for r in range(world)]
        sizes = samplers[0].global_sizes()
        per_rank = [list(s) for s in samplers]
        assert len(sizes) == len(per_rank[0]) == len(samplers[0])
        for step, want in enumerate(sizes):
            assert sum(len(pr[step]) for pr in per_rank) == want      # the verb loss divides by exactly this
    # the imSitu train tail over 8 ranks: unequal shards (rows * world is not the denominator)
    s = [ShardedBatchSampler(1974, 6144, r, 8) for r in range(8)]
    assert sorted({len(next(iter(x))) for x in s}) == [246, 247] and s[0].global_sizes() == [1974]


def test_feature_cache_roundtrip():
    cache = FeatureCache(5, 8, torch.device("cpu"))
    fv, fn = torch.arange(24.).view(3, 8), -torch.arange(24.).view(3, 8)
    assert cache.lookup(["a", "b", "c"]) is None and not cache.has(["a"])
    cache.store(["a", "b", "c"], fv, fn)
    assert len(cache) == 3 and cache.has(["c", "a"])
    got_v, got_n = cache.lookup(["c", "a"])
    assert torch.equal(got_v, fv[[2, 0]]) and torch.equal(got_n, fn[[2, 0]])
    cache.store(["b"], fv[:1] * 0 + 7, fn[:1])                       # overwriting keeps the row
    assert len(cache) == 3 and torch.equal(cache.lookup(["b"])[0], torch.full((1, 8), 7.))
    assert cache.lookup(["a", "zzz"]) is None                         # one miss -> the batch is recomputed
    cache.store(["d", "e"], fv[:2], fn[:2])
    try:
        cache.store(["f"], fv[:1], fn[:1])
        raise AssertionError("capacity must be enforced")
    except ValueError:
        pass
    assert cache.hits == 3 and cache.misses == 5
