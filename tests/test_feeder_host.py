"""Host side of the feeder (dataset.py:366-393): the native packer against numpy/torch, bit for
bit, and the iterator's batching/sharding against DistributedSampler's rule.  No GPU needed (the packer makes no CUDA
calls); the device leg is in tests/test_gpu_feeder.py."""
import random

import numpy as np
import pytest
import torch

import mil_b200
from mil_b200 import feeder


def _bags(lens, L, dtype, seed=0):
    rng = np.random.default_rng(seed)
    return [(rng.standard_normal((n, L)) * 3).astype(dtype) for n in lens]


def _expect(bags, keep, out_dtype):
    rows = [b if k is None else b[k] for b, k in zip(bags, keep)]
    cat = torch.from_numpy(np.concatenate(rows, 0).astype(np.float32) if bags[0].dtype != np.float64
                           else np.concatenate(rows, 0)).float()     # reference: torch.from_numpy(feature).float()
    return cat.to(out_dtype)


@pytest.mark.parametrize("src", [np.float32, np.float64, np.float16])
@pytest.mark.parametrize("dst", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("threads", [1, 5])
def test_pack_is_bit_exact(src, dst, threads):
    lens = [1, 37, 300, 2, 129, 64]
    bags = _bags(lens, 96, src, seed=3)
    out = torch.full((sum(lens) + 5, 96), 7.0, dtype=dst)
    off = torch.zeros(len(lens) + 1, dtype=torch.int32)
    n = feeder.pack_bags_host(bags, out, off, n_threads=threads)
    assert n == sum(lens)
    assert off.tolist() == [0] + np.cumsum(lens).tolist()                 # offsets: bit-exact
    want = _expect(bags, [None] * len(lens), dst)
    assert torch.equal(out[:n].view(torch.int16 if dst == torch.bfloat16 else torch.int32),
                       want.view(torch.int16 if dst == torch.bfloat16 else torch.int32))
    assert float(out[n:].min()) == 7.0                                    # rows past the batch untouched


def test_special_values_round_like_torch():
    a = np.array([[0.0, -0.0, np.inf, -np.inf, np.nan, 1e-45, 3.3895314e38, 1.0039062, 1.0117188, 65504.0, -1e-40, 2.0 ** -133]],
                 dtype=np.float32)
    out = torch.empty((1, a.shape[1]), dtype=torch.bfloat16)
    off = torch.zeros(2, dtype=torch.int32)
    feeder.pack_bags_host([a], out, off)
    want = torch.from_numpy(a).to(torch.bfloat16)
    same = out.view(torch.int16) == want.view(torch.int16)
    assert bool((same | (torch.isnan(out) & torch.isnan(want))).all())
    h = np.array([[6e-8, -6.1e-5, 65504.0, np.inf, 0.0, 1.0009766]], dtype=np.float16)      # subnormal halves too
    out = torch.empty((1, h.shape[1]), dtype=torch.float32)
    feeder.pack_bags_host([h], out, off)
    assert torch.equal(out, torch.from_numpy(h).float())


def test_augmentation_subset_and_strided_sources():
    rng = random.Random(5)
    lens = [50, 200, 11]
    bags = _bags(lens, 64, np.float32, seed=9)
    wide = np.zeros((200, 80), dtype=np.float32)
    wide[:, :64] = bags[1]
    bags[1] = wide[:, :64]                                                # row pitch 320 B, rows still contiguous
    keep = [feeder.augmentation_rows(50, "Biopsy", rng), feeder.augmentation_rows(200, "Resection", rng), None]
    assert len(keep[0]) == 45 and len(keep[1]) == 160                     # dataset.py:378,381: int(n * 0.9), int(n * 0.8)
    assert all(np.all(np.diff(k) > 0) for k in keep[:2])                  # `sorted(random.sample(...))`
    out = torch.empty((45 + 160 + 11, 64), dtype=torch.bfloat16)
    off = torch.zeros(4, dtype=torch.int32)
    n = feeder.pack_bags_host(bags, out, off, keep_rows=keep, n_threads=3)
    assert n == 216 and off.tolist() == [0, 45, 205, 216]
    assert torch.equal(out.view(torch.int16), _expect(bags, keep, torch.bfloat16).view(torch.int16))


def test_bf16_torch_sources_pass_through():
    a = torch.randn(33, 32).to(torch.bfloat16)
    b = torch.randn(5, 32).to(torch.bfloat16)
    out = torch.empty((38, 32), dtype=torch.bfloat16)
    off = torch.zeros(3, dtype=torch.int32)
    feeder.pack_bags_host([a, b], out, off)
    assert torch.equal(out, torch.cat([a, b]))


def test_errors_are_raised_not_swallowed():
    bags = _bags([4, 4], 16, np.float32)
    off = torch.zeros(3, dtype=torch.int32)
    with pytest.raises(mil_b200.MilB200Error):                            # destination too small
        feeder.pack_bags_host(bags, torch.empty((7, 16), dtype=torch.bfloat16), off)
    with pytest.raises(mil_b200.MilB200Error):                            # unsorted subset
        feeder.pack_bags_host(bags, torch.empty((8, 16), dtype=torch.bfloat16), off,
                              keep_rows=[np.array([2, 1], dtype=np.int32), None])
    with pytest.raises(mil_b200.MilB200Error):                            # empty bag
        feeder.pack_bags_host([bags[0], np.zeros((0, 16), np.float32)], torch.empty((8, 16), dtype=torch.bfloat16), off)
    with pytest.raises(ValueError):                                       # feature width mismatch
        feeder.pack_bags_host(bags, torch.empty((8, 32), dtype=torch.bfloat16), off)
    with pytest.raises(ValueError):                                       # mixed source dtypes
        feeder.pack_bags_host([bags[0], bags[1].astype(np.float64)], torch.empty((8, 16), dtype=torch.bfloat16), off)


def test_feeder_batches_follow_distributed_sampler(tmp_path):
    lens = [7, 30, 2, 19, 5, 11, 23]
    bags = _bags(lens, 32, np.float32, seed=1)
    paths = []
    for i, b in enumerate(bags):                                          # the reference's storage: one .npy per slide
        p = tmp_path / f"patient{i}.npy"
        np.save(p, b)
        paths.append(str(p))
    seen = []
    for rank in (0, 1):
        f = feeder.PackedBagFeeder(paths, batch_bags=2, L_feat=32, device="cpu", dtype=torch.bfloat16, rank=rank, world=2)
        # train_ddp.py:191 without shuffle: DistributedSampler pads 7 -> 8 indices with the head of the list
        assert f.order() == (list(range(7)) + [0])[rank::2]
        assert len(f.batches()) == 2                                      # every rank runs the same number of steps
        for X, off, ids in f:
            assert off.dtype == torch.int32 and off[0] == 0
            assert off.tolist() == [0] + np.cumsum([lens[i] for i in ids]).tolist()
            want = _expect([bags[i] for i in ids], [None] * len(ids), torch.bfloat16)
            assert torch.equal(X.view(torch.int16), want.view(torch.int16))
            seen += ids
    assert sorted(seen) == [0] + list(range(7))                           # the padding index is seen twice
    for world in (2, 3, 4, 8):                                            # equal step counts for any world size
        counts = {len(feeder.PackedBagFeeder(paths, batch_bags=2, L_feat=32, device="cpu", rank=r, world=world,
                                             shuffle=True, seed=5).batches()) for r in range(world)}
        assert len(counts) == 1, (world, counts)
    f = feeder.PackedBagFeeder(paths, batch_bags=3, L_feat=32, device="cpu", shuffle=True, seed=3)
    f.set_epoch(0)
    a = f.order()
    f.set_epoch(1)
    assert sorted(a) == list(range(7)) and f.order() != a                 # reshuffled per epoch, still a permutation


def test_abandoned_iteration_stops_the_worker_and_the_feeder_can_be_reused():
    import threading
    bags = _bags([5, 9, 3, 12, 7, 4, 8, 6], 16, np.float32, seed=2)
    f = feeder.PackedBagFeeder(bags, batch_bags=2, L_feat=16, device="cpu", dtype=torch.float32)
    before = threading.active_count()
    it = iter(f)
    X, off, ids = next(it)
    assert ids == [0, 1] and off.tolist() == [0, 5, 14]
    it.close()                                  # the consumer walks away after one batch
    assert threading.active_count() == before   # ... and the packing thread is gone, not parked on a full queue
    assert [ids for _, _, ids in f] == [[0, 1], [2, 3], [4, 5], [6, 7]]
