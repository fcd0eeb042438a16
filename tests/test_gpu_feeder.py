"""Device leg of the feeder: batches that arrive through pinned staging + the copy stream are the
same bits as a direct pack, the rotating buffers are not overwritten while a batch is still in use, and the packed
batch drives the pool exactly like a hand-built one."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _bags(lens, L, seed):
    rng = np.random.default_rng(seed)
    return [rng.standard_normal((n, L)).astype(np.float32) for n in lens]


def test_feeder_batches_are_bit_exact_and_buffers_rotate_safely():
    import mil_b200
    from mil_b200 import feeder
    lens = [300, 17, 1, 950, 64, 129, 511, 2, 77, 400, 33]
    bags = _bags(lens, 256, seed=4)
    f = feeder.PackedBagFeeder(bags, batch_bags=3, L_feat=256, device="cuda", dtype=torch.bfloat16)
    torch.manual_seed(0)
    m = mil_b200.ABMIL(None, L=256).cuda().eval()              # fp32 master weights, bf16 instances
    held = []
    for X, off, ids in f:
        want = torch.cat([torch.from_numpy(bags[i]) for i in ids]).to(torch.bfloat16).cuda()
        want_off = torch.tensor([0] + np.cumsum([lens[i] for i in ids]).tolist(), dtype=torch.int32, device="cuda")
        assert torch.equal(off, want_off)                                   # bag offsets: bit-exact
        assert torch.equal(X.view(torch.int16), want.view(torch.int16))
        M = m.forward_csr(X, off)                                           # consumer work on the current stream
        M_direct = m.forward_csr(want, want_off)
        assert torch.equal(M, M_direct)
        held.append((X, want))
        if len(held) >= 2:                                                  # the previous batch's view is still intact
            Xp, wp = held[-2]
            assert Xp.data_ptr() != X.data_ptr()
    assert len(held) == 4


def test_feeder_feeds_the_trainer_step():
    from mil_b200 import feeder
    from mil_b200.dp import AbmilTrainer
    import mil_b200
    lens = [700, 45, 1200, 333, 90, 2048]
    bags = _bags(lens, 1024, seed=8)
    torch.manual_seed(1)
    module = mil_b200.ABMIL(None, L=1024).cuda()
    tr_a = AbmilTrainer(1024, 192, torch.bfloat16, device="cuda")
    tr_b = AbmilTrainer(1024, 192, torch.bfloat16, device="cuda")
    tr_a.load_from(module)
    tr_b.load_from(module)
    f = feeder.PackedBagFeeder(bags, batch_bags=3, L_feat=1024, device="cuda")
    for X, off, ids in f:
        tr_a.step(X, off)
        want = torch.cat([torch.from_numpy(bags[i]) for i in ids]).to(torch.bfloat16).cuda()
        want_off = torch.tensor([0] + np.cumsum([lens[i] for i in ids]).tolist(), dtype=torch.int32, device="cuda")
        tr_b.step(want, want_off)
    assert torch.equal(tr_a.params, tr_b.params)                            # deterministic kernels: identical updates
