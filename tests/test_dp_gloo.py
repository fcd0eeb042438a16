"""CPU test of the N>1 host logic (world_size 2, gloo): bag sharding, the flat gradient buffer layout, the single
all-reduce and the 1/world scaling + Adam update.  Per-rank gradients come from the float64 oracle (the CUDA kernels
cannot run here); the check is that the data-parallel result equals the single-process result on the union of the
shards — the property train_ddp.py:79 relies on DDP for."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import mil_oracle as mo

L_FEAT, D = 64, 192
N_BAGS = 7


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _bags():
    lens = mo.ragged_lengths(N_BAGS, 3, 60, 5)
    rs = np.random.RandomState(9)
    return lens, [rs.standard_normal((int(n), L_FEAT)).astype(np.float32) for n in lens]


def _flat_grad(p, bags, idx):
    from mil_b200.dp import flat_layout
    lay = flat_layout(L_FEAT, D)
    flat = np.zeros(2 * D * L_FEAT + 3 * D + 1, dtype=np.float64)
    for i in idx:
        g = mo.abmil_backward(p, bags[i], np.ones(L_FEAT), need_dx=False)
        for k, (a, b) in lay.items():
            flat[a:b] += np.asarray(g[k], dtype=np.float64).reshape(-1)
    return flat


def _worker(rank, world, port, balance, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import mil_b200  # noqa: F401  (loads the package; no CUDA call is made)
        from mil_b200.dp import AbmilTrainer, flat_layout, shard_bags
        p = mo.procedural_state(mo.abmil_shapes(L_FEAT, D), 3)
        lens, bags = _bags()
        tr = AbmilTrainer(L_FEAT, D, torch.float32, device="cpu", process_group=dist.group.WORLD, world_size=world)
        lay = flat_layout(L_FEAT, D)
        if rank == 0:                                # only rank 0 holds the real parameters before the broadcast
            for k, (a, b) in lay.items():
                tr.params[a:b] = torch.from_numpy(p[k].reshape(-1))
        tr.broadcast_params()
        mine = shard_bags(lens, rank, world, balance=balance)
        pr = {k: tr.params[a:b].numpy().reshape(p[k].shape) for k, (a, b) in lay.items()}
        tr.grads.copy_(torch.from_numpy(_flat_grad(pr, bags, mine)).float())
        tr.allreduce_grads()
        out[rank] = (mine, tr.grads.clone().numpy(), tr.params.clone().numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("balance", [False, True])
def test_world2_allreduce_equals_single_process(balance):
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), balance, out), nprocs=world, join=True)
    lens, bags = _bags()
    p = mo.procedural_state(mo.abmil_shapes(L_FEAT, D), 3)
    shards = [out[r][0] for r in range(world)]
    assert sorted(shards[0] + shards[1]) == list(range(N_BAGS)) and not set(shards[0]) & set(shards[1])
    if balance:
        loads = [int(sum(lens[i] for i in s)) for s in shards]
        assert abs(loads[0] - loads[1]) <= int(max(lens))
    else:
        assert shards[0] == list(range(0, N_BAGS, 2))                     # DistributedSampler's stride
    ref = _flat_grad(p, bags, range(N_BAGS))
    for r in range(world):
        assert np.allclose(out[r][2], out[0][2])                           # broadcast: identical replicas
        g = out[r][1]
        assert np.abs(g - ref).max() <= 1e-5 * np.abs(ref).max()           # sum over ranks == single process
    # DDP averages: Adam on grad/world equals the single-process Adam on the mean gradient
    m0 = v0 = np.zeros_like(ref)
    a = mo.adam_step(out[0][2].astype(np.float64), out[0][1].astype(np.float64) / world, m0, v0, 1)
    b = mo.adam_step(out[0][2].astype(np.float64), ref / world, m0, v0, 1)
    assert np.abs(a[0] - b[0]).max() <= 1e-9


def test_flat_layout_matches_state_dict():
    import mil_b200
    from mil_b200.dp import flat_layout
    m = mil_b200.ABMIL(None, L=L_FEAT, D=D)
    lay = flat_layout(L_FEAT, D)
    assert set(lay) == set(m.state_dict())
    assert all(b - a == m.state_dict()[k].numel() for k, (a, b) in lay.items())
    spans = sorted(lay.values())
    assert spans[0][0] == 0 and all(spans[i][1] == spans[i + 1][0] for i in range(len(spans) - 1))
