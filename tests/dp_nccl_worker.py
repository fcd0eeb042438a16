"""Worker of tests/test_gpu_dp_nccl.py (launched with torch.distributed.run, one rank per GPU, NCCL).

Checks on real hardware what DDP guarantees upstream (train_ddp.py:79): after the exchange every rank holds the gradient
of the UNION of the ranks' bags — here against a single-process run over all bags on rank 0 — and the replicas' parameters
stay identical after the optimiser step, for (1) the gated-pool trainer over NCCL, (2) the same trainer over the
symmetric-memory exchange kernel (csrc/exchange.cu), (3) the CT+pathology aggregator trainer (FusionTrainer) over NCCL."""
import os
import sys
from argparse import Namespace

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mil_b200  # noqa: E402
from mil_b200.dp import AbmilTrainer, shard_bags  # noqa: E402


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def same_on_all_ranks(t, world):
    """bit-identical across ranks: max and min of every element agree"""
    hi, lo = t.clone(), t.clone()
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    return bool(torch.equal(hi, lo))


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    pg = dist.group.WORLD
    L_feat = 1024
    lens = np.asarray([900, 120, 2500, 33, 1700, 640, 64, 3000, 77, 1300])
    off_all = np.concatenate([[0], np.cumsum(lens)])
    g = torch.Generator().manual_seed(7)                      # every rank draws the same union, then keeps its shard
    X_all = torch.randn(int(off_all[-1]), L_feat, generator=g)
    torch.manual_seed(11)
    module = mil_b200.ABMIL(None, L=L_feat).to(dev)
    mine = shard_bags(lens.tolist(), rank, world, balance=True)
    Xs = torch.cat([X_all[off_all[b]:off_all[b + 1]] for b in mine]).to(dev)
    offs = torch.tensor(np.concatenate([[0], np.cumsum(lens[mine])]), dtype=torch.int32, device=dev)
    for dtype, tol in ((torch.float32, 1e-5), (torch.bfloat16, 2e-3)):
        # (1) NCCL: all-reduced gradients == single-process gradients over the union
        tr = AbmilTrainer(L_feat, 192, dtype, device=dev, process_group=pg, world_size=world, lr=1e-3)
        tr.load_from(module)
        tr.broadcast_params()
        tr.forward_backward(Xs.to(dtype), offs)
        tr.allreduce_grads()
        one = AbmilTrainer(L_feat, 192, dtype, device=dev, lr=1e-3)
        one.load_from(module)
        one.params.copy_(tr.params)
        Xu = X_all.to(dev).to(dtype)
        offu = torch.tensor(off_all, dtype=torch.int32, device=dev)
        one.forward_backward(Xu, offu)
        e = rel(tr.grads, one.grads)
        assert e <= tol, (str(dtype), "all-reduced grads vs union", e)
        assert same_on_all_ranks(tr.grads, world)
        # (2) the symmetric-memory exchange kernel: same sums, same update, replicas bit-identical
        ts = AbmilTrainer(L_feat, 192, dtype, device=dev, process_group=pg, world_size=world, lr=1e-3)
        ts.load_from(module)
        ts.broadcast_params()
        has_symm = ts.enable_symmetric_exchange()
        tn = AbmilTrainer(L_feat, 192, dtype, device=dev, process_group=pg, world_size=world, lr=1e-3)
        tn.load_from(module)
        tn.broadcast_params()
        for step in range(3):
            ts.step(Xs.to(dtype), offs)
            tn.step(Xs.to(dtype), offs)
            if step == 0 and has_symm:      # same parameters, same local gradients: the two exchanges must produce the same sums
                assert rel(ts.grads, tn.grads) <= 1e-6, ("symm vs nccl sums", rel(ts.grads, tn.grads))
        torch.cuda.synchronize()
        if has_symm:
            # three Adam steps later the two replicas sets have moved the same way wherever the gradient is not float
            # noise (Adam turns noise-level gradients into +-lr steps, and those feed back into later gradients)
            live = tn.grads.abs() > 1e-3 * tn.grads.abs().max()
            p0 = module_flat(module, ts)
            assert rel((ts.params - p0)[live], (tn.params - p0)[live]) <= 5e-2
            assert same_on_all_ranks(ts.params, world) and same_on_all_ranks(ts.grads, world)
        assert same_on_all_ranks(tn.params, world)
        if rank == 0:
            print(f"[dp] {dtype}: all-reduced grads vs single-process union rel err {e:.2e}; symmetric exchange "
                  f"{'checked' if has_symm else 'unavailable (NCCL path only)'}", flush=True)
        del ts, tn, tr, one
    # (3) the CT+pathology aggregator (FusionTrainer, collapsed program): DP over patients
    ns = Namespace(modality=["CT", "pathology"], model_CT="resnetMC3_18", model_pathology="ABMIL", model_CI="none",
                   aggregator="ABMIL", num_classes=2, alignment_base="none", clinical_features=list("abcdefghi"))
    torch.manual_seed(5)
    m = mil_b200.get_model(ns).to(dev).eval()
    Bp, Nc = 2, 160
    plen = [[700, 90], [1500, 333]]
    gg = torch.Generator().manual_seed(3)
    data = []
    for r in range(world if world <= 2 else 2):
        data.append((torch.randn(Bp, Nc, 512, generator=gg), torch.randn(sum(plen[r]), 768, generator=gg),
                     torch.randn(Bp, 512, generator=gg) * 0.05,
                     torch.tensor([[0.0, 1.0], [1.0, 0.0]]) if r == 0 else torch.tensor([[1.0, 0.0], [1.0, 0.0]])))
    r = rank % 2
    ft = mil_b200.FusionTrainer(m, n_text_tokens=1, compute_dtype=torch.float32, process_group=pg, world_size=world, lr=1e-3)
    ft.broadcast_params()
    ct, xp, xt, lab = (t.to(dev) for t in data[r])
    ft.forward_backward_bags(ct, xp, plen[r], xt, lab)
    dist.all_reduce(ft.grads)
    if world == 2:
        f1 = mil_b200.FusionTrainer(m, n_text_tokens=1, compute_dtype=torch.float32, lr=1e-3)
        f1.params.copy_(ft.params)
        f1._refresh_compute_copy()
        ctu = torch.cat([data[0][0], data[1][0]]).to(dev)
        xpu = torch.cat([data[0][1], data[1][1]]).to(dev)
        xtu = torch.cat([data[0][2], data[1][2]]).to(dev)
        labu = torch.cat([data[0][3], data[1][3]]).to(dev)
        f1.forward_backward_bags(ctu, xpu, plen[0] + plen[1], xtu, labu)
        # DDP averages: sum over ranks / world of per-rank means (2 patients each) == mean over the 4 patients
        ef = rel(ft.grads / world, f1.grads)
        assert ef <= 2e-5, ("fusion DP grads vs single-process union", ef)
        if rank == 0:
            print(f"[dp] FusionTrainer: averaged grads vs single-process union rel err {ef:.2e}", flush=True)
    ft.grads.div_(1.0)          # (the reduce above already ran; reduce_and_update would reduce twice)
    assert same_on_all_ranks(ft.grads, world)
    dist.barrier()
    if rank == 0:
        print("DP_NCCL_OK", flush=True)
    dist.destroy_process_group()


def module_flat(module, tr):
    ref = AbmilTrainer(tr.L, tr.D, tr.dtype, device=tr.device)
    ref.load_from(module)
    return ref.params


if __name__ == "__main__":
    try:
        main()
    except BaseException:
        import traceback
        print(f"[rank {os.environ.get('RANK')}] " + traceback.format_exc()[-1500:], flush=True)
        raise
