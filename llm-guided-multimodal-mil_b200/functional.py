"""autograd.Functions over the C ABI.  Every forward/backward here is a sequence of libmilb200 kernel
launches on the current CUDA stream; nothing is computed with eager PyTorch ops except dtype casts of
tiny parameter vectors and views/slices of the gradient buffers."""
from __future__ import annotations

import torch

from . import _lib as L

# --------------------------------------------------------------------------------------------------
# gated-attention MIL pool over a CSR batch of bags           (reference: model/dim1/ABMIL.py:47-64)
# --------------------------------------------------------------------------------------------------


def _f32(t):
    return t if t.dtype == torch.float32 else t.float()


def pack_gate_weights(Wv, bv, Wu, bu, dtype):
    """[Wv; Wu] -> Wcat (2D, L) in `dtype`, [bv; bu] -> bcat fp32 (one kernel)."""
    D, Lf = Wv.shape
    Wcat = torch.empty((2 * D, Lf), dtype=dtype, device=Wv.device)
    bcat = torch.empty((2 * D,), dtype=torch.float32, device=Wv.device)
    Wv, Wu, bv, bu = (t.contiguous() for t in (Wv, Wu, bv, bu))
    if not (Wv.dtype == Wu.dtype == bv.dtype == bu.dtype):
        raise L.MilB200Error("gate parameters must share one dtype")
    L.check(L.lib().milb200_pack_gate_weights(L.ptr(Wv), L.ptr(Wu), L.ptr(bv), L.ptr(bu), L.dtype_code(Wv),
                                              Lf, D, L.ptr(Wcat), L.dtype_code(Wcat), L.ptr(bcat),
                                              L.stream_ptr()), "pack_gate_weights")
    return Wcat, bcat


def gated_scores(X, Wcat, bcat, ww, bw):
    """s[i] = (tanh(x_i Wv^T + bv) * sigmoid(x_i Wu^T + bu)) . ww + bw   -> fp32 [total_n]."""
    n, Lf = X.shape
    D = Wcat.shape[0] // 2
    s = torch.empty((n,), dtype=torch.float32, device=X.device)
    code = L.dtype_code(X)
    nb = L.lib().milb200_gated_score_workspace_bytes(n, Lf, D, code, 0)
    ws = L.workspace(nb, X.device)
    L.check(L.lib().milb200_gated_score_fwd(L.ptr(X), L.ptr(Wcat), L.ptr(bcat), L.ptr(ww), L.ptr(bw), L.ptr(s),
                                            n, Lf, D, code, L.ptr(ws), ws.numel(), L.stream_ptr()),
            "gated_score_fwd")
    return s


def segment_softmax_pool(X, s, offsets, want_lowp=False):
    """M[b] = sum_i softmax_b(s)_i x_i over CSR offsets.  Returns (M fp32 [B,L], M_lowp|None, argmax int32 [B],
    lse fp32 [B])."""
    n, Lf = X.shape
    B = offsets.numel() - 1
    M = torch.empty((B, Lf), dtype=torch.float32, device=X.device)
    Ml = torch.empty((B, Lf), dtype=X.dtype, device=X.device) if want_lowp else None
    am = torch.empty((B,), dtype=torch.int32, device=X.device)
    lse = torch.empty((B,), dtype=torch.float32, device=X.device)
    nb = L.lib().milb200_pool_workspace_bytes(n, B, Lf)
    ws = L.workspace(nb, X.device)
    L.check(L.lib().milb200_segment_softmax_pool_fwd(L.ptr(X), L.ptr(s), L.ptr(offsets), B, n, Lf,
                                                     L.dtype_code(X), L.ptr(M), L.ptr(Ml), L.ptr(am), L.ptr(lse),
                                                     L.ptr(ws), ws.numel(), L.stream_ptr()),
            "segment_softmax_pool_fwd")
    return M, Ml, am, lse


def segment_softmax_pool_bwd(X, s, offsets, dM, M, want_attn):
    n, Lf = X.shape
    B = offsets.numel() - 1
    ds = torch.empty((n,), dtype=torch.float32, device=X.device)
    attn = torch.empty((n,), dtype=torch.float32, device=X.device) if want_attn else None
    nb = L.lib().milb200_pool_workspace_bytes(n, B, Lf)
    ws = L.workspace(nb, X.device)
    L.check(L.lib().milb200_segment_softmax_pool_bwd(L.ptr(X), L.ptr(s), L.ptr(offsets), B, n, Lf,
                                                     L.dtype_code(X), L.ptr(dM), L.ptr(M), L.ptr(ds), L.ptr(attn),
                                                     L.ptr(ws), ws.numel(), L.stream_ptr()),
            "segment_softmax_pool_bwd")
    return ds, attn


def gated_scores_bwd(X, Wcat, bcat, ww, bw, ds, attn, dM, offsets, need_dx, grad_out=None):
    """Backward of gated_scores (+ the pooling term of dX).  `grad_out`, if given, is a flat fp32 buffer of
    2D*L + 2D + D + 1 elements (laid out dWcat | dbcat | dww | dbw) that the kernels write in place — the
    data-parallel trainer passes a slice of its flat gradient buffer."""
    n, Lf = X.shape
    D = Wcat.shape[0] // 2
    B = offsets.numel() - 1
    need = 2 * D * Lf + 2 * D + D + 1
    if grad_out is None:
        grad_out = torch.empty((need,), dtype=torch.float32, device=X.device)
    elif grad_out.numel() != need or grad_out.dtype != torch.float32:
        raise L.MilB200Error("gated_scores_bwd: grad_out must be fp32 with 2D*L+3D+1 elements")
    dWcat = grad_out[:2 * D * Lf].view(2 * D, Lf)
    dbcat = grad_out[2 * D * Lf:2 * D * Lf + 2 * D]
    dww = grad_out[2 * D * Lf + 2 * D:2 * D * Lf + 3 * D]
    dbw = grad_out[2 * D * Lf + 3 * D:]
    dX = torch.empty_like(X) if need_dx else None
    code = L.dtype_code(X)
    nb = L.lib().milb200_gated_score_workspace_bytes(n, Lf, D, code, 1)
    ws = L.workspace(nb, X.device)
    L.check(L.lib().milb200_gated_score_bwd(L.ptr(X), L.ptr(Wcat), L.ptr(bcat), L.ptr(ww), L.ptr(bw), L.ptr(ds),
                                            L.ptr(attn) if need_dx else None, L.ptr(dM) if need_dx else None,
                                            L.ptr(offsets), B, n, Lf, D, code, L.ptr(dWcat), L.ptr(dbcat),
                                            L.ptr(dww), L.ptr(dbw), L.ptr(dX), L.ptr(ws), ws.numel(),
                                            L.stream_ptr()), "gated_score_bwd")
    return dX, dWcat, dbcat, dww, dbw


class _AbmilPoolCSR(torch.autograd.Function):
    """M[b] = ABMIL(x[offsets[b]:offsets[b+1]]) for every bag of a packed batch, eval-mode semantics
    (ABMIL.py:47-64 applied bag by bag, which is how train_ddp.py:75 / test_ddp.py:73 run it)."""

    @staticmethod
    def forward(ctx, X, offsets, Wv, bv, Wu, bu, ww, bw):
        X = X.contiguous()
        Wcat, bcat = pack_gate_weights(Wv, bv, Wu, bu, X.dtype)
        wwf = _f32(ww).reshape(-1).contiguous()
        bwf = _f32(bw).reshape(-1).contiguous()
        s = gated_scores(X, Wcat, bcat, wwf, bwf)
        M, Ml, am, lse = segment_softmax_pool(X, s, offsets, want_lowp=X.dtype != torch.float32)
        ctx.save_for_backward(X, offsets, Wcat, bcat, wwf, bwf, s, M)
        ctx.param_dtypes = (Wv.dtype, bv.dtype, Wu.dtype, bu.dtype, ww.dtype, bw.dtype)
        ctx.D = Wv.shape[0]
        ctx.mark_non_differentiable(am, s)
        return (M if Ml is None else Ml), am, s

    @staticmethod
    def backward(ctx, dM_out, _dam, _ds):
        X, offsets, Wcat, bcat, wwf, bwf, s, M = ctx.saved_tensors
        D = ctx.D
        need_dx = ctx.needs_input_grad[0]
        dM = _f32(dM_out).contiguous()
        ds, attn = segment_softmax_pool_bwd(X, s, offsets, dM, M, want_attn=need_dx)
        dX, dWcat, dbcat, dww, dbw = gated_scores_bwd(X, Wcat, bcat, wwf, bwf, ds, attn, dM, offsets, need_dx)
        dt = ctx.param_dtypes
        return (dX, None, dWcat[:D].to(dt[0]), dbcat[:D].to(dt[1]), dWcat[D:].to(dt[2]), dbcat[D:].to(dt[3]),
                dww.view(1, D).to(dt[4]), dbw.view(1).to(dt[5]))


def abmil_pool_csr(X, offsets, Wv, bv, Wu, bu, ww, bw):
    """Returns (M [B,L] in X.dtype, argmax int32 [B] (index within the bag), scores fp32 [total_n])."""
    if offsets.dtype != torch.int32:
        raise L.MilB200Error("offsets must be int32 CSR offsets on the device")
    return _AbmilPoolCSR.apply(X, offsets, Wv, bv, Wu, bu, ww, bw)


# --------------------------------------------------------------------------------------------------
# dropout (train mode only) and the dense-batch sum pool quirk
# --------------------------------------------------------------------------------------------------
def _dropout_raw(x, p, seed, offset):
    out = torch.empty_like(x)
    L.check(L.lib().milb200_dropout(L.ptr(x), L.ptr(out), x.numel(), float(p), int(seed), int(offset),
                                    L.dtype_code(x), L.stream_ptr()), "dropout")
    return out


class _Dropout(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, p):
        x = x.contiguous()
        # one 63-bit seed per call from torch's CPU generator (so torch.manual_seed controls it)
        ctx.seed = int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item())
        ctx.p = p
        return _dropout_raw(x, p, ctx.seed, 0)

    @staticmethod
    def backward(ctx, g):
        return _dropout_raw(g.contiguous(), ctx.p, ctx.seed, 0), None


def dropout(x, p):
    """nn.Dropout(p) in train mode; Philox mask regenerated (not stored) in backward."""
    if p <= 0.0:
        return x
    return _Dropout.apply(x, p)


class _DenseSumPool(torch.autograd.Function):
    """x (B,N,L) -> (B,1,L) = sum over N: the reference's behaviour for dense batches B>1."""

    @staticmethod
    def forward(ctx, x):
        B, N, Lf = x.shape
        X = x.contiguous().view(B * N, Lf)
        off = torch.arange(0, (B + 1) * N, N, dtype=torch.int32, device=x.device)
        M = torch.empty((B, Lf), dtype=torch.float32, device=x.device)
        Ml = torch.empty((B, Lf), dtype=x.dtype, device=x.device) if x.dtype != torch.float32 else None
        nb = L.lib().milb200_pool_workspace_bytes(B * N, B, Lf)
        ws = L.workspace(nb, x.device)
        L.check(L.lib().milb200_segment_sum_fwd(L.ptr(X), L.ptr(off), B, B * N, Lf, L.dtype_code(X), L.ptr(M),
                                                L.ptr(Ml), L.ptr(ws), ws.numel(), L.stream_ptr()),
                "segment_sum_fwd")
        ctx.save_for_backward(off)
        ctx.shape = (B, N, Lf)
        ctx.dtype = x.dtype
        return (M if Ml is None else Ml).view(B, 1, Lf)

    @staticmethod
    def backward(ctx, g):
        (off,) = ctx.saved_tensors
        B, N, Lf = ctx.shape
        gf = _f32(g).contiguous().view(B, Lf)
        out = torch.empty((B * N, Lf), dtype=ctx.dtype, device=g.device)
        L.check(L.lib().milb200_bag_broadcast(L.ptr(gf), None, L.ptr(off), B, B * N, Lf, L.dtype_code(out),
                                              L.ptr(out), L.stream_ptr()), "bag_broadcast")
        return out.view(B, N, Lf)


def dense_sum_pool(x):
    return _DenseSumPool.apply(x)
