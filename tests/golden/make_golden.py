#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference modules
from /root/reference on CPU (build container only — the GPU box has no /root/reference and never
runs this).  Re-run:  python tests/golden/make_golden.py

Everything (weights, inputs, upstream gradients) is generated procedurally with
numpy.random.RandomState so the fixtures only need to hold the *outputs*; the tests regenerate the
same weights/inputs with oracle.mil_oracle.procedural_state / the helpers below.

Import recipe: three sys.modules stubs for un-vendored deps
(`nystrom_attention`, `clip`, `model.dim3`), a `Tensor.cuda` identity shim (the reference hard-codes
.cuda(), aggregator.py:160,168) and `model.dim1.gatedAttention = ABMIL` (the name
aggregator_wMask.py:24 imports does not exist upstream).
"""
from __future__ import annotations

import ast
import os
import sys
import types
from argparse import Namespace

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
from oracle.mil_oracle import procedural_state  # noqa: E402


# ------------------------------------------------------------------ reference import under stubs
def import_reference():
    if not os.path.isdir(REF):
        raise SystemExit("make_golden.py needs /root/reference (build container only)")
    sys.path.insert(0, REF)
    nys = types.ModuleType("nystrom_attention")

    class NystromAttention(nn.Module):  # never executed (TransMIL is off-path)
        def __init__(self, *a, **k):
            super().__init__()

    nys.NystromAttention = NystromAttention
    sys.modules["nystrom_attention"] = nys
    sys.modules["clip"] = types.ModuleType("clip")

    dim3 = types.ModuleType("model.dim3")

    class _PassThroughEncoder(nn.Module):
        """Stands in for the CT encoders (out of scope): returns its input, which the fixtures
        supply already in the encoder's *output* layout."""
        def __init__(self, args, weights=None, progress=None):
            super().__init__()

        def forward(self, x, mask=None):
            return x

    for n in ("Resnet2plus1D_18", "ResnetMC3_18", "ResnetMC3_18_wMask", "ResNeXt", "medicalNet",
              "SwinUNETR", "SwinUNETR_wMask", "MViT_v2"):
        setattr(dim3, n, _PassThroughEncoder)
    sys.modules["model.dim3"] = dim3
    torch.Tensor.cuda = lambda self, *a, **k: self

    import model.dim1 as dim1
    dim1.gatedAttention = dim1.ABMIL
    import model.sam.transformer as tr
    import model.aggregator as agg
    import model.aggregator_clip as agg_clip
    import model.aggregator_wMask as agg_wmask
    return dim1, tr, agg, agg_clip, agg_wmask


def load_state(module, seed):
    shapes = {k: tuple(v.shape) for k, v in module.state_dict().items()}
    sd = procedural_state(shapes, seed)
    module.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    return sd


def rnd(seed, *shape, scale=1.0):
    return (np.random.RandomState(seed).standard_normal(shape) * scale).astype(np.float32)


def grads_of(module):
    return {k: (p.grad.detach().numpy().copy() if p.grad is not None else None)
            for k, p in module.named_parameters()}


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"wrote {path}  ({os.path.getsize(path)/1024:.1f} KiB)")


def pack_grads(g, full_limit=60_000):
    """Full gradient arrays when small, [sum, sumsq, first 16] digests ('d:' prefix) when large, so
    fixtures stay small."""
    out = {}
    for k, v in g.items():
        if v is None:
            out["g:" + k] = np.zeros(0)
        elif v.size <= full_limit:
            out["g:" + k] = v
        else:
            out["d:" + k] = grad_digest(v)
    return out


def grad_digest(g):
    """Compact digest of a big gradient: [sum, sum of squares, first 16 values (flattened)]."""
    f = g.astype(np.float64).reshape(-1)
    head = np.zeros(16); head[:min(16, f.size)] = f[:16]
    return np.concatenate([[f.sum(), (f * f).sum()], head])


# ------------------------------------------------------------------ fixtures
def fx_abmil(dim1):
    """ABMIL fwd+bwd at several (L,N); loss = sum(M * dM)."""
    cases = [("abmil_L96_N37", 96, 37, 11), ("abmil_L768_N100", 768, 100, 12),
             ("abmil_L1024_N257", 1024, 257, 13), ("abmil_L512_N300", 512, 300, 14),
             ("abmil_L1024_N1", 1024, 1, 15)]
    for name, L, N, seed in cases:
        m = dim1.ABMIL(None, L=L).double().eval()
        load_state(m, seed)
        x = torch.from_numpy(rnd(seed + 100, 1, N, L)).double().requires_grad_(True)
        dM = torch.from_numpy(rnd(seed + 200, 1, L)).double()
        M = m(x)
        (M * dM).sum().backward()
        with torch.no_grad():
            xs = x.squeeze(0)
            s = m.attention_weights(m.attention_V(xs) * m.attention_U(xs)).reshape(-1)
        g = grads_of(m)
        save(name, L=L, N=N, seed=seed, M=M.detach().numpy(), s=s.numpy(),
             argmax=int(torch.argmax(s)), dx=x.grad.numpy().reshape(N, L).astype(np.float64) if N * L <= 60_000
             else grad_digest(x.grad.numpy()), **pack_grads(g))


def fx_abmil_quirks(dim1):
    m = dim1.ABMIL(None, L=96).double().eval()
    load_state(m, 21)
    x = torch.from_numpy(rnd(121, 3, 17, 96)).double()
    save("abmil_dense_batched_B3", L=96, N=17, B=3, seed=21, M=m(x).detach().numpy())
    m2 = dim1.ABMIL_v2(None).double().eval()
    load_state(m2, 22)
    x = torch.from_numpy(rnd(122, 1, 29, 768)).double()
    cls = torch.tensor([[1.0]], dtype=torch.float64)
    save("abmil_v2_N29", N=29, seed=22, M=m2(x, cls).detach().numpy())


def fx_attention(tr):
    for name, ds, seed in [("attention_ds1", 1, 31), ("attention_ds2", 2, 32)]:
        m = tr.Attention(64, 8, downsample_rate=ds).double()
        load_state(m, seed)
        q = torch.from_numpy(rnd(seed + 100, 1, 5, 64)).double().requires_grad_(True)
        k = torch.from_numpy(rnd(seed + 200, 1, 33, 64)).double().requires_grad_(True)
        v = torch.from_numpy(rnd(seed + 300, 1, 33, 64)).double().requires_grad_(True)
        do = torch.from_numpy(rnd(seed + 400, 1, 5, 64)).double()
        out = m(q, k, v)
        (out * do).sum().backward()
        save(name, seed=seed, ds=ds, out=out.detach().numpy(), dq=q.grad.numpy(), dk=k.grad.numpy(),
             dv=v.grad.numpy(), **pack_grads(grads_of(m)))


def fx_block(tr):
    for name, skip, T, seed in [("block_skip_pe_T3", True, 3, 41), ("block_pe_T1", False, 1, 42),
                                ("block_pe_T10", False, 10, 43)]:
        m = tr.TwoWayAttentionBlock(64, 8, mlp_dim=128, skip_first_layer_pe=skip).double()
        load_state(m, seed)
        N = 41
        qs = torch.from_numpy(rnd(seed + 100, 1, T, 64)).double().requires_grad_(True)
        ks = torch.from_numpy(rnd(seed + 200, 1, N, 64)).double().requires_grad_(True)
        qpe = torch.from_numpy(rnd(seed + 300, 1, T, 64)).double()
        kpe = torch.from_numpy(rnd(seed + 400, 1, N, 64)).double()
        dq = torch.from_numpy(rnd(seed + 500, 1, T, 64)).double()
        dk = torch.from_numpy(rnd(seed + 600, 1, N, 64)).double()
        oq, ok = m(qs, ks, qpe, kpe)
        ((oq * dq).sum() + (ok * dk).sum()).backward()
        save(name, seed=seed, skip=int(skip), T=T, N=N, oq=oq.detach().numpy(), ok=ok.detach().numpy(),
             dqs=qs.grad.numpy(), dks=ks.grad.numpy(),
             **pack_grads(grads_of(m)))


def fx_transformer(tr):
    args = Namespace(alignment_base="none", model_CT="resnetMC3_18")
    for name, T, N, five_d, seed in [("twoway_T1_N50", 1, 50, False, 51),
                                     ("twoway_T10_N50", 10, 50, False, 52),
                                     ("twoway_ct5d_T1", 1, 12, True, 53)]:
        m = tr.TwoWayTransformer(args=args, depth=2, embedding_dim=64, num_heads=8, mlp_dim=128).double()
        load_state(m, seed)
        if five_d:
            img = torch.from_numpy(rnd(seed + 100, 1, 64, N, 2, 3)).double().requires_grad_(True)
        else:
            img = torch.from_numpy(rnd(seed + 100, 1, N, 64)).double().requires_grad_(True)
        pe = torch.from_numpy(rnd(seed + 200, 1, N, 64)).double()
        pt = torch.from_numpy(rnd(seed + 300, 1, T, 64)).double().requires_grad_(True)
        dq = torch.from_numpy(rnd(seed + 500, 1, T, 64)).double()
        dk = torch.from_numpy(rnd(seed + 600, 1, N, 64)).double()
        oq, ok = m(img, pe, pt)
        ((oq * dq).sum() + (ok * dk).sum()).backward()
        save(name, seed=seed, T=T, N=N, five_d=int(five_d), oq=oq.detach().numpy(),
             ok=ok.detach().numpy(), dimg=img.grad.numpy(), dpt=pt.grad.numpy(),
             **pack_grads(grads_of(m)))


class _TextStub(nn.Module):
    """clinic_extractor stand-in: returns the synthetic (1,T,512) text embedding it is given
    (the CLIP text tower is a frozen upstream encoder, out of scope)."""
    def forward(self, x):
        return x


def fx_aggregator(agg):
    for name, T, N, seed in [("aggregator_T1_N70", 1, 70, 61), ("aggregator_T10_N45", 10, 45, 62)]:
        args = Namespace(modality=["CT", "pathology"], model_CT="resnetMC3_18", model_pathology="ABMIL",
                         model_CI="none", aggregator="ABMIL", num_classes=2, alignment_base="none",
                         clinical_features=list("abcdefghi"))
        m = agg.aggregator(args)
        m.clinic_extractor = _TextStub()
        m = m.float().eval()
        load_state(m, seed)
        x_ct = torch.from_numpy(rnd(seed + 100, 1, 512, 160, 1, 2)).requires_grad_(True)
        x_p = torch.from_numpy(rnd(seed + 200, 1, N, 768)).requires_grad_(True)
        x_t = torch.from_numpy(rnd(seed + 300, 1, T, 512, scale=0.05))
        label = torch.tensor([[0.0, 1.0]])
        prob, ct2ci, pth2ci = m([x_ct, x_p], x_t)
        loss = nn.BCELoss()(prob, label) + nn.CosineEmbeddingLoss()(
            ct2ci.squeeze(0), pth2ci.squeeze(0), torch.ones(T))
        loss.backward()
        g = grads_of(m)
        dig = {"g:" + k: grad_digest(v) for k, v in g.items() if v is not None}
        dead = np.array([k for k, v in g.items() if v is None])
        save(name, seed=seed, T=T, N=N, prob=prob.detach().numpy(), ct2ci=ct2ci.detach().numpy(),
             pth2ci=pth2ci.detach().numpy(), loss=float(loss), dx_p=grad_digest(x_p.grad.numpy()),
             dx_ct=grad_digest(x_ct.grad.numpy()), dead=dead, **dig)


def fx_aggregator_clip(agg_clip):
    args = Namespace(modality=["CT", "pathology"], model_CT="resnetMC3_18", model_pathology="ABMIL",
                     num_classes=2)
    m = agg_clip.aggregator(args).double().eval()
    load_state(m, 71)
    x_ct = torch.from_numpy(rnd(171, 1, 512)).double()
    x_p = torch.from_numpy(rnd(271, 1, 83, 768)).double().requires_grad_(True)
    a, b, prob = m([x_ct, x_p])
    ((a * b).sum() + prob[0, 1]).backward()
    save("aggregator_clip_ctpath", seed=71, N=83, x_ct=a.detach().numpy(), x_path=b.detach().numpy(),
         prob=prob.detach().numpy(), dx_p=x_p.grad.numpy(),
         **pack_grads(grads_of(m)))
    args = Namespace(modality=["pathology"], model_pathology="ABMIL", num_classes=2)
    m = agg_clip.aggregator(args).double().eval()
    load_state(m, 72)
    x_p = torch.from_numpy(rnd(272, 1, 64, 768)).double()
    pooled, prob = m([x_p])
    save("aggregator_clip_path", seed=72, N=64, pooled=pooled.detach().numpy(), prob=prob.detach().numpy())


def fx_aggregator_wmask(agg_wmask):
    args = Namespace(modality=["CT", "pathology"], model_CT="resnetMC3_18_wMask", model_pathology="ABMIL",
                     num_classes=2, clinical_features=list("abcdefghi"))
    m = agg_wmask.aggregator_wMask(args).double().eval()
    load_state(m, 81)
    # the stub CT encoder returns cat([x, mask], dim=1) unchanged -> supply 384+384 so x_CT is (1,768)
    x_ct = torch.from_numpy(rnd(181, 1, 384)).double()
    mask = torch.from_numpy(rnd(182, 1, 384)).double()
    x_p = torch.from_numpy(rnd(281, 1, 57, 768)).double()
    prob = m([x_ct, x_p], mask)
    save("aggregator_wmask_ctpath", seed=81, N=57, prob=prob.detach().numpy())


def fx_clip_logits():
    """clip/model.py:354-368 executed unmodified via CLIP.forward on a duck-typed self
    (the towers are out of scope; encode_* return the given features)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_clip_model", os.path.join(REF, "clip", "model.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    img = torch.from_numpy(rnd(91, 24, 512)).double().requires_grad_(True)
    txt = torch.from_numpy(rnd(92, 24, 512)).double().requires_grad_(True)
    fake = types.SimpleNamespace(encode_image=lambda x: x, encode_text=lambda x: x,
                                 logit_scale=torch.tensor(np.log(1 / 0.07), dtype=torch.float64,
                                                          requires_grad=True))
    li, lt = mod.CLIP.forward(fake, img, txt)
    dli = torch.from_numpy(rnd(93, 24, 24)).double()
    dlt = torch.from_numpy(rnd(94, 24, 24)).double()
    ((li * dli).sum() + (lt * dlt).sum()).backward()
    save("clip_logits_b24", li=li.detach().numpy(), lt=lt.detach().numpy(), dimg=img.grad.numpy(),
         dtxt=txt.grad.numpy(), dscale=float(fake.logit_scale.grad))


def fx_cliploss():
    """utils.py:247-284 `CLIPloss_v1.forward` executed unmodified: the class source is lifted from
    the reference file at run time (utils.py itself imports SimpleITK etc., absent here) and run with
    a stub `clip` (tokenize = identity on indices, encode_text = table lookup of synthetic features)."""
    src = open(os.path.join(REF, "utils.py")).read()
    tree = ast.parse(src)
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "CLIPloss_v1")
    b, I = 6, 9
    feats = rnd(95, b, I, 512, scale=0.3)
    state = {"b": -1}

    clip_stub = types.SimpleNamespace()

    def tokenize(texts):
        state["b"] += 1
        return torch.full((len(texts), 1), state["b"])
    clip_stub.tokenize = tokenize
    ns = {"torch": torch, "clip": clip_stub}
    exec(compile(ast.Module(body=[cls], type_ignores=[]), "ref_utils_CLIPloss_v1", "exec"), ns)
    model = types.SimpleNamespace(encode_text=lambda tok: torch.from_numpy(feats[int(tok[0, 0])]))
    fake = types.SimpleNamespace(args=Namespace(gpu=None), clinical_info=list("abcdefghi"), model=model,
                                 criterion=torch.nn.CrossEntropyLoss())
    out = torch.from_numpy(rnd(96, b, 512, scale=0.3)).requires_grad_(True)
    ci = torch.zeros(b, 1, I)
    loss = ns["CLIPloss_v1"].forward(fake, out, ci)
    loss.backward()
    save("cliploss_v1_b6", loss=float(loss), dout=out.grad.numpy())


def main():
    torch.manual_seed(1234)
    torch.set_num_threads(4)
    dim1, tr, agg, agg_clip, agg_wmask = import_reference()
    fx_abmil(dim1)
    fx_abmil_quirks(dim1)
    fx_attention(tr)
    fx_block(tr)
    fx_transformer(tr)
    fx_aggregator(agg)
    fx_aggregator_clip(agg_clip)
    fx_aggregator_wmask(agg_wmask)
    fx_clip_logits()
    fx_cliploss()


if __name__ == "__main__":
    main()
