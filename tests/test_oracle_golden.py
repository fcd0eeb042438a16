"""CPU tests: the oracle (oracle/*.py) must reproduce every golden fixture that the unmodified
reference produced (tests/golden/make_golden.py).  This is what pins the oracle."""
import math

import numpy as np
import pytest
import torch

from oracle import fusion_oracle as fo
from oracle import mil_oracle as mo
from tests.helpers import check_grads, digest, load_golden, rel_err, rnd

TOL64 = 1e-9      # float64 oracle vs float64 reference
TOL32 = 2e-4      # float32 reference run (full aggregator fixture)


def _abmil_params(L, seed):
    return mo.procedural_state(mo.abmil_shapes(L), seed)


@pytest.mark.parametrize("name", ["abmil_L96_N37", "abmil_L768_N100", "abmil_L1024_N257",
                                  "abmil_L512_N300", "abmil_L1024_N1"])
def test_abmil_numpy_vs_reference(name):
    fx = load_golden(name)
    L, N, seed = int(fx["L"]), int(fx["N"]), int(fx["seed"])
    p = _abmil_params(L, seed)
    x = rnd(seed + 100, 1, N, L)[0]
    dM = rnd(seed + 200, 1, L)
    f = mo.abmil_forward(p, x)
    assert rel_err(f["M"], fx["M"]) < TOL64
    assert rel_err(f["s"], fx["s"]) < TOL64
    assert f["argmax"] == int(fx["argmax"])          # bit-exact index
    g = mo.abmil_backward(p, x, dM)
    if fx["dx"].shape == (N, L):
        assert rel_err(g["x"], fx["dx"]) < TOL64
    else:
        assert np.allclose(digest(g["x"]), fx["dx"], rtol=1e-8, atol=1e-10)
    # attention_weights.bias has zero true gradient (softmax shift invariance): absolute check
    assert abs(g["attention_weights.bias"][0]) < 1e-12
    n = check_grads(fx, g, TOL64, prefix_filter=lambda k: k != "attention_weights.bias")
    assert n == 5


def test_abmil_torch_oracle_matches_numpy_oracle():
    p = _abmil_params(96, 5)
    x = rnd(6, 23, 96)
    sd = {"a." + k: v for k, v in fo.to_torch(p).items()}
    M = fo.abmil(sd, "a", torch.from_numpy(x).double()[None])
    assert rel_err(M.numpy(), mo.abmil_forward(p, x)["M"]) < 1e-12


def test_abmil_csr_equals_loop_and_backward_sums():
    L = 96
    p = _abmil_params(L, 7)
    lens = mo.ragged_lengths(5, 1, 40, 3)
    off = mo.offsets_from_lengths(lens)
    X = rnd(8, int(off[-1]), L)
    dM = rnd(9, 5, L)
    M, s, am = mo.abmil_forward_csr(p, X, off)
    g = mo.abmil_backward_csr(p, X, off, dM)
    # autograd cross-check of the analytic backward on the summed loss
    sd = fo.to_torch({"a." + k: v for k, v in p.items()}, requires_grad=True)
    Xt = torch.from_numpy(X).double().requires_grad_(True)
    loss = 0
    for b in range(5):
        Mb = fo.abmil(sd, "a", Xt[off[b]:off[b + 1]][None])
        assert rel_err(Mb.detach().numpy()[0], M[b]) < 1e-12
        loss = loss + (Mb * torch.from_numpy(dM[b]).double()).sum()
    loss.backward()
    assert rel_err(g["x"], Xt.grad.numpy()) < 1e-10
    for k in p:
        if k == "attention_weights.bias":
            continue
        assert rel_err(g[k], sd["a." + k].grad.numpy()) < 1e-10, k


def test_abmil_quirks():
    fx = load_golden("abmil_dense_batched_B3")
    x = rnd(121, 3, 17, 96)
    assert rel_err(mo.abmil_dense_batched(None, x), fx["M"]) < 1e-12          # F2: plain sum pool
    p = _abmil_params(96, 21)
    sd = {"a." + k: v for k, v in fo.to_torch(p).items()}
    assert rel_err(fo.abmil(sd, "a", torch.from_numpy(x).double()).numpy(), fx["M"]) < 1e-12
    fx = load_golden("abmil_v2_N29")
    p = _abmil_params(768, 22)
    f = mo.abmil_v2_forward(p, rnd(122, 1, 29, 768)[0], np.array([[1.0]]))
    assert f["M"].shape == (1, 769)
    assert rel_err(f["M"], fx["M"]) < TOL64


def test_masked_equals_unpadded():
    p = _abmil_params(96, 33)
    lens = [5, 17, 1]
    Xp = rnd(34, 3, 17, 96)
    M = mo.abmil_forward_masked(p, Xp, lens)
    for b, n in enumerate(lens):
        assert rel_err(M[b], mo.abmil_forward(p, Xp[b, :n])["M"][0]) < 1e-14


def _attn_shapes(E, ds):
    I = E // ds
    return {"q_proj.weight": (I, E), "q_proj.bias": (I,), "k_proj.weight": (I, E), "k_proj.bias": (I,),
            "v_proj.weight": (I, E), "v_proj.bias": (I,), "out_proj.weight": (E, I), "out_proj.bias": (E,)}


def block_shapes(E, mlp, ds=2, prefix=""):
    sh = {}
    for n, d in (("self_attn", 1), ("cross_attn_token_to_image", ds), ("cross_attn_image_to_token", ds)):
        for k, v in _attn_shapes(E, d).items():
            sh[f"{prefix}{n}.{k}"] = v
    for n in ("norm1", "norm2", "norm3", "norm4"):
        sh[f"{prefix}{n}.weight"] = (E,)
        sh[f"{prefix}{n}.bias"] = (E,)
    sh[f"{prefix}mlp.lin1.weight"] = (mlp, E); sh[f"{prefix}mlp.lin1.bias"] = (mlp,)
    sh[f"{prefix}mlp.lin2.weight"] = (E, mlp); sh[f"{prefix}mlp.lin2.bias"] = (E,)
    return sh


def transformer_shapes(E, mlp, depth=2, ds=2, prefix=""):
    sh = {}
    for i in range(depth):
        sh.update(block_shapes(E, mlp, ds, f"{prefix}layers.{i}."))
    for k, v in _attn_shapes(E, ds).items():
        sh[f"{prefix}final_attn_token_to_image.{k}"] = v
    sh[f"{prefix}norm_final_attn.weight"] = (E,)
    sh[f"{prefix}norm_final_attn.bias"] = (E,)
    return sh


def _t(a, grad=False):
    t = torch.from_numpy(np.asarray(a)).double()
    return t.requires_grad_(True) if grad else t


@pytest.mark.parametrize("name", ["attention_ds1", "attention_ds2"])
def test_attention_vs_reference(name):
    fx = load_golden(name)
    seed, ds = int(fx["seed"]), int(fx["ds"])
    sd = fo.to_torch({"m." + k: v for k, v in mo.procedural_state(_attn_shapes(64, ds), seed).items()},
                     requires_grad=True)
    q, k, v = (_t(rnd(seed + o, 1, n, 64), True) for o, n in ((100, 5), (200, 33), (300, 33)))
    out = fo.attention(sd, "m", q, k, v, 8)
    assert rel_err(out.detach().numpy(), fx["out"]) < TOL64
    (out * _t(rnd(seed + 400, 1, 5, 64))).sum().backward()
    for nm, t in (("dq", q), ("dk", k), ("dv", v)):
        assert rel_err(t.grad.numpy(), fx[nm]) < TOL64
    assert check_grads(fx, {k_[2:]: v_.grad.numpy() for k_, v_ in sd.items()}, TOL64) == 7


@pytest.mark.parametrize("name", ["block_skip_pe_T3", "block_pe_T1", "block_pe_T10"])
def test_block_vs_reference(name):
    fx = load_golden(name)
    seed, T, N, skip = int(fx["seed"]), int(fx["T"]), int(fx["N"]), bool(fx["skip"])
    sd = fo.to_torch(mo.procedural_state(block_shapes(64, 128, prefix="m."), seed), requires_grad=True)
    qs, ks = _t(rnd(seed + 100, 1, T, 64), True), _t(rnd(seed + 200, 1, N, 64), True)
    qpe, kpe = _t(rnd(seed + 300, 1, T, 64)), _t(rnd(seed + 400, 1, N, 64))
    oq, ok = fo.two_way_block(sd, "m", qs, ks, qpe, kpe, 8, skip)
    assert rel_err(oq.detach().numpy(), fx["oq"]) < TOL64
    assert rel_err(ok.detach().numpy(), fx["ok"]) < TOL64
    ((oq * _t(rnd(seed + 500, 1, T, 64))).sum() + (ok * _t(rnd(seed + 600, 1, N, 64))).sum()).backward()
    assert rel_err(qs.grad.numpy(), fx["dqs"]) < TOL64
    assert rel_err(ks.grad.numpy(), fx["dks"]) < TOL64
    grads = {k[2:]: (v.grad.numpy() if v.grad is not None else None) for k, v in sd.items()}
    if T == 1:   # F10: softmax over one key -> exactly zero grads for i2t q/k projections
        for nm in ("q_proj", "k_proj"):
            g = grads[f"cross_attn_image_to_token.{nm}.weight"]
            assert g is None or np.abs(g).max() == 0.0
    assert check_grads(fx, grads, 1e-8) > 30


@pytest.mark.parametrize("name", ["twoway_T1_N50", "twoway_T10_N50", "twoway_ct5d_T1"])
def test_transformer_vs_reference(name):
    fx = load_golden(name)
    seed, T, N, five_d = int(fx["seed"]), int(fx["T"]), int(fx["N"]), bool(fx["five_d"])
    sd = fo.to_torch(mo.procedural_state(transformer_shapes(64, 128, prefix="m."), seed), requires_grad=True)
    img = _t(rnd(seed + 100, 1, 64, N, 2, 3) if five_d else rnd(seed + 100, 1, N, 64), True)
    pe, pt = _t(rnd(seed + 200, 1, N, 64)), _t(rnd(seed + 300, 1, T, 64), True)
    oq, ok = fo.two_way_transformer(sd, "m", img, pe, pt, depth=2, num_heads=8)
    assert rel_err(oq.detach().numpy(), fx["oq"]) < TOL64
    assert rel_err(ok.detach().numpy(), fx["ok"]) < TOL64
    ((oq * _t(rnd(seed + 500, 1, T, 64))).sum() + (ok * _t(rnd(seed + 600, 1, N, 64))).sum()).backward()
    assert rel_err(img.grad.numpy(), fx["dimg"]) < 1e-8
    assert rel_err(pt.grad.numpy(), fx["dpt"]) < 1e-8
    grads = {k[2:]: (v.grad.numpy() if v.grad is not None else None) for k, v in sd.items()}
    assert check_grads(fx, grads, 1e-8) > 60


def aggregator_shapes(num_classes=2):
    """state_dict layout of model/aggregator.py for modality=[CT,pathology], model_pathology=ABMIL,
    aggregator=ABMIL, with the CT extractor / clinic_extractor parameter-free (stubs)."""
    E = 512
    sh = {}
    for tw in ("TwoWayTransformer_CT", "TwoWayTransformer_Pth", "TwoWayTransformer_Both"):
        sh.update(transformer_shapes(E, 2048, prefix=tw + "."))
    for fc in ("fc_CI2CT.0", "fc_CI2Pth.0", "fc_CI.0"):
        sh[fc + ".weight"] = (E, E); sh[fc + ".bias"] = (E,)
    sh["fc_pathology.0.weight"] = (E, 768); sh["fc_pathology.0.bias"] = (E,)
    for ab in ("extractor_pathology", "aggregator"):
        for k, v in mo.abmil_shapes(E).items():
            sh[f"{ab}.{k}"] = v
    sh["prompt_embedding"] = (1, E)
    sh["fc.1.weight"] = (num_classes, E); sh["fc.1.bias"] = (num_classes,)
    return sh


@pytest.mark.parametrize("name", ["aggregator_T1_N70", "aggregator_T10_N45"])
def test_aggregator_vs_reference(name):
    fx = load_golden(name)
    seed, T, N = int(fx["seed"]), int(fx["T"]), int(fx["N"])
    sd = fo.to_torch(mo.procedural_state(aggregator_shapes(), seed), dtype=torch.float64, requires_grad=True)
    x_ct = _t(rnd(seed + 100, 1, 512, 160, 1, 2), True)
    x_p = _t(rnd(seed + 200, 1, N, 768), True)
    x_t = _t(rnd(seed + 300, 1, T, 512, scale=0.05))
    prob, ct2ci, pth2ci = fo.aggregator_fusion_forward(sd, x_ct, x_p, x_t)
    assert rel_err(prob.detach().numpy(), fx["prob"]) < TOL32
    assert rel_err(ct2ci.detach().numpy(), fx["ct2ci"]) < TOL32
    assert rel_err(pth2ci.detach().numpy(), fx["pth2ci"]) < TOL32
    label = torch.tensor([[0.0, 1.0]], dtype=torch.float64)
    loss = torch.nn.BCELoss()(prob, label) + torch.nn.CosineEmbeddingLoss()(
        ct2ci.squeeze(0), pth2ci.squeeze(0), torch.ones(T, dtype=torch.float64))
    assert abs(float(loss) - float(fx["loss"])) < TOL32 * max(1.0, abs(float(fx["loss"])))
    # numpy loss helpers agree with torch's
    l_np = mo.cosine_embedding_loss_pos(ct2ci.detach().numpy()[0], pth2ci.detach().numpy()[0])
    assert abs(l_np - float(torch.nn.CosineEmbeddingLoss()(ct2ci.squeeze(0), pth2ci.squeeze(0),
                                                           torch.ones(T, dtype=torch.float64)))) < 1e-9
    loss.backward()
    dead = set(str(k) for k in fx["dead"])
    live = 0
    for key, ref in fx.items():
        if not key.startswith("g:"):
            continue
        name_ = key[2:]
        g = sd[name_].grad
        assert g is not None, name_
        d = digest(g.numpy())
        rms = math.sqrt(max(ref[1], 1e-300))
        if T == 1 and "cross_attn_image_to_token" in name_ and (".q_proj" in name_ or ".k_proj" in name_):
            assert ref[1] == 0.0 and d[1] == 0.0        # F10
            continue
        if name_.endswith("k_proj.bias") or name_.endswith("attention_weights.bias"):
            assert d[1] <= max(1e-10, 100 * ref[1])          # zero true gradient: float noise only
            live += 1
            continue
        assert abs(d[1] - ref[1]) <= 2e-3 * ref[1] + 1e-15, name_   # fp32 reference run: noise floor
        assert np.abs(d[2:] - ref[2:]).max() <= 2e-3 * max(np.abs(ref[2:]).max(), rms / math.sqrt(g.numel())) + 1e-9, name_
        live += 1
    # F12: the unused modules receive no gradient in the reference
    for name_ in dead:
        assert name_ in sd and sd[name_].grad is None, name_
    assert live > 80 and len(dead) > 100
    assert np.allclose(digest(x_p.grad.numpy())[1], fx["dx_p"][1], rtol=2e-3)


def clip_agg_shapes(ctpath=True, C=2):
    sh = {f"extractor_pathology.{k}": v for k, v in mo.abmil_shapes(768).items()}
    if ctpath:
        sh.update({"fc_CT.1.weight": (512, 512), "fc_CT.1.bias": (512,),
                   "fc_pathology.1.weight": (512, 768), "fc_pathology.1.bias": (512,),
                   "fc.1.weight": (C, 512), "fc.1.bias": (C,)})
    else:
        sh.update({"fc.1.weight": (C, 768), "fc.1.bias": (C,)})
    return sh


def test_aggregator_clip_vs_reference():
    fx = load_golden("aggregator_clip_ctpath")
    sd = fo.to_torch(mo.procedural_state(clip_agg_shapes(True), 71), requires_grad=True)
    x_ct, x_p = _t(rnd(171, 1, 512)), _t(rnd(271, 1, 83, 768), True)
    a, b, prob = fo.aggregator_clip_forward(sd, x_ct, [x_p])
    for got, key in ((a, "x_ct"), (b, "x_path"), (prob, "prob")):
        assert rel_err(got.detach().numpy(), fx[key]) < TOL64
    ((a * b).sum() + prob[0, 1]).backward()
    assert rel_err(x_p.grad.numpy(), fx["dx_p"]) < 1e-8
    grads = {k: (v.grad.numpy() if v.grad is not None else None) for k, v in sd.items()}
    assert check_grads(fx, grads, 1e-8, prefix_filter=lambda k: not k.endswith("attention_weights.bias")) >= 10
    fx = load_golden("aggregator_clip_path")
    sd = fo.to_torch(mo.procedural_state(clip_agg_shapes(False), 72))
    pooled, prob = fo.aggregator_clip_pathology_forward(sd, _t(rnd(272, 1, 64, 768)))
    assert rel_err(pooled.numpy(), fx["pooled"]) < TOL64 and rel_err(prob.numpy(), fx["prob"]) < TOL64


def wmask_shapes(C=2):
    sh = {f"extractor_pathology.{k}": v for k, v in mo.abmil_shapes(768).items()}
    sh.update({"fc.1.weight": (384, 1536), "fc.1.bias": (384,), "fc.4.weight": (C, 384), "fc.4.bias": (C,)})
    return sh


def test_aggregator_wmask_vs_reference():
    fx = load_golden("aggregator_wmask_ctpath")
    sd = fo.to_torch(mo.procedural_state(wmask_shapes(), 81))
    x_ct = torch.cat([_t(rnd(181, 1, 384)), _t(rnd(182, 1, 384))], dim=1)
    pooled = fo.abmil(sd, "extractor_pathology", _t(rnd(281, 1, 57, 768)))
    prob = fo.wmask_head_forward(sd, [x_ct, pooled])
    assert rel_err(prob.numpy(), fx["prob"]) < TOL64


def test_clip_logits_vs_reference():
    fx = load_golden("clip_logits_b24")
    img, txt = rnd(91, 24, 512), rnd(92, 24, 512)
    ls = math.log(1 / 0.07)
    li, lt = mo.clip_cosine_logits(img, txt, ls)
    assert rel_err(li, fx["li"]) < TOL64 and rel_err(lt, fx["lt"]) < TOL64
    dimg, dtxt, dscale = mo.clip_cosine_logits_bwd(img, txt, ls, rnd(93, 24, 24), rnd(94, 24, 24))
    assert rel_err(dimg, fx["dimg"]) < 1e-8 and rel_err(dtxt, fx["dtxt"]) < 1e-8
    assert abs(dscale - float(fx["dscale"])) < 1e-8 * abs(float(fx["dscale"]))


def test_cliploss_vs_reference():
    fx = load_golden("cliploss_v1_b6")
    out, feats = rnd(96, 6, 512, scale=0.3), rnd(95, 6, 9, 512, scale=0.3)
    loss, logits = mo.cliploss_v1(out, feats)
    assert logits.shape == (9, 6, 6)
    assert abs(loss - float(fx["loss"])) < 1e-5 * abs(float(fx["loss"]))     # reference ran in fp32
    assert rel_err(mo.cliploss_v1_bwd(out, feats), fx["dout"]) < 1e-5


def test_head_bce_and_adam_against_torch():
    x, W, b = rnd(1, 4, 512), rnd(2, 2, 512, scale=0.05), rnd(3, 2, scale=0.05)
    tgt = np.eye(2)[[0, 1, 1, 0]]
    loss, prob, dx, dW, db = mo.sigmoid_head_bce(x, W, b, tgt)
    xt, Wt, bt = _t(x, True), _t(W, True), _t(b, True)
    lt = torch.nn.BCELoss()(torch.sigmoid(xt @ Wt.t() + bt), _t(tgt))
    lt.backward()
    assert abs(loss - float(lt)) < 1e-12
    assert rel_err(dx, xt.grad.numpy()) < 1e-10 and rel_err(dW, Wt.grad.numpy()) < 1e-10
    assert rel_err(db, bt.grad.numpy()) < 1e-10
    p = torch.nn.Parameter(_t(W).clone())
    opt = torch.optim.Adam([p], lr=1e-5, betas=(0.9, 0.999), weight_decay=1e-7)
    pn, m, v = W.astype(np.float64), np.zeros_like(W, dtype=np.float64), np.zeros_like(W, dtype=np.float64)
    for step in (1, 2, 3):
        g = rnd(10 + step, 2, 512)
        p.grad = _t(g)
        opt.step()
        pn, m, v = mo.adam_step(pn, g, m, v, step)
    assert rel_err(pn, p.detach().numpy()) < 1e-12
    # the learnable-prompt configuration's optimiser (train_ddp.py:103-108)
    p = torch.nn.Parameter(_t(W).clone())
    opt = torch.optim.SGD([p], lr=1e-3, weight_decay=1e-7)
    pn = W.astype(np.float64)
    for step in (1, 2, 3):
        g = rnd(20 + step, 2, 512)
        p.grad = _t(g)
        opt.step()
        pn = mo.sgd_step(pn, g)
    assert rel_err(pn, p.detach().numpy()) < 1e-12


def test_pe_matches_reference_formula():
    a = mo.sinusoid_pe(300, 512)
    b = fo.sinusoid_pe(300, 512).numpy()
    assert np.abs(a - b).max() < 2e-4       # fp32 table upstream; sin/cos of up to 299 rad in fp32
