"""numpy float64 restatement of the reference's gated-attention MIL pool and the small
loss/logit maths around it.

TEST INFRASTRUCTURE ONLY — see oracle/README.md.  Never imported by the product package.
Pinned against fixtures produced by the unmodified reference (tests/golden/make_golden.py).

Reference lines restated (paths relative to the upstream repo):
  * ABMIL.forward                     model/dim1/ABMIL.py:47-64
  * ABMIL_v2.forward (concat tail)    model/dim1/ABMIL_v2.py:49-69
  * sinusoidal PE table               model/aggregator.py:99-106
  * CLIP.forward cosine logits        clip/model.py:354-368
  * CLIPloss_v1 logits + CE           utils.py:277-282
  * BCELoss / CosineEmbeddingLoss use train_ddp.py:99,102,319-326
  * Adam(lr, betas, weight_decay)     train_ddp.py:111-118
"""
from __future__ import annotations

import math
import numpy as np

F64 = np.float64


def _f(a):
    return np.asarray(a, dtype=F64)


def _sigmoid(z):
    # numerically stable logistic
    out = np.empty_like(z)
    pos = z >= 0
    out[pos] = 1.0 / (1.0 + np.exp(-z[pos]))
    ez = np.exp(z[~pos])
    out[~pos] = ez / (1.0 + ez)
    return out


# --------------------------------------------------------------------------------------
# gated-attention MIL pool, one bag                               model/dim1/ABMIL.py:47-64
# --------------------------------------------------------------------------------------
def gated_scores(x, Wv, bv, Wu, bu, ww, bw):
    """s_i = (tanh(x_i Wv^T + bv) * sigmoid(x_i Wu^T + bu)) . ww + bw      (ABMIL.py:52-54)

    x (N,L); Wv,Wu (D,L); bv,bu (D,); ww (D,) [the (1,D) weight row]; bw scalar.
    Returns s (N,), V (N,D), U (N,D)."""
    x, Wv, bv, Wu, bu, ww = map(_f, (x, Wv, bv, Wu, bu, ww))
    V = np.tanh(x @ Wv.T + bv)
    U = _sigmoid(x @ Wu.T + bu)
    s = (V * U) @ ww.reshape(-1) + float(np.asarray(bw).reshape(-1)[0])
    return s, V, U


def softmax_pool(x, s):
    """a = softmax_N(s); M = a^T x                                      (ABMIL.py:56-59)
    Returns M (L,), a (N,), argmax (first index of the max score, torch.argmax tie rule)."""
    x, s = _f(x), _f(s)
    m = s.max()
    e = np.exp(s - m)
    a = e / e.sum()
    return a @ x, a, int(np.argmax(s))


def abmil_forward(p, x):
    """p: dict with the reference state_dict keys (attention_V.0.weight, ...). x (N,L).
    Returns dict(M (1,L), s, a, argmax, V, U)."""
    s, V, U = gated_scores(x, p["attention_V.0.weight"], p["attention_V.0.bias"],
                           p["attention_U.0.weight"], p["attention_U.0.bias"],
                           p["attention_weights.weight"], p["attention_weights.bias"])
    M, a, am = softmax_pool(x, s)
    return dict(M=M[None, :], s=s, a=a, argmax=am, V=V, U=U)


def abmil_backward(p, x, dM, need_dx=True):
    """Analytic backward of abmil_forward for upstream gradient dM (1,L) or (L,).
    Formulas: SURVEY Appendix A (derived from ABMIL.py:52-59).
    Returns dict of parameter grads under state_dict keys, plus 'x' (N,L) if need_dx and 'ds'."""
    x = _f(x)
    dM = _f(dM).reshape(-1)
    f = abmil_forward(p, x)
    a, V, U, M = f["a"], f["V"], f["U"], f["M"].reshape(-1)
    ww = _f(p["attention_weights.weight"]).reshape(-1)
    g = x @ dM                                  # g_i = dM . x_i
    ds = a * (g - float(dM @ M))                # softmax bwd; a.g == dM.M
    dG = ds[:, None] * ww[None, :]              # grad wrt (V*U)
    dVp = dG * U * (1.0 - V * V)                # through tanh
    dUp = dG * V * U * (1.0 - U)                # through sigmoid
    out = {
        "attention_V.0.weight": dVp.T @ x,
        "attention_V.0.bias": dVp.sum(0),
        "attention_U.0.weight": dUp.T @ x,
        "attention_U.0.bias": dUp.sum(0),
        "attention_weights.weight": (ds @ (V * U))[None, :],
        "attention_weights.bias": np.array([ds.sum()]),
        "ds": ds,
    }
    if need_dx:
        out["x"] = (a[:, None] * dM[None, :]
                    + dVp @ _f(p["attention_V.0.weight"])
                    + dUp @ _f(p["attention_U.0.weight"]))
    return out


def abmil_v2_forward(p, x, bprc):
    """ABMIL_v2: pool (L fixed at 768 upstream) then cat([M, BpRc_class], dim=1)
    (ABMIL_v2.py:66).  bprc (1,1)."""
    f = abmil_forward(p, x)
    f["M"] = np.concatenate([f["M"], _f(bprc).reshape(1, -1)], axis=1)
    return f


def abmil_dense_batched(p, x):
    """The reference called on a dense (B,N,L) batch with B>1: squeeze(0) is a no-op, A is
    (B,1,N) after the transpose and softmax(dim=1) normalises over the size-1 axis, so every
    weight is exactly 1 and M[b] = sum_i x[b,i]  (ABMIL.py:48,56-59).
    Returns M (B,1,L)."""
    x = _f(x)
    assert x.ndim == 3 and x.shape[0] > 1
    return x.sum(axis=1, keepdims=True)


# --------------------------------------------------------------------------------------
# packed CSR batches = "loop the reference over bags with batch 1" (train_ddp.py:75)
# --------------------------------------------------------------------------------------
def abmil_forward_csr(p, X, offsets):
    """X (sumN,L) packed, offsets (B+1,) int. Returns M (B,L), s (sumN,), argmax (B,) local idx."""
    offsets = np.asarray(offsets, dtype=np.int64)
    B = len(offsets) - 1
    Ms, ss, am = [], [], []
    for b in range(B):
        lo, hi = int(offsets[b]), int(offsets[b + 1])
        f = abmil_forward(p, _f(X)[lo:hi])
        Ms.append(f["M"][0]); ss.append(f["s"]); am.append(f["argmax"])
    return np.stack(Ms), np.concatenate(ss), np.asarray(am, dtype=np.int64)


def abmil_backward_csr(p, X, offsets, dM, need_dx=True):
    """Sum of per-bag parameter grads (what one backward over the sum of per-bag losses gives)."""
    offsets = np.asarray(offsets, dtype=np.int64)
    B = len(offsets) - 1
    acc, dxs, dss = None, [], []
    for b in range(B):
        lo, hi = int(offsets[b]), int(offsets[b + 1])
        g = abmil_backward(p, _f(X)[lo:hi], _f(dM)[b], need_dx=need_dx)
        dss.append(g.pop("ds"))
        if need_dx:
            dxs.append(g.pop("x"))
        if acc is None:
            acc = g
        else:
            for k in g:
                acc[k] = acc[k] + g[k]
    acc["ds"] = np.concatenate(dss)
    if need_dx:
        acc["x"] = np.concatenate(dxs, axis=0)
    return acc


def abmil_forward_masked(p, Xpad, lengths):
    """Padded (B,Nmax,L) bags with valid lengths: the pool of each *unpadded* bag
    (cfg 4 oracle: reference ABMIL on x[b,:len_b])."""
    Xpad = _f(Xpad)
    return np.stack([abmil_forward(p, Xpad[b, :int(n)])["M"][0] for b, n in enumerate(lengths)])


# --------------------------------------------------------------------------------------
# sinusoidal positional table                                  model/aggregator.py:99-106
# --------------------------------------------------------------------------------------
def sinusoid_pe(n_pos, dim=512):
    """pe[p,2i]=sin(p*w_i), pe[p,2i+1]=cos(p*w_i), w_i=exp(-2i*ln(1e4)/dim). Returns (1,n_pos,dim).
    The reference builds the table in float32; callers compare with fp32 tolerance."""
    pos = np.arange(n_pos, dtype=F64)[:, None]
    div = np.exp(np.arange(0, dim, 2, dtype=F64) * -(math.log(10000.0) / dim))
    pe = np.zeros((n_pos, dim), dtype=F64)
    pe[:, 0::2] = np.sin(pos * div)
    pe[:, 1::2] = np.cos(pos * div)
    return pe[None]


# --------------------------------------------------------------------------------------
# CLIP-style logits and losses
# --------------------------------------------------------------------------------------
def clip_cosine_logits(img, txt, logit_scale):
    """clip/model.py:359-368: L2-normalise both, logits_per_image = exp(logit_scale) I T^T,
    logits_per_text = its transpose."""
    img, txt = _f(img), _f(txt)
    i = img / np.linalg.norm(img, axis=-1, keepdims=True)
    t = txt / np.linalg.norm(txt, axis=-1, keepdims=True)
    li = math.exp(float(logit_scale)) * (i @ t.T)
    return li, li.T.copy()


def clip_cosine_logits_bwd(img, txt, logit_scale, d_li, d_lt):
    """Analytic gradient of clip_cosine_logits wrt img, txt, logit_scale."""
    img, txt = _f(img), _f(txt)
    ni = np.linalg.norm(img, axis=-1, keepdims=True)
    nt = np.linalg.norm(txt, axis=-1, keepdims=True)
    i, t = img / ni, txt / nt
    sc = math.exp(float(logit_scale))
    G = _f(d_li) + _f(d_lt).T                    # total grad on the (b_i, b_t) cosine matrix * sc
    d_i = sc * (G @ t)
    d_t = sc * (G.T @ i)
    d_img = (d_i - i * (d_i * i).sum(-1, keepdims=True)) / ni
    d_txt = (d_t - t * (d_t * t).sum(-1, keepdims=True)) / nt
    d_scale = float((G * (sc * (i @ t.T))).sum())
    return d_img, d_txt, d_scale


def cliploss_v1(output, text_feat):
    """utils.py:277-282.  output (b,512); text_feat (b,I,512) [feature_by_CLIP].
    logits[i] = output @ text_feat[:,i,:].T  -> (I,b,b)  (no normalisation, no temperature);
    labels = eye(b) repeated I times; torch CrossEntropyLoss with *probability* targets and the
    class axis = dim 1:  loss = -(1/(I*b)) * sum_{i,c,k} labels[i,c,k] * log_softmax(logits, axis=1)[i,c,k]
    (mean over the I*b non-class positions).  Returns (loss, logits)."""
    output, text_feat = _f(output), _f(text_feat)
    logits = np.einsum("rd,cid->irc", output, text_feat)       # (I, b_row, b_col)
    # torch: input (N=I, C=b_row, d1=b_col): softmax over axis 1
    z = logits - logits.max(axis=1, keepdims=True)
    lsm = z - np.log(np.exp(z).sum(axis=1, keepdims=True))
    b = output.shape[0]
    eye = np.eye(b)[None]
    loss = -(eye * lsm).sum() / (logits.shape[0] * b)
    return float(loss), logits


def cliploss_v1_bwd(output, text_feat):
    """d loss / d output for cliploss_v1 (text features are frozen upstream, utils.py:272)."""
    output, text_feat = _f(output), _f(text_feat)
    logits = np.einsum("rd,cid->irc", output, text_feat)
    z = logits - logits.max(axis=1, keepdims=True)
    sm = np.exp(z); sm /= sm.sum(axis=1, keepdims=True)
    b = output.shape[0]
    # d/dlogits[i,r,c] = (softmax_r[i,r,c]*sum_r' eye[r',c] - eye[r,c]) / (I*b)
    dlog = (sm - np.eye(b)[None]) / (logits.shape[0] * b)
    return np.einsum("irc,cid->rd", dlog, text_feat)


def sigmoid_head_bce(x, W, b, target):
    """prob = sigmoid(x W^T + b)  (aggregator.py:200; eval-mode dropout = identity),
    loss = BCELoss(prob, target) mean over all elements (train_ddp.py:99,319).
    Returns (loss, prob, dx, dW, db)."""
    x, W, b, target = _f(x), _f(W), _f(b), _f(target)
    z = x @ W.T + b
    prob = _sigmoid(z)
    # torch clamps log at -100
    lp = np.maximum(np.log(prob), -100.0)
    l1p = np.maximum(np.log1p(-prob), -100.0)
    loss = float(-(target * lp + (1 - target) * l1p).mean())
    dz = (prob - target) / prob.size
    return loss, prob, dz @ W, dz.T @ x, dz.sum(0)


def cosine_embedding_loss_pos(a, b):
    """CosineEmbeddingLoss(a, b, target=+1) = mean(1 - cos(a_i, b_i))  (train_ddp.py:102,326)."""
    a, b = _f(a), _f(b)
    eps = 1e-12   # ATen cosine_embedding_loss EPSILON
    cos = (a * b).sum(-1) / np.sqrt(((a * a).sum(-1) + eps) * ((b * b).sum(-1) + eps))
    return float((1.0 - cos).mean())


def adam_step(param, grad, m, v, step, lr=1e-5, b1=0.9, b2=0.999, eps=1e-8, wd=1e-7):
    """torch.optim.Adam (L2 weight decay added to the gradient), train_ddp.py:111-118.
    step is the 1-based step count *after* increment. Returns (param, m, v)."""
    param, grad, m, v = map(_f, (param, grad, m, v))
    g = grad + wd * param
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    mhat = m / (1 - b1 ** step)
    vhat = v / (1 - b2 ** step)
    return param - lr * mhat / (np.sqrt(vhat) + eps), m, v


def sgd_step(param, grad, lr=1e-3, wd=1e-7):
    """torch.optim.SGD(lr, weight_decay) without momentum, as train_ddp.py:103-108 builds it."""
    param, grad = map(_f, (param, grad))
    return param - lr * (grad + wd * param)


# --------------------------------------------------------------------------------------
# synthetic-data helpers shared by tests, golden generation and the bench (no reference analogue:
# the real data are private, dataset.py:366-393 only fixes the shapes)
# --------------------------------------------------------------------------------------
def procedural_state(shapes, seed, scale=None):
    """Fill a {name: shape} dict deterministically (sorted by name) from RandomState(seed).
    Linear-like tensors get U(-1/sqrt(fan_in), 1/sqrt(fan_in)) like nn.Linear's default init;
    names containing 'norm' get weight 1+0.1n / bias 0.1n so LayerNorm affine terms are exercised."""
    rs = np.random.RandomState(seed)
    out = {}
    for name in sorted(shapes):
        shp = tuple(shapes[name])
        if "norm" in name:
            base = 1.0 if name.endswith("weight") else 0.0
            out[name] = (base + 0.1 * rs.standard_normal(shp)).astype(np.float32)
            continue
        if name.endswith("weight") and len(shp) >= 2:
            fan_in = shp[-1]
        elif name.endswith("bias"):
            fan_in = None
        else:
            fan_in = shp[-1] if len(shp) else 1
        if scale is not None:
            bound = scale
        elif fan_in is None:
            bound = 0.05
        else:
            bound = 1.0 / math.sqrt(fan_in)
        out[name] = rs.uniform(-bound, bound, size=shp).astype(np.float32)
    return out


def abmil_shapes(L, D=192, K=1):
    return {
        "attention_V.0.weight": (D, L), "attention_V.0.bias": (D,),
        "attention_U.0.weight": (D, L), "attention_U.0.bias": (D,),
        "attention_weights.weight": (K, D), "attention_weights.bias": (K,),
    }


def ragged_lengths(B, lo, hi, seed):
    """Bag sizes n_b ~ randint[lo, hi] (inclusive), reproducible everywhere."""
    return np.random.RandomState(seed).randint(lo, hi + 1, size=B).astype(np.int64)


def offsets_from_lengths(lengths):
    off = np.zeros(len(lengths) + 1, dtype=np.int32)
    off[1:] = np.cumsum(lengths)
    return off


def score_margin(s, offsets):
    """min over bags of (top1 - top2) of the raw scores — the argmax guard."""
    mg = np.inf
    for b in range(len(offsets) - 1):
        seg = np.sort(np.asarray(s[offsets[b]:offsets[b + 1]], dtype=F64))
        if len(seg) > 1:
            mg = min(mg, seg[-1] - seg[-2])
    return mg
