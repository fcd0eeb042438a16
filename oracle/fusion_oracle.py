"""torch *functional* CPU restatement of the reference's cross-modal fusion path and aggregator
glue.  Works in any dtype (float64 for ground truth, float32 for the CPU-baseline timing);
autograd provides the backward.

TEST INFRASTRUCTURE ONLY — see oracle/README.md.  Never imported by the product package.
Everything is a pure function of a flat ``state_dict``-style mapping ``sd`` (reference key names)
and a key ``prefix``, so the same fixture weights drive the reference, the oracle and the CUDA modules.

Reference lines restated (paths relative to the upstream repo):
  * Attention.forward                 model/sam/transformer.py:418-450
  * MLPBlock.forward                  model/sam/common.py:25-26
  * TwoWayAttentionBlock.forward      model/sam/transformer.py:278-309
  * TwoWayTransformer.forward         model/sam/transformer.py:58-120
  * ABMIL.forward                     model/dim1/ABMIL.py:47-64
  * aggregator.forward (CT+pathology) model/aggregator.py:134-203
  * aggregator_clip.forward           model/aggregator_clip.py:79-118
  * aggregator_wMask head             model/aggregator_wMask.py:67-70,114
"""
from __future__ import annotations

import math
import torch


def _lin(sd, prefix, x):
    return x @ sd[prefix + ".weight"].t() + sd[prefix + ".bias"]


def layer_norm(sd, prefix, x, eps=1e-5):
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * sd[prefix + ".weight"] + sd[prefix + ".bias"]


def attention(sd, prefix, q, k, v, num_heads):
    """transformer.py:418-450: project, split heads, softmax(QK^T/sqrt(c_head)) V, merge, out-proj.
    q (B,Nq,E), k,v (B,Nk,E)."""
    q = _lin(sd, prefix + ".q_proj", q)
    k = _lin(sd, prefix + ".k_proj", k)
    v = _lin(sd, prefix + ".v_proj", v)
    B, Nq, C = q.shape
    ch = C // num_heads

    def split(t):
        return t.reshape(t.shape[0], t.shape[1], num_heads, ch).permute(0, 2, 1, 3)

    qh, kh, vh = split(q), split(k), split(v)
    att = torch.softmax((qh @ kh.transpose(-1, -2)) / math.sqrt(ch), dim=-1)
    out = (att @ vh).permute(0, 2, 1, 3).reshape(B, Nq, C)
    return _lin(sd, prefix + ".out_proj", out)


def mlp_block(sd, prefix, x):
    """common.py:26 with act=ReLU (the activation TwoWayTransformer passes, transformer.py:18)."""
    return _lin(sd, prefix + ".lin2", torch.relu(_lin(sd, prefix + ".lin1", x)))


def two_way_block(sd, prefix, queries, keys, query_pe, key_pe, num_heads, skip_first_layer_pe):
    """transformer.py:278-309."""
    if skip_first_layer_pe:
        queries = attention(sd, prefix + ".self_attn", queries, queries, queries, num_heads)
    else:
        q = queries + query_pe
        queries = queries + attention(sd, prefix + ".self_attn", q, q, queries, num_heads)
    queries = layer_norm(sd, prefix + ".norm1", queries)

    q = queries + query_pe
    k = keys + key_pe
    queries = queries + attention(sd, prefix + ".cross_attn_token_to_image", q, k, keys, num_heads)
    queries = layer_norm(sd, prefix + ".norm2", queries)

    queries = layer_norm(sd, prefix + ".norm3", queries + mlp_block(sd, prefix + ".mlp", queries))

    q = queries + query_pe
    k = keys + key_pe
    keys = keys + attention(sd, prefix + ".cross_attn_image_to_token", k, q, queries, num_heads)
    keys = layer_norm(sd, prefix + ".norm4", keys)
    return queries, keys


def two_way_transformer(sd, prefix, image_embedding, image_pe, point_embedding,
                        depth=2, num_heads=8):
    """transformer.py:58-120 with alignment_base != 'CT' and model_CT == 'resnetMC3_18':
    a 5-D CT feature map (B,C,T,h,w) becomes tokens by mean over (h,w) then permute (:91-95)."""
    if image_embedding.dim() == 5:
        image_embedding = image_embedding.mean(dim=(3, 4)).permute(0, 2, 1)
    queries, keys = point_embedding, image_embedding
    for i in range(depth):
        queries, keys = two_way_block(sd, f"{prefix}.layers.{i}", queries, keys,
                                      point_embedding, image_pe, num_heads, i == 0)
    q = queries + point_embedding
    k = keys + image_pe
    queries = queries + attention(sd, prefix + ".final_attn_token_to_image", q, k, keys, num_heads)
    queries = layer_norm(sd, prefix + ".norm_final_attn", queries)
    return queries, keys


def abmil(sd, prefix, x):
    """ABMIL.py:47-64 in eval mode (dropout = identity). x (1,N,L) or (N,L) -> (1,L).
    A dense (B>1,N,L) input reproduces the upstream quirk (softmax over a size-1 axis)."""
    x = x.squeeze(0)
    A = _lin(sd, prefix + ".attention_weights",
             torch.tanh(_lin(sd, prefix + ".attention_V.0", x))
             * torch.sigmoid(_lin(sd, prefix + ".attention_U.0", x)))
    A = torch.softmax(A.transpose(-2, -1), dim=1)
    return A @ x


def sinusoid_pe(n_pos, dim, dtype=torch.float32):
    """aggregator.py:99-106 (built in float32 upstream, then cast)."""
    pe = torch.zeros((n_pos, dim))
    position = torch.arange(0, n_pos).unsqueeze(1).float()
    div_term = torch.exp(torch.arange(0, dim, 2, dtype=torch.float) * -(math.log(10000.0) / dim))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe.unsqueeze(0).to(dtype)


def aggregator_fusion_forward(sd, x_ct_feat, x_path, x_text, num_heads=8, depth=2, pe_fn=None):
    """aggregator.py:134-203, CT+pathology branch, aggregator='ABMIL', eval mode.
    pe_fn(n, E) -> (1, n, E) overrides the position table (tests of reduced-precision storage hand in the table the
    kernels read, e.g. the bf16-rounded one).
    x_ct_feat: the CT extractor's output (1,512,c,h,w) (the encoder itself is out of scope);
    x_path (1,N,768); x_text: clinic_extractor output (1,T,512).
    Returns (prob (1,C), x_CT2CI (1,T,512), x_Pth2CI (1,T,512))."""
    dt = x_path.dtype
    xin_path = torch.tanh(_lin(sd, "fc_pathology.0", x_path))                      # :141
    c = x_ct_feat.shape[2]                                                          # :156
    E = xin_path.shape[-1]
    pe_of = (lambda n_: sinusoid_pe(n_, E, dt)) if pe_fn is None else (lambda n_: pe_fn(n_, E))
    ct2ci, ci2ct = two_way_transformer(sd, "TwoWayTransformer_Both", x_ct_feat,
                                       pe_of(c),
                                       torch.tanh(_lin(sd, "fc_CI2CT.0", x_text)),
                                       depth, num_heads)                            # :160
    n = xin_path.shape[1]
    pth2ci, ci2pth = two_way_transformer(sd, "TwoWayTransformer_Both", xin_path,
                                         pe_of(n),
                                         torch.tanh(_lin(sd, "fc_CI2Pth.0", x_text)),
                                         depth, num_heads)                          # :168
    bag = torch.cat([ct2ci, ci2ct, pth2ci, ci2pth], dim=1)                          # :173
    pooled = abmil(sd, "aggregator", bag)                                           # :199
    prob = torch.sigmoid(_lin(sd, "fc.1", pooled))                                  # :200
    return prob, ct2ci, pth2ci


def aggregator_clip_forward(sd, x_ct_feat, x_path_bags):
    """aggregator_clip.py:82-96 (CT+pathology), eval mode.  x_ct_feat (B,512) = CT extractor output;
    x_path_bags: list of (1,N_b,768) bags (the reference runs batch 1; B bags are stacked here).
    Returns (x_CT (B,512), x_pathology (B,512), prob (B,C))."""
    x_ct = torch.relu(_lin(sd, "fc_CT.1", x_ct_feat))
    pooled = torch.cat([abmil(sd, "extractor_pathology", xb) for xb in x_path_bags], dim=0)
    x_p = torch.relu(_lin(sd, "fc_pathology.1", pooled))
    x = (x_ct + x_p) / 2
    return x_ct, x_p, torch.sigmoid(_lin(sd, "fc.1", x))


def aggregator_clip_pathology_forward(sd, x_path):
    """aggregator_clip.py:109-118 (pathology only): returns (pooled (1,768), prob (1,C))."""
    pooled = abmil(sd, "extractor_pathology", x_path)
    return pooled, torch.sigmoid(_lin(sd, "fc.1", pooled))


def wmask_head_forward(sd, feats):
    """aggregator_wMask.py:67-70,114 (eval): sigmoid(Linear(ReLU(Linear(cat feats))))."""
    x = torch.cat(feats, dim=1)
    return torch.sigmoid(_lin(sd, "fc.4", torch.relu(_lin(sd, "fc.1", x))))


def to_torch(sd_np, dtype=torch.float64, requires_grad=False):
    out = {}
    for k, v in sd_np.items():
        t = torch.as_tensor(v).to(dtype).clone()
        if requires_grad:
            t.requires_grad_(True)
        out[k] = t
    return out
