"""Cross-modal fusion modules — drop-ins for model/sam/transformer.py (TwoWayTransformer,
TwoWayAttentionBlock, Attention).

Same constructor arguments, forward signatures, outputs and ``state_dict`` keys as upstream
(``layers.{i}.self_attn.q_proj.weight`` ...).  The ``nn.Linear`` / ``nn.LayerNorm`` children are parameter
containers only; every forward/backward runs in libmilb200:

  * q/k/v/out projections and the MLP      -> ``milb200_linear_fwd/bwd`` (tcgen05 GEMM for bf16 image-side
    rows, FFMA for fp32), with the ``x + pe`` sums of transformer.py:291-292,303-304 folded into the call;
  * softmax(QK^T/sqrt(c))V per head        -> ``milb200_attention_fwd/bwd`` (one side is always the <=16
    text tokens: token->image streams K,V once with an online softmax, image->token keeps K,V in smem);
  * residual + LayerNorm                   -> ``milb200_layernorm_fwd/bwd`` (one kernel, residual fused).

The reference runs one bag per call (batch 1, train_ddp.py:75); a leading batch B>1 is looped.
"""
from __future__ import annotations

from typing import Tuple, Type

import torch
from torch import Tensor, nn

import os

from ... import functional as F
from ..._lib import MilB200Error
from ...tape import Tape
from .common import MLPBlock


def _use_tape():
    """MILB200_NO_TAPE=1 falls back to one autograd node per operator (same kernels, same numerics; ~8x the host
    time per bag) — kept for debugging and for the tape-vs-eager equivalence test."""
    return os.environ.get("MILB200_NO_TAPE", "0") != "1"


def _use_collapsed():
    """MILB200_FUSION_COLLAPSED=0 keeps the T = 1 CT+pathology branch on the projected-keys program (round-1 path: six
    N x 512 x 256 GEMMs per TwoWayTransformer call) instead of the collapsed one (csrc/xfusion.cu)."""
    return os.environ.get("MILB200_FUSION_COLLAPSED", "1") != "0"


class Attention(nn.Module):
    """model/sam/transformer.py:395-450."""

    def __init__(self, embedding_dim: int, num_heads: int, downsample_rate: int = 1) -> None:
        super().__init__()
        self.embedding_dim = embedding_dim
        self.internal_dim = embedding_dim // downsample_rate
        self.num_heads = num_heads
        assert self.internal_dim % num_heads == 0, "num_heads must divide embedding_dim."
        self.q_proj = nn.Linear(embedding_dim, self.internal_dim)
        self.k_proj = nn.Linear(embedding_dim, self.internal_dim)
        self.v_proj = nn.Linear(embedding_dim, self.internal_dim)
        self.out_proj = nn.Linear(self.internal_dim, embedding_dim)

    def forward(self, q: Tensor, k: Tensor, v: Tensor, q_add: Tensor = None, k_add: Tensor = None) -> Tensor:
        """``q_add`` / ``k_add`` (extension, optional): positional terms added to q / k inside the projection
        kernels, so callers need not materialise ``queries + query_pe`` / ``keys + key_pe``."""
        if q.dim() != 3 or k.dim() != 3 or v.dim() != 3:
            raise MilB200Error("Attention.forward expects (B, N_tokens, C) tensors")
        qp = F.linear(q, self.q_proj.weight, self.q_proj.bias, add=q_add)          # transformer.py:430
        kp = F.linear(k, self.k_proj.weight, self.k_proj.bias, add=k_add)          # :431
        vp = F.linear(v, self.v_proj.weight, self.v_proj.bias)                     # :432
        outs = [F.attention_core(qp[b], kp[b], vp[b], self.num_heads) for b in range(q.shape[0])]   # :434-446
        o = outs[0].unsqueeze(0) if len(outs) == 1 else torch.stack(outs, dim=0)
        return F.linear(o, self.out_proj.weight, self.out_proj.bias)               # :448


    def emit(self, tape: Tape, q: int, k: int, v: int, q_add: int = None, k_add: int = None) -> int:
        """The same computation recorded on a tape (slot ids in, slot id out)."""
        qp = tape.linear(q, self.q_proj, add=q_add)
        kp = tape.linear(k, self.k_proj, add=k_add)
        vp = tape.linear(v, self.v_proj)
        o = tape.attention(qp, kp, vp, self.num_heads)
        return tape.linear(o, self.out_proj)


    def emit_single_key(self, tape: Tape, v: int) -> int:
        """Attention over exactly ONE key (a single text token): softmax over one score is 1.0 for every query and head
        (transformer.py:441-443), so the output row is out_proj(v_proj(v)) for every query and q_proj / k_proj drop out of
        the computation with exactly-zero gradients, as in the reference.  Returns a 1-row slot."""
        for lin in (self.q_proj, self.k_proj):      # still parameters of the program: they receive (exactly) zero gradients
            tape.param(lin.weight)
            tape.param(lin.bias)
        return tape.linear(tape.linear(v, self.v_proj), self.out_proj)


class TwoWayAttentionBlock(nn.Module):
    """model/sam/transformer.py:236-309."""

    def __init__(self, embedding_dim: int, num_heads: int, mlp_dim: int = 2048,
                 activation: Type[nn.Module] = nn.ReLU, attention_downsample_rate: int = 2,
                 skip_first_layer_pe: bool = False) -> None:
        super().__init__()
        self.self_attn = Attention(embedding_dim, num_heads)
        self.norm1 = nn.LayerNorm(embedding_dim)
        self.cross_attn_token_to_image = Attention(embedding_dim, num_heads, downsample_rate=attention_downsample_rate)
        self.norm2 = nn.LayerNorm(embedding_dim)
        self.mlp = MLPBlock(embedding_dim, mlp_dim, activation)
        self.norm3 = nn.LayerNorm(embedding_dim)
        self.norm4 = nn.LayerNorm(embedding_dim)
        self.cross_attn_image_to_token = Attention(embedding_dim, num_heads, downsample_rate=attention_downsample_rate)
        self.skip_first_layer_pe = skip_first_layer_pe

    @staticmethod
    def _ln(norm, x, residual=None):
        if abs(norm.eps - 1e-5) > 1e-12:
            raise MilB200Error("layernorm kernels are built for eps = 1e-5 (nn.LayerNorm default)")
        return F.layernorm(x, norm.weight, norm.bias, residual=residual)

    def forward(self, queries: Tensor, keys: Tensor, query_pe: Tensor, key_pe: Tensor) -> Tuple[Tensor, Tensor]:
        # (1) token self-attention (:281-288)
        if self.skip_first_layer_pe:
            queries = self._ln(self.norm1, self.self_attn(q=queries, k=queries, v=queries))
        else:
            attn_out = self.self_attn(q=queries, k=queries, v=queries, q_add=query_pe, k_add=query_pe)
            queries = self._ln(self.norm1, queries, residual=attn_out)
        # (2) tokens attend to the image bag (:290-295)
        attn_out = self.cross_attn_token_to_image(q=queries, k=keys, v=keys, q_add=query_pe, k_add=key_pe)
        queries = self._ln(self.norm2, queries, residual=attn_out)
        # (3) MLP on the tokens (:297-300)
        queries = self._ln(self.norm3, queries, residual=self.mlp(queries))
        # (4) the image bag attends to the tokens (:302-307)
        attn_out = self.cross_attn_image_to_token(q=keys, k=queries, v=queries, q_add=key_pe, k_add=query_pe)
        keys = self._ln(self.norm4, keys, residual=attn_out)
        return queries, keys


    def emit(self, tape: Tape, queries: int, keys: int, query_pe: int, key_pe: int, single_token: bool = False):
        """forward() recorded on a tape.  single_token: the program is specialised for T = 1 (one clinical prompt, the
        reference's active configuration, dataset.py:479-480): every attention whose KEYS are the tokens has one key."""
        if single_token:
            sa = self.self_attn.emit_single_key(tape, queries)
            queries = tape.layernorm(sa, self.norm1) if self.skip_first_layer_pe else \
                tape.layernorm(queries, self.norm1, residual=sa)
        elif self.skip_first_layer_pe:
            queries = tape.layernorm(self.self_attn.emit(tape, queries, queries, queries), self.norm1)
        else:
            attn_out = self.self_attn.emit(tape, queries, queries, queries, q_add=query_pe, k_add=query_pe)
            queries = tape.layernorm(queries, self.norm1, residual=attn_out)
        keys_pe = tape.add(keys, key_pe)      # transformer.py:292 and :304 are the same tensor: build it once
        attn_out = self.cross_attn_token_to_image.emit(tape, queries, keys_pe, keys, q_add=query_pe)
        queries = tape.layernorm(queries, self.norm2, residual=attn_out)
        queries = tape.layernorm(queries, self.norm3, residual=self.mlp.emit(tape, queries))
        if single_token:
            # image -> token attention with one key: the same row for every instance, added by the LayerNorm kernel
            row = self.cross_attn_image_to_token.emit_single_key(tape, queries)
            keys = tape.layernorm(keys, self.norm4, residual=row)
        else:
            attn_out = self.cross_attn_image_to_token.emit(tape, keys_pe, queries, queries, k_add=query_pe)
            keys = tape.layernorm(keys, self.norm4, residual=attn_out)
        return queries, keys


class TwoWayTransformer(nn.Module):
    """model/sam/transformer.py:10-120.  ``args`` supplies ``alignment_base`` and ``model_CT`` like upstream."""

    def __init__(self, args, depth: int, embedding_dim: int, num_heads: int, mlp_dim: int,
                 activation: Type[nn.Module] = nn.ReLU, attention_downsample_rate: int = 2) -> None:
        super().__init__()
        self.args = args
        self.depth = depth
        self.embedding_dim = embedding_dim
        self.num_heads = num_heads
        self.mlp_dim = mlp_dim
        self.layers = nn.ModuleList()
        for i in range(depth):
            self.layers.append(TwoWayAttentionBlock(embedding_dim=embedding_dim, num_heads=num_heads, mlp_dim=mlp_dim,
                                                    activation=activation,
                                                    attention_downsample_rate=attention_downsample_rate,
                                                    skip_first_layer_pe=(i == 0)))
        self.final_attn_token_to_image = Attention(embedding_dim, num_heads, downsample_rate=attention_downsample_rate)
        self.norm_final_attn = nn.LayerNorm(embedding_dim)

    def _ct_to_tokens(self, x: Tensor) -> Tensor:
        """(B, C, T, h, w) CT feature map -> (B, tokens, C)  (transformer.py:79-95)."""
        model_ct = getattr(self.args, "model_CT", None)
        if model_ct == "resnetMC3_18":
            return F.ct_tokens(x)                                   # mean over (h, w), permute: one kernel
        if model_ct == "medicalNet":
            return x.flatten(2).permute(0, 2, 1).contiguous()       # pure data movement
        return x                                                     # upstream leaves other encoders untouched

    def emit(self, tape: Tape, image: int, image_pe: int, points: int, single_token: bool = False):
        """The token/image program of forward() on a tape: returns (queries slot, keys slot)."""
        queries, keys = points, image
        for layer in self.layers:
            queries, keys = layer.emit(tape, queries, keys, points, image_pe, single_token=single_token)
        attn_out = self.final_attn_token_to_image.emit(tape, queries, keys, keys, q_add=points, k_add=image_pe)
        queries = tape.layernorm(queries, self.norm_final_attn, residual=attn_out)
        return queries, keys

    # ---- collapsed program: key/value projections folded into the token side (csrc/xfusion.cu) --------------------------
    @staticmethod
    def _emit_t2i(tape: Tape, att: "Attention", norm, queries: int, points: int, keys: int, pe: int, bag_layout: bool) -> int:
        """queries = LN(queries + cross_attn_token_to_image(q=queries + points, k=keys + pe, v=keys))  (transformer.py:290-295,
        114-118) without projecting the image tokens: the projected queries are pulled through k_proj.weight per head
        (U = Wk_h^T q_h), ONE pass over the keys pools them per head, and v_proj is applied to the 8 pooled rows."""
        qx = tape.linear(queries, att.q_proj, add=points)                       # [S*T, 256]
        u = tape.headdiag_u(qx, att.k_proj, "SJ")                               # [S*T*8, 512]
        tape.param(att.k_proj.bias)     # constant over the keys: cancels in the softmax, exactly-zero gradient (as upstream, up to noise)
        pooled = tape.t2i_pool(keys, pe, u, bag_layout=bag_layout)              # [S*T*8, 512]
        o = tape.headdiag_o(pooled, att.v_proj, "ST")                           # [S*T, 256]
        return tape.layernorm(queries, norm, residual=tape.linear(o, att.out_proj))

    def emit_collapsed(self, tape: Tape, keys: int, pe: int, points: int):
        """forward() for ONE text token per segment over a segmented key stream (every image-side bag that shares these
        weights is a segment: the CT bag and the pathology bag of aggregator.py:160,168, of one or several patients).
        Returns (slot of the final token rows [S, E] fp32, slot of the final keys — written in the packed-bag layout)."""
        queries = points
        last = len(self.layers) - 1
        for i, layer in enumerate(self.layers):
            sa = layer.self_attn.emit_single_key(tape, queries)                              # :281-288, one key
            queries = tape.layernorm(sa, layer.norm1) if layer.skip_first_layer_pe else \
                tape.layernorm(queries, layer.norm1, residual=sa)
            queries = self._emit_t2i(tape, layer.cross_attn_token_to_image, layer.norm2, queries, points, keys, pe, False)
            queries = tape.layernorm(queries, layer.norm3, residual=layer.mlp.emit(tape, queries))   # :297-300
            row = layer.cross_attn_image_to_token.emit_single_key(tape, queries)             # :302-307, one key per segment
            keys = tape.ln_seg(keys, row, layer.norm4, out_rows_key="NBAG" if i == last else None)
        queries = self._emit_t2i(tape, self.final_attn_token_to_image, self.norm_final_attn, queries, points, keys, pe, True)
        return queries, keys

    def _tape(self, single_token=False):
        name = "_tape_cache_t1" if single_token else "_tape_cache"
        t = getattr(self, name, None)
        if t is None:
            t = Tape()
            E = self.embedding_dim
            img, pe, pts = t.input("N", E), t.input("N", E), t.input("T", E)
            q, k = self.emit(t, img, pe, pts, single_token=single_token)
            bq, bk = t.buffer(lambda r: r["T"], E), t.buffer(lambda r: r["N"], E)
            t.output(q, bq, lambda r: 0)
            t.output(k, bk, lambda r: 0)
            object.__setattr__(self, name, t)      # not a submodule / parameter: keep it out of nn.Module state
        return t

    def forward(self, image_embedding: Tensor, image_pe: Tensor, point_embedding: Tensor) -> Tuple[Tensor, Tensor]:
        if getattr(self.args, "alignment_base", None) == "CT":
            if point_embedding.dim() == 5:
                point_embedding = self._ct_to_tokens(point_embedding)
        elif image_embedding.dim() == 5:
            image_embedding = self._ct_to_tokens(image_embedding)
        if (_use_tape() and image_embedding.dim() == 3 and image_embedding.shape[0] == 1 and point_embedding.shape[0] == 1
                and image_pe.shape == image_embedding.shape):
            # one bag per call (train_ddp.py:75): the whole program is a single native call each way
            n, t_ = image_embedding.shape[1], point_embedding.shape[1]
            q, k = self._tape(single_token=(t_ == 1 and n > 1)).run({"N": n, "T": t_},
                                                                    [image_embedding[0], image_pe[0], point_embedding[0]])
            return q.unsqueeze(0), k.unsqueeze(0)
        queries, keys = point_embedding, image_embedding
        for layer in self.layers:                                                     # :105-111
            queries, keys = layer(queries=queries, keys=keys, query_pe=point_embedding, key_pe=image_pe)
        attn_out = self.final_attn_token_to_image(q=queries, k=keys, v=keys, q_add=point_embedding,
                                                  k_add=image_pe)                     # :114-116
        queries = TwoWayAttentionBlock._ln(self.norm_final_attn, queries, residual=attn_out)   # :117-118
        return queries, keys
