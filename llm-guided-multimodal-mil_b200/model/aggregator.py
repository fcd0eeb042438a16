"""Multimodal aggregator — drop-in for model/aggregator.py:9-209 (``aggregator(args).forward(x_list, x_CI)``).

Same ``args`` fields, forward signature, outputs and ``state_dict`` keys (``fc_pathology.0``, ``fc_CI2CT.0``,
``fc_CI2Pth.0``, ``fc_CI.0``, ``TwoWayTransformer_{CT,Pth,Both}``, ``extractor_pathology``, ``aggregator``,
``prompt_embedding``, ``fc.1``).  Differences, all outside the hot path:
  * the CT encoders (model/dim3) and the clinical-text encoder (model/dim1/CLIP.py, simpleFCs) are upstream,
    frozen or cuDNN-bound components that this library does not rebuild: by default ``x_list[0]`` is the CT
    encoder's OUTPUT feature map (1, 512, c, h, w) and ``x_CI`` the text encoder's OUTPUT (1, T, 512); real
    encoders can be injected through the constructor;
  * ``pe`` is generated on the device by a kernel and cached, instead of a 205 MB host table copied
    host->device on every forward; the attribute ``pe`` is kept, shape (1, 100000, 512);
  * TransMIL aggregators raise: upstream they return a tuple that ``self.fc`` cannot consume.
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from .. import functional as F
from .._lib import MilB200Error
from ..abmil import ABMIL, ABMIL_v2
from .encoders import PrecomputedFeatures
from .sam.transformer import TwoWayTransformer, _use_tape, _use_collapsed
from ..tape import Tape

_CT_MODELS = ("resnet2plus1d_18", "resnetMC3_18", "medicalNet", "SwinUNETR", "MViT")
_CI_MODELS = ("simpleFCs_v1", "simpleFCs_v1d", "simpleFCs_v2", "simpleFCs_v2d", "CLIP")


def _make_pool(kind, args, L, where):
    if kind == "ABMIL":
        return ABMIL(args, L=L)
    if kind == "ABMIL_v2":
        # upstream calls ABMIL_v2(args, L=...) which raises TypeError; keep that behaviour explicit
        raise TypeError(f"{where}: ABMIL_v2.__init__() got an unexpected keyword argument 'L' (as upstream)")
    if kind in ("TransMIL", "TransMIL_seperate"):
        raise NotImplementedError(f"{where}='{kind}': Nystrom-attention TransMIL is outside the hot path "
                                  "(needs the un-vendored nystrom_attention package; SURVEY F5)")
    return None


class aggregator(nn.Module):
    def __init__(self, args, extractor_CT: nn.Module = None, clinic_extractor: nn.Module = None):
        super().__init__()
        self.args = args
        embedding_dim = 512
        self.max_seq_len = 100000
        self.embedding_dim = embedding_dim
        tw = dict(args=args, depth=2, embedding_dim=embedding_dim, num_heads=8, mlp_dim=2048)

        if "CT" in args.modality:                                                     # aggregator.py:17-42
            self.extractor_CT = extractor_CT if extractor_CT is not None else PrecomputedFeatures()
            self.TwoWayTransformer_CT = TwoWayTransformer(**tw)
        self.fc_CI2CT = nn.Sequential(nn.Linear(embedding_dim, embedding_dim), nn.Tanh())           # :44
        if "pathology" in args.modality:                                              # :46-64
            self.fc_pathology = nn.Sequential(nn.Linear(768, embedding_dim), nn.Tanh())
            pool = _make_pool(getattr(args, "model_pathology", None), args, embedding_dim, "model_pathology")
            if pool is not None:
                self.extractor_pathology = pool                                       # built but unused upstream too
            self.TwoWayTransformer_Pth = TwoWayTransformer(**tw)
        self.fc_CI2Pth = nn.Sequential(nn.Linear(embedding_dim, embedding_dim), nn.Tanh())          # :66
        self.fc_CI = nn.Sequential(nn.Linear(embedding_dim, embedding_dim), nn.Tanh())              # :68
        self.TwoWayTransformer_Both = TwoWayTransformer(**tw)                                       # :70-76
        pool = _make_pool(getattr(args, "aggregator", None), args, embedding_dim, "aggregator")     # :79-96
        if pool is not None:
            self.aggregator = pool
        self.clinic_extractor = clinic_extractor if clinic_extractor is not None else PrecomputedFeatures()  # :108-122
        self.prompt_embedding = nn.Parameter(torch.randn(1, embedding_dim))           # :124 (unused, kept for the ABI)
        self.fc = nn.Sequential(nn.Dropout(0.25), nn.Linear(embedding_dim, args.num_classes))      # :128-131
        self._pe_cache = {}

    # ---- sinusoidal position table (aggregator.py:99-106), device resident -----------------------------
    def _pe(self, n, like):
        key = (like.device, like.dtype)
        tab = self._pe_cache.get(key)
        if tab is None or tab.shape[1] < n:
            rows = min(self.max_seq_len, max(int(n), 4096 if tab is None else 2 * tab.shape[1]))
            if n > self.max_seq_len:
                raise MilB200Error(f"bag of {n} instances exceeds the position table ({self.max_seq_len}, aggregator.py:99)")
            tab = F.sinusoid_pe(rows, self.embedding_dim, like.dtype, like.device)
            self._pe_cache[key] = tab
        return tab[:, :n]

    @property
    def pe(self):
        """(1, 100000, 512) fp32, like the upstream attribute (generated on first access, on the device)."""
        dev = self.fc[1].weight.device
        return self._pe(self.max_seq_len, torch.empty(0, dtype=torch.float32, device=dev))

    # ---- small fused pieces ------------------------------------------------------------------------------
    @staticmethod
    def _fc_tanh(seq, x):
        return F.linear(x, seq[0].weight, seq[0].bias, act="tanh")

    def _head(self, x0):
        if self.training and self.fc[0].p > 0:
            x0 = F.dropout(x0, self.fc[0].p)
        return F.linear(x0, self.fc[1].weight, self.fc[1].bias, act="sigmoid")      # torch.sigmoid(self.fc(x0)), :200

    def _fuse(self, transformer, x_img, text_tokens):
        if x_img.dim() == 5:
            b, t, c, h, w = x_img.shape                                               # :156 (t = 512 channels, c = slices)
            n = c * h * w if getattr(self.args, "model_CT", None) == "medicalNet" else c
        else:
            n = x_img.shape[1]
        return transformer(x_img, self._pe(n, text_tokens), text_tokens)

    # ---- CT + pathology branch as ONE native program (csrc/tape.cu) ----------------------------------------
    def _fusion_tape(self, single_token=False):
        """aggregator.py:141,160,168,173 on a tape: fc_pathology, fc_CI2CT / fc_CI2Pth, both TwoWayTransformer_Both
        calls, and the four results written in place into the packed multi-modal bag (no torch.cat copy).
        single_token: the T = 1 specialisation (exact, see Attention.emit_single_key)."""
        name = "_tape_cache_t1" if single_token else "_tape_cache"
        t = getattr(self, name, None)
        if t is None:
            t = Tape()
            E = self.embedding_dim
            ct, pe_ct = t.input("Nc", E), t.input("Nc", E)
            xp, pe_p = t.input("Np", 768), t.input("Np", E)
            txt = t.input("T", E)
            xin_p = t.linear(xp, self.fc_pathology[0], act="tanh")                                   # :141
            # the CT branch (160 tokens: ~180 tiny kernels) is independent of the pathology branch until the bag is
            # assembled: lane 1 runs it on a second stream / as a parallel branch of the replayed graph
            with t.lane(1):
                q1, k1 = self.TwoWayTransformer_Both.emit(t, ct, pe_ct, t.linear(txt, self.fc_CI2CT[0], act="tanh"),
                                                          single_token=single_token)                # :160
            q2, k2 = self.TwoWayTransformer_Both.emit(t, xin_p, pe_p, t.linear(txt, self.fc_CI2Pth[0], act="tanh"),
                                                      single_token=single_token)                    # :168
            bag = t.buffer(lambda r: 2 * r["T"] + r["Nc"] + r["Np"], E)                             # :173 row order
            t.output(q1, bag, lambda r: 0)
            t.output(k1, bag, lambda r: r["T"])
            t.output(q2, bag, lambda r: r["T"] + r["Nc"])
            t.output(k2, bag, lambda r: 2 * r["T"] + r["Nc"])
            object.__setattr__(self, name, t)
        return t

    # ---- the same branch with the image-side projections folded into the token side (T = 1; csrc/xfusion.cu) ------------
    def _fusion_tape_v2(self):
        """fc_pathology, fc_CI2CT / fc_CI2Pth and BOTH TwoWayTransformer_Both calls as ONE segmented program: the CT bag and
        the pathology bag (of one or several patients) are segments of one key stream; the token side (one row per
        segment) runs in fp32; the final keys and token rows land in the packed multi-modal bag (aggregator.py:173)."""
        t = getattr(self, "_tape_cache_v2", None)
        if t is None:
            t = Tape()
            # MILB200_FUSION_STREAM=bf16 (experiment switch): the key stream and the CT tokens in the program dtype too
            stream_f32 = os.environ.get("MILB200_FUSION_STREAM", "f32") != "bf16"
            t.stream_f32 = stream_f32
            E = self.embedding_dim
            # Storage: the patch features (the big HBM stream) and the packed bag are in the program dtype; the key stream
            # between them, the CT tokens, the position table and the whole token side are fp32 — in a bf16 program the
            # only reduced-precision steps are then the tensor-core operands of fc_pathology and of the gated pool
            xp, ct = t.input("NP", 768), t.input("NC", E, f32=stream_f32)
            pe = t.input("NPE", E, f32=True)
            txt = t.input("BT", E, f32=True)
            keys = t.join(t.linear(xp, self.fc_pathology[0], act="tanh", out_f32=stream_f32), ct, "NK")    # :141 | CT tokens
            points = t.join(t.linear(txt, self.fc_CI2CT[0], act="tanh"),                             # :160 third argument
                            t.linear(txt, self.fc_CI2Pth[0], act="tanh"), "ST")                      # :168 third argument
            q, k = self.TwoWayTransformer_Both.emit_collapsed(t, keys, pe, points)
            full = t.tok_scatter(q, k)
            bag = t.buffer(lambda r: r["NBAG"], E)
            t.output(k, bag, lambda r: 0)
            t.output(full, bag, lambda r: 0)
            # the final token rows also leave in fp32 (rows: CT segments, then pathology segments): x_CT2CI / x_Pth2CI and
            # the cosine loss on them never see the packed bag's storage dtype
            t.output(q, t.buffer(lambda r: r["ST"], E, f32=True), lambda r: 0)
            object.__setattr__(self, "_tape_cache_v2", t)
        return t

    @staticmethod
    def fusion_layout(n_ct, n_path_list, T=1):
        """Row bookkeeping of the segmented program for B patients with n_ct CT tokens each and n_path_list[b] pathology
        rows: (rows dict, segment table, per-patient bag offsets).  Bag b holds [T | n_ct | T | n_path_b] rows in the order
        of aggregator.py:173; segments are ordered CT(0..B-1), pathology(0..B-1)."""
        B = len(n_path_list)
        n_p = int(sum(n_path_list))
        bag_off = [0]
        for n in n_path_list:
            bag_off.append(bag_off[-1] + 2 * T + n_ct + int(n))
        segs, pstart = [], 0
        for b in range(B):
            segs.append((n_p + b * n_ct, n_ct, bag_off[b] + T, bag_off[b]))
        for b, n in enumerate(n_path_list):
            segs.append((pstart, int(n), bag_off[b] + 2 * T + n_ct, bag_off[b] + T + n_ct))
            pstart += int(n)
        rows = {"NP": n_p, "NC": B * n_ct, "NK": n_p + B * n_ct, "BT": B * T, "ST": 2 * B * T, "SJ": 2 * B * T * 8,
                "NBAG": bag_off[-1]}
        return rows, (tuple(segs), T), bag_off

    def _pe_table(self, n, device):
        """The cached fp32 position table (whole: its address never changes between calls), at least n rows."""
        like = torch.empty(0, dtype=torch.float32, device=device)
        self._pe(n, like)
        return self._pe_cache[(like.device, like.dtype)][0]

    def _forward_fused_v2(self, x_ct_tokens, x_path, x_text):
        Nc, Np = x_ct_tokens.shape[1], x_path.shape[1]
        rows, segs, _ = self.fusion_layout(Nc, [Np], 1)
        pe = self._pe_table(max(Nc, Np), x_path.device)
        rows["NPE"] = pe.shape[0]
        tape = self._fusion_tape_v2()
        ct_in = x_ct_tokens[0].float() if tape.stream_f32 else x_ct_tokens[0]
        bag, tok = tape.run(rows, [x_path[0], ct_in, pe, x_text[0].float()], segs=segs)
        return bag.unsqueeze(0), tok[0:1].unsqueeze(0), tok[1:2].unsqueeze(0)          # x0, x_CT2CI, x_Pth2CI (fp32)

    def forward_bags(self, ct_tokens, x_path, path_lens, x_text):
        """The CT+pathology branch for B patients in ONE launch set (B200-native entry; the reference runs batch 1,
        train_ddp.py:75).  ct_tokens (B, Nc, 512): per-slice CT tokens (F.ct_tokens of the encoder's feature map);
        x_path (sum Np, 768): the patients' patch features packed row-wise; path_lens: their row counts (host ints);
        x_text (B, 1, 512): one clinical-text embedding per patient.  Returns (prob (B, C), x_CT2CI (B, 1, 512),
        x_Pth2CI (B, 1, 512)), all fp32 — row b equals forward([ct_b, path_b], text_b)."""
        B, Nc = int(ct_tokens.shape[0]), int(ct_tokens.shape[1])
        path_lens = [int(n) for n in path_lens]
        if len(path_lens) != B or x_text.shape[0] != B or x_text.shape[1] != 1 or 2 * B > 16:
            raise MilB200Error("forward_bags: B patients (<= 8), one text token each, len(path_lens) == B")
        if x_path.dim() != 2 or x_path.shape[0] != sum(path_lens) or min(path_lens) < 2 or Nc < 2:
            raise MilB200Error("forward_bags: x_path must be the packed (sum Np, 768) matrix; bags need >= 2 rows")
        rows, segs, bag_off = self.fusion_layout(Nc, path_lens, 1)
        pe = self._pe_table(max(Nc, max(path_lens)), x_path.device)
        rows["NPE"] = pe.shape[0]
        E = self.embedding_dim
        tape = self._fusion_tape_v2()
        ct_in = ct_tokens.reshape(B * Nc, E)
        bag, tok = tape.run(rows, [x_path, ct_in.float() if tape.stream_f32 else ct_in, pe, x_text.reshape(B, E).float()],
                            segs=segs)
        key = (tuple(bag_off), bag.device)
        cached = self.__dict__.setdefault("_bag_off_cache", {})
        if key not in cached:
            if len(cached) > 64:
                cached.clear()
            cached[key] = torch.tensor(bag_off, dtype=torch.int32, device=bag.device)
        prob = self._head(self.aggregator.forward_csr(bag, cached[key], out_fp32=True))           # :199-200
        return prob, tok[:B].unsqueeze(1), tok[B:].unsqueeze(1)

    def _forward_fused(self, x_ct_tokens, x_path, x_text):
        T, Nc, Np = x_text.shape[1], x_ct_tokens.shape[1], x_path.shape[1]
        like = x_text
        (bag,) = self._fusion_tape(single_token=(T == 1 and Nc > 1 and Np > 1)).run({"T": T, "Nc": Nc, "Np": Np},
                                         [x_ct_tokens[0], self._pe(Nc, like)[0], x_path[0], self._pe(Np, like)[0], x_text[0]])
        x0 = bag.unsqueeze(0)
        return x0, x0[:, :T], x0[:, T + Nc:2 * T + Nc]

    def forward(self, x_list, x_CI):
        mod = self.args.modality
        has_ct, has_path = "CT" in mod, "pathology" in mod
        x_input_CT = x_input_pathology = None
        if (has_ct and has_path and _use_tape() and getattr(self.args, "model_CT", None) == "resnetMC3_18"
                and getattr(self.args, "alignment_base", None) != "CT" and getattr(self.args, "aggregator", "-") != "-"):
            x_ct = self.extractor_CT(x_list[0])
            x_text = self.clinic_extractor(x_CI)
            if x_ct.dim() == 5 and x_ct.shape[0] == 1 and x_list[1].dim() == 3 and x_list[1].shape[0] == 1 \
                    and x_text.dim() == 3 and x_text.shape[0] == 1 and x_ct.dtype == x_list[1].dtype == x_text.dtype:
                if x_text.shape[1] == 1 and x_ct.shape[2] > 1 and x_list[1].shape[1] > 1 and _use_collapsed():
                    x0, x_CT2CI, x_Pth2CI = self._forward_fused_v2(F.ct_tokens(x_ct), x_list[1], x_text)
                    # pooled vector and head in fp32: with bf16 bags the storage of the bag is the only low-precision step
                    off = torch.arange(0, 2, dtype=torch.int32, device=x0.device) * x0.shape[1]
                    return self._head(self.aggregator.forward_csr(x0[0], off, out_fp32=True)), x_CT2CI, x_Pth2CI
                x0, x_CT2CI, x_Pth2CI = self._forward_fused(F.ct_tokens(x_ct), x_list[1], x_text)
                return self._head(self.aggregator(x0)), x_CT2CI, x_Pth2CI
        if has_ct:
            x_input_CT = self.extractor_CT(x_list[0])                                 # :140,146
        if has_path:
            x_input_pathology = self._fc_tanh(self.fc_pathology, x_list[1 if has_ct else 0])      # :141,149
        x_CI_prompted = self.clinic_extractor(x_CI)                                   # :151 -> (1, T, 512)

        if has_ct and has_path:
            x_CT2CI, x_CI2CT = self._fuse(self.TwoWayTransformer_Both, x_input_CT,
                                          self._fc_tanh(self.fc_CI2CT, x_CI_prompted))            # :160
            x_Pth2CI, x_CI2Pth = self._fuse(self.TwoWayTransformer_Both, x_input_pathology,
                                            self._fc_tanh(self.fc_CI2Pth, x_CI_prompted))         # :168
            x0 = torch.cat([x_CT2CI, x_CI2CT, x_Pth2CI, x_CI2Pth], dim=1)             # :173 the multi-modal bag
        elif has_ct:
            # upstream reads an undefined name here; this is the evident intent
            x_CT2CI, x_CI2CT = self._fuse(self.TwoWayTransformer_CT, x_input_CT,
                                          self._fc_tanh(self.fc_CI2CT, x_CI_prompted))            # :179
            x0 = torch.cat([x_CT2CI, x_CI2CT], dim=1)
        elif has_path:
            x_Pth2CI, x_CI2Pth = self._fuse(self.TwoWayTransformer_Pth, x_input_pathology,
                                            self._fc_tanh(self.fc_CI2Pth, x_CI_prompted))         # :190
            x0 = torch.cat([x_Pth2CI, x_CI2Pth], dim=1)
        elif "CI" in mod:
            x0 = self._fc_tanh(self.fc_CI, x_CI_prompted)                             # :195
        else:
            raise MilB200Error(f"aggregator: unsupported modality {mod}")

        if getattr(self.args, "aggregator", "-") != "-":
            x0 = self.aggregator(x0)                                                  # :199 gated-attention MIL pool
        x = self._head(x0)                                                            # :200
        if has_ct and has_path:
            return x, x_CT2CI, x_Pth2CI
        if has_ct:
            return x, x_CT2CI
        if has_path:
            return x, x_Pth2CI
        return x
