"""Native training step of the CT + pathology `aggregator` (BASELINE configs[2]; train_ddp.py:297-348 with batch size 1):

    prob, x_CT2CI, x_Pth2CI = model([x_ct, x_path], x_text)
    loss = BCELoss(prob, label) + CosineEmbeddingLoss(x_CT2CI, x_Pth2CI, +1)
    loss.backward(); all-reduce(grads); Adam / SGD

without the autograd graph: every parameter of the branch lives in ONE flat fp32 buffer (fusion-tape parameters | gated
pool | head), the backward kernels write into ONE flat gradient buffer, DDP's exchange is ONE all-reduce and the
optimiser ONE kernel — the same design as `dp.AbmilTrainer`, which this class embeds for the gated pool.  Through
nn.Module + torch.autograd the step is bound by PyTorch's per-node host cost (~100 AccumulateGrad nodes, see
tools/host_segments.py); here the host issues ~15 C calls.  `load_from` / `store_to` move parameters from / to the
reference-shaped module, so checkpoints keep the reference's state_dict.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L
from .dp import AbmilTrainer


def _ptrs(n):
    return (C.c_void_p * n)()


class FusionTrainer:
    def __init__(self, model, n_text_tokens=1, compute_dtype=torch.bfloat16, lr=1e-5, betas=(0.9, 0.999), eps=1e-8,
                 weight_decay=1e-7, optimizer="adam", process_group=None, world_size=1, cosine_loss=True):
        if not hasattr(model, "_fusion_tape") or not hasattr(model, "aggregator"):
            raise L.MilB200Error("FusionTrainer needs the CT+pathology `aggregator` with a gated-attention pool")
        if optimizer not in ("adam", "sgd"):
            raise ValueError("optimizer must be 'adam' or 'sgd'")
        self.model, self.T = model, int(n_text_tokens)
        self.dtype = compute_dtype
        self.device = model.fc[1].weight.device
        self.lr, self.betas, self.eps, self.wd, self.optimizer = lr, betas, eps, weight_decay, optimizer
        self.pg, self.world = process_group, world_size
        self.cosine_loss = cosine_loss
        self.E = model.embedding_dim
        self.C = model.fc[1].weight.shape[0]
        self.tape = model._fusion_tape(single_token=(self.T == 1))
        c = self.tape._freeze()
        self.n_tape = c["total"]
        D = model.aggregator.attention_V[0].weight.shape[0]
        self.n_pool = 2 * D * self.E + 3 * D + 1
        up8 = lambda n: (n + 7) // 8 * 8
        self.o_pool = up8(self.n_tape)
        self.o_head_w = up8(self.o_pool + self.n_pool)
        self.o_head_b = up8(self.o_head_w + self.C * self.E)
        self.numel = up8(self.o_head_b + self.C)
        z = lambda: torch.zeros(self.numel, dtype=torch.float32, device=self.device)
        self.params, self.grads, self.exp_avg, self.exp_avg_sq = z(), z(), z(), z()
        # the gated pool runs through AbmilTrainer on views of the same flat buffers (its instances are the tape's output,
        # so it also returns the gradient of the packed bag)
        self.pool = AbmilTrainer(self.E, D, compute_dtype, device=self.device, need_input_grad=True)
        sl = slice(self.o_pool, self.o_pool + self.n_pool)
        self.pool.params, self.pool.grads = self.params[sl], self.grads[sl]
        self.wc = torch.empty(self.n_tape, dtype=compute_dtype, device=self.device) if compute_dtype != torch.float32 else None
        self.step_count = 0
        self._buf = {}
        self.load_from(model)

    # ---- parameters <-> module ----------------------------------------------------------------------------------
    def _module_tensors(self):
        """(flat offset, tensor) for every parameter outside the pool, in flat-buffer order."""
        c = self.tape._freeze()
        out = [(off, p) for p, off in zip(self.tape.params, c["offsets"])]
        out.append((self.o_head_w, self.model.fc[1].weight))
        out.append((self.o_head_b, self.model.fc[1].bias))
        return out

    @torch.no_grad()
    def load_from(self, model=None):
        for off, p in self._module_tensors():
            self.params[off:off + p.numel()].copy_(p.detach().reshape(-1))
        self.pool.load_from(self.model.aggregator)
        self._refresh_compute_copy()

    @torch.no_grad()
    def store_to(self, model=None):
        for off, p in self._module_tensors():
            p.copy_(self.params[off:off + p.numel()].view(p.shape))
        self.pool.store_to(self.model.aggregator)

    def named_grads(self):
        """{state_dict name: gradient view} for the tests (tape / head parameters by module name; pool under aggregator.*)."""
        names = {id(p): n for n, p in self.model.named_parameters()}
        out = {names[id(p)]: self.grads[off:off + p.numel()].view(p.shape) for off, p in self._module_tensors()}
        gv = self.pool.grad_views()
        D = self.pool.D
        out["aggregator.attention_V.0.weight"], out["aggregator.attention_U.0.weight"] = gv["Wcat"][:D], gv["Wcat"][D:]
        out["aggregator.attention_V.0.bias"], out["aggregator.attention_U.0.bias"] = gv["bcat"][:D], gv["bcat"][D:]
        out["aggregator.attention_weights.weight"] = gv["ww"].view(1, D)
        out["aggregator.attention_weights.bias"] = gv["bw"]
        return out

    def _refresh_compute_copy(self):
        if self.wc is not None:
            L.check(L.lib().milb200_cast(L.ptr(self.params), L.F32, L.ptr(self.wc), L.dtype_code(self.wc), self.n_tape,
                                         L.stream_ptr()), "cast")

    def broadcast_params(self):
        if self.world > 1:
            torch.distributed.broadcast(self.params, src=0, group=self.pg)
            self._refresh_compute_copy()

    # ---- buffers of one bag shape (pointer-stable, so the tape replays as a CUDA graph) ----------------------------
    def _buffers(self, rows):
        key = tuple(sorted(rows.items()))
        b = self._buf.get(key)
        if b is None:
            if len(self._buf) >= 4:
                self._buf.pop(next(iter(self._buf)))
            t, c, lib = self.tape, self.tape._freeze(), L.lib()
            slots = t._slots(rows)
            code = L.BF16 if self.dtype == torch.bfloat16 else L.F32
            mk = lambda r, cc: torch.empty((int(r), int(cc)), dtype=self.dtype, device=self.device)
            b = dict(slots=slots, code=code,
                     arena=torch.empty(lib.milb200_tape_arena_bytes(c["ops"], c["n_ops"], slots, c["n_slots"], code),
                                       dtype=torch.uint8, device=self.device),
                     out=[mk(fn(rows), cols) for fn, cols in t.buffers],
                     gout=[mk(slots[s].rows, slots[s].cols) for s, _, _ in t.outputs],
                     stage=[mk(slots[s].rows, slots[s].cols) for s in t.inputs],
                     z=torch.empty((1, self.C), dtype=torch.float32, device=self.device),
                     prob=torch.empty((1, self.C), dtype=torch.float32, device=self.device),
                     dz=torch.empty((1, self.C), dtype=torch.float32, device=self.device),
                     dM=torch.empty((1, self.E), dtype=torch.float32, device=self.device),
                     loss=torch.zeros(2, dtype=torch.float32, device=self.device),
                     cos_rows=torch.empty(max(self.T, 1), dtype=torch.float32, device=self.device),
                     da=mk(self.T, self.E), db=mk(self.T, self.E),
                     offsets=torch.tensor([0, int(t.buffers[0][0](rows))], dtype=torch.int32, device=self.device))
            self._buf[key] = b
        return b

    # ---- one training step ---------------------------------------------------------------------------------------
    def forward_backward(self, ct_tokens, x_path, x_text, label):
        """ct_tokens (Nc, E): per-slice CT tokens (F.ct_tokens of the encoder's feature map); x_path (Np, 768) patch
        features; x_text (T, E) clinical-text embeddings; label (C,) one-hot float.  All on the device, compute dtype.
        Leaves the gradients in `self.grads`; returns (loss tensor [2] = (BCE, cosine), prob (1, C))."""
        m, t, lib = self.model, self.tape, L.lib()
        T, Nc, Np = x_text.shape[0], ct_tokens.shape[0], x_path.shape[0]
        if T != self.T:
            raise L.MilB200Error(f"FusionTrainer was built for {self.T} text token(s), got {T}")
        rows = {"T": T, "Nc": Nc, "Np": Np}
        b = self._buffers(rows)
        c = t._freeze()
        slots, code, n_slots = b["slots"], b["code"], c["n_slots"]
        like = x_text
        srcs = [ct_tokens, m._pe(Nc, like)[0], x_path, m._pe(Np, like)[0], x_text]
        inputs = []
        for j, (src, st) in enumerate(zip(srcs, b["stage"])):
            if j in (1, 3):                      # cached position tables: stable address, used in place
                inputs.append(src.contiguous())
            else:
                st.copy_(src)
                inputs.append(st)
        wcomp = self.wc if self.wc is not None else self.params
        esz = inputs[0].element_size()
        ext = _ptrs(n_slots)
        for s, x in zip(t.inputs, inputs):
            ext[s] = x.data_ptr()
        for s, bi, fn in t.outputs:
            ext[s] = b["out"][bi].data_ptr() + int(fn(rows)) * t.slot_cols[s] * esz
        ws = L.workspace(lib.milb200_tape_workspace_bytes(c["ops"], c["n_ops"], slots, n_slots, code, 0), self.device)
        L.check(lib.milb200_tape_forward(c["ops"], c["n_ops"], slots, n_slots, c["params"], c["n_params"], ext, L.ptr(wcomp),
                                         L.ptr(self.params), L.ptr(b["arena"]), b["arena"].numel(), L.ptr(ws), ws.numel(),
                                         code, L.stream_ptr()), "tape_forward")
        bag = b["out"][0]                                                     # aggregator.py:173 row order
        # gated pool over the multi-modal bag (aggregator.py:199) and the head (:200) + BCE (train_ddp.py:99,319)
        M = self.pool.forward(bag, b["offsets"])                              # (1, E) fp32
        hw = self.params[self.o_head_w:self.o_head_w + self.C * self.E]
        hb = self.params[self.o_head_b:self.o_head_b + self.C]
        lws = L.workspace(lib.milb200_linear_workspace_bytes(1, self.C, self.E, L.F32, 1), self.device)
        L.check(lib.milb200_linear_fwd(L.ptr(M), None, L.ptr(hw), L.ptr(hb), L.ptr(b["z"]), 1, self.C, self.E, L.ACT_NONE,
                                       L.F32, L.ptr(lws), lws.numel(), L.stream_ptr()), "head linear_fwd")
        tgt = label.reshape(1, self.C).to(torch.float32).contiguous()
        L.check(lib.milb200_sigmoid_bce_fwd_bwd(L.ptr(b["z"]), L.ptr(tgt), L.ptr(b["prob"]), L.ptr(b["loss"][0:1]),
                                                L.ptr(b["dz"]), self.C, L.stream_ptr()), "sigmoid_bce")
        gw = self.grads[self.o_head_w:self.o_head_w + self.C * self.E]
        gb = self.grads[self.o_head_b:self.o_head_b + self.C]
        L.check(lib.milb200_linear_bwd(L.ptr(M), None, L.ptr(hw), L.ptr(b["z"]), L.ptr(b["dz"]), L.ptr(b["dM"]), L.ptr(gw),
                                       L.ptr(gb), 1, self.C, self.E, L.ACT_NONE, L.F32, 0, L.ptr(lws), lws.numel(),
                                       L.stream_ptr()), "head linear_bwd")
        dbag = self.pool.backward(b["dM"])                                    # (2T + Nc + Np, E), compute dtype
        if self.cosine_loss:
            a_rows, b_rows = bag[0:T], bag[T + Nc:2 * T + Nc]                 # x_CT2CI, x_Pth2CI (aggregator.py:160,168)
            L.check(lib.milb200_cosine_embedding_fwd_bwd(L.ptr(a_rows), L.ptr(b_rows), L.ptr(b["loss"][1:2]),
                                                         L.ptr(b["cos_rows"]), L.ptr(b["da"]), L.ptr(b["db"]), T, self.E,
                                                         code, L.stream_ptr()), "cosine_embedding")
            for rows_view, g in ((dbag[0:T], b["da"]), (dbag[T + Nc:2 * T + Nc], b["db"])):     # both losses reach these rows
                L.check(lib.milb200_add(L.ptr(rows_view), L.ptr(g), L.ptr(rows_view), T * self.E, code, L.stream_ptr()), "add")
        # backward of the fusion program, seeded with the gradient of the packed bag
        ext2, gext, seeds = _ptrs(n_slots), _ptrs(n_slots), _ptrs(n_slots)
        for s, x in zip(t.inputs, inputs):
            ext2[s] = x.data_ptr()
        for i, (s, bi, fn) in enumerate(t.outputs):
            off = int(fn(rows)) * t.slot_cols[s] * esz
            ext2[s] = b["out"][bi].data_ptr() + off
            seeds[s] = dbag.data_ptr() + off
            gext[s] = b["gout"][i].data_ptr()
        ws = L.workspace(lib.milb200_tape_workspace_bytes(c["ops"], c["n_ops"], slots, n_slots, code, 1), self.device)
        L.check(lib.milb200_tape_backward(c["ops"], c["n_ops"], slots, n_slots, c["params"], c["n_params"], ext2, gext, seeds,
                                          L.ptr(wcomp), L.ptr(self.params), L.ptr(self.grads), L.ptr(b["arena"]),
                                          b["arena"].numel(), L.ptr(ws), ws.numel(), code, L.stream_ptr()), "tape_backward")
        return b["loss"], b["prob"]

    def reduce_and_update(self):
        """ONE all-reduce(sum) of the flat gradient, then the fused optimiser step with grad_scale = 1/world."""
        if self.world > 1:
            torch.distributed.all_reduce(self.grads, group=self.pg)
        self.step_count += 1
        if self.optimizer == "sgd":
            L.check(L.lib().milb200_sgd_step(L.ptr(self.params), L.ptr(self.grads), self.numel, self.lr, self.wd,
                                             1.0 / self.world, L.stream_ptr()), "sgd_step")
        else:
            L.check(L.lib().milb200_adam_step(L.ptr(self.params), L.ptr(self.grads), L.ptr(self.exp_avg),
                                              L.ptr(self.exp_avg_sq), self.numel, self.lr, self.betas[0], self.betas[1],
                                              self.eps, self.wd, 1.0 / self.world, self.step_count, L.stream_ptr()),
                    "adam_step")
        self._refresh_compute_copy()

    def step(self, ct_tokens, x_path, x_text, label):
        out = self.forward_backward(ct_tokens, x_path, x_text, label)
        self.reduce_and_update()
        return out
