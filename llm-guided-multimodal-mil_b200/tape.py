"""Static operator programs ("tapes"): host side of csrc/tape.cu.

A module describes its forward ONCE as ops over numbered tensor slots (``Tape.linear / attention / layernorm``); a call
then costs one ``milb200_tape_forward`` and, in backward, one ``milb200_tape_backward`` instead of one autograd node,
several allocations and a ctypes call per kernel.  Row counts are symbolic (``"T"``, ``"N"`` ...) and bound per call,
so one tape serves every bag size.  The whole program is a single ``torch.autograd.Function``; parameter gradients come
back as views of one flat fp32 buffer.
"""
from __future__ import annotations

import ctypes as C
import os
import weakref
from collections import OrderedDict

import torch

from . import _lib as L
from . import functional as F

_ACT = {None: L.ACT_NONE, "none": L.ACT_NONE, "tanh": L.ACT_TANH, "relu": L.ACT_RELU, "sigmoid": L.ACT_SIGMOID}


class Tape:
    def __init__(self):
        self.ops = []            # (kind, in0, in1, in2, out, p0, p1, a0, lane)
        self._lane = 0           # ops emitted now go to this lane (0 or 1); lanes run concurrently on two streams
        self.slot_rows = []      # symbolic row key per slot
        self.slot_cols = []
        self.slot_ext = []       # 0 internal, 1 external input, 2 external output (lives in an output buffer)
        self.slot_f32 = []       # True: the slot is fp32 whatever the program dtype (the text-token side)
        self.params = []         # nn.Parameter objects, in flat-buffer order
        self._pidx = {}
        self.inputs = []         # slot ids of external inputs, in call order
        self.outputs = []        # (slot, buffer index, row-offset function(rows) -> int)
        self.buffers = []        # (rows function(rows) -> int, cols)
        self._c = None

    # ---- building ------------------------------------------------------------------------------------------
    def slot(self, rows_key, cols, ext=0, f32=False):
        self.slot_rows.append(rows_key)
        self.slot_cols.append(int(cols))
        self.slot_ext.append(ext)
        self.slot_f32.append(bool(f32))
        return len(self.slot_cols) - 1

    def input(self, rows_key, cols, f32=False):
        s = self.slot(rows_key, cols, ext=1, f32=f32)
        self.inputs.append(s)
        return s

    def param(self, p):
        if p is None:
            return -1
        k = id(p)
        if k not in self._pidx:
            self._pidx[k] = len(self.params)
            self.params.append(p)
        return self._pidx[k]

    def lane(self, n):
        """Context manager: ops emitted inside run on lane `n` (0 or 1).  Lanes are executed on two streams (parallel
        branches of the replayed CUDA graph); the executor orders every cross-lane use of a slot, a slot gradient or a
        parameter gradient, so ANY assignment is correct — it only pays off for independent sub-programs such as the
        CT and pathology branches of aggregator.py:160,168."""
        tape = self

        class _Lane:
            def __enter__(self_inner):
                self_inner.prev = tape._lane
                tape._lane = int(n)

            def __exit__(self_inner, *exc):
                tape._lane = self_inner.prev
        if n not in (0, 1):
            raise L.MilB200Error("tape.lane: two lanes are available (0, 1)")
        if os.environ.get("MILB200_TAPE_LANES", "1") == "0":      # switch: keep the whole program on one stream
            n = 0
        return _Lane()

    def linear(self, x, lin, act=None, add=None, out_f32=False):
        """out = act((x [+ add]) W^T + b) for an ``nn.Linear`` parameter container.  out_f32: the result is an fp32 slot even
        when x is stored in the program dtype (bf16 tensor-core operands, fp32 result: the head of an fp32 key stream)."""
        out = self.slot(self.slot_rows[x], lin.weight.shape[0], f32=self.slot_f32[x] or out_f32)
        self.ops.append((L.OP_LINEAR, x, -1 if add is None else add, -1, out, self.param(lin.weight), self.param(lin.bias),
                         _ACT[act], self._lane))
        return out

    def attention(self, q, k, v, heads):
        out = self.slot(self.slot_rows[q], self.slot_cols[q])
        self.ops.append((L.OP_ATTENTION, q, k, v, out, -1, -1, int(heads), self._lane))
        return out

    def layernorm(self, x, ln, residual=None):
        if abs(ln.eps - 1e-5) > 1e-12:
            raise L.MilB200Error("layernorm kernels are built for eps = 1e-5 (nn.LayerNorm default)")
        out = self.slot(self.slot_rows[x], self.slot_cols[x], f32=self.slot_f32[x])
        self.ops.append((L.OP_LAYERNORM, x, -1 if residual is None else residual, -1, out, self.param(ln.weight),
                         self.param(ln.bias), 0, self._lane))
        return out

    def add(self, a, b):
        """out = a + b as a slot of its own, for sums that several ops consume (keys + key_pe feeds two projections
        per block: materialise it once instead of once per consumer)."""
        out = self.slot(self.slot_rows[a], self.slot_cols[a], f32=self.slot_f32[a])
        self.ops.append((L.OP_ADD, a, b, -1, out, -1, -1, 0, self._lane))
        return out

    # ---- segment ops of the collapsed cross-modal attention (csrc/xfusion.cu) ---------------------------------------
    def join(self, a, b, rows_key):
        """out = rows of a, then rows of b (`rows_key` names the total).  An internal `a` is produced in place at the head
        of the result (no copy); `b` is copied."""
        out = self.slot(rows_key, self.slot_cols[a], f32=self.slot_f32[a])
        self.ops.append((L.OP_JOIN, a, b, -1, out, -1, -1, 0, self._lane))
        return out

    def headdiag_u(self, x, lin, rows_key):
        """out[r*8 + h, :] = sum_c x[r, h*32 + c] W[h*32 + c, :] — the per-head image of the projected queries under the
        key projection (U = Wk_h^T q_h).  x [R, 256] fp32, `lin` the k_proj container ([256, 512]); rows_key names R*8."""
        out = self.slot(rows_key, lin.weight.shape[1], f32=True)
        self.ops.append((L.OP_HEADDIAG_U, x, -1, -1, out, self.param(lin.weight), -1, 8, self._lane))
        return out

    def headdiag_o(self, y, lin, rows_key):
        """out[r, h*32 + c] = y[r*8 + h, :] . W[h*32 + c, :] + b — the value projection applied to the per-head pooled
        keys.  y [R*8, 512] fp32, `lin` the v_proj container; rows_key names R."""
        out = self.slot(rows_key, lin.weight.shape[0], f32=True)
        self.ops.append((L.OP_HEADDIAG_O, y, -1, -1, out, self.param(lin.weight), self.param(lin.bias), 8, self._lane))
        return out

    def t2i_pool(self, keys, pe, u, bag_layout=False):
        """Pool[(seg, t, h), :] = sum_n softmax_n(scale (keys + pe) . U[(seg, t, h)]) keys[n] over each segment's rows."""
        out = self.slot(self.slot_rows[u], self.slot_cols[u], f32=True)
        self.ops.append((L.OP_T2I_POOL, keys, pe, u, out, -1, -1, 1 if bag_layout else 0, self._lane))
        return out

    def ln_seg(self, keys, rows, ln, out_rows_key=None):
        """out = LN(keys + rows[segment]) (one residual row per segment).  out_rows_key: the result is written in the
        packed-bag layout (segment out_start) into a slot of that many rows."""
        if abs(ln.eps - 1e-5) > 1e-12:
            raise L.MilB200Error("layernorm kernels are built for eps = 1e-5 (nn.LayerNorm default)")
        bag = out_rows_key is not None
        # the packed bag is stored in the program dtype (it feeds the gated pool's tensor-core GEMMs); inside the stream the
        # result keeps the storage of its input
        out = self.slot(out_rows_key if bag else self.slot_rows[keys], self.slot_cols[keys],
                        f32=False if bag else self.slot_f32[keys])
        self.ops.append((L.OP_LN_SEG, keys, rows, -1, out, self.param(ln.weight), self.param(ln.bias), 1 if bag else 0,
                         self._lane))
        return out

    def tok_scatter(self, tokens, bag):
        """The token rows [n_segs*T, 512] written into their rows (segment tok_row) of the packed bag, in place: the result
        names the same external memory as `bag` (declare both with Tape.output at the same buffer offset)."""
        out = self.slot(self.slot_rows[bag], self.slot_cols[bag], f32=self.slot_f32[bag])
        self.ops.append((L.OP_TOK_SCATTER, tokens, bag, -1, out, -1, -1, 0, self._lane))
        return out

    def buffer(self, rows_fn, cols, f32=False):
        """An output buffer the caller receives; f32: fp32 whatever the program dtype (holds fp32 slots)."""
        self.buffers.append((rows_fn, int(cols)))
        self.__dict__.setdefault("buffer_f32", []).append(bool(f32))
        return len(self.buffers) - 1

    def buffer_dtype(self, i, dtype):
        return torch.float32 if self.__dict__.get("buffer_f32", [False] * len(self.buffers))[i] else dtype

    def output(self, slot, buf, row_offset_fn):
        """Have `slot` written in place at row `row_offset_fn(rows)` of output buffer `buf`."""
        if self.slot_ext[slot] != 0:
            raise L.MilB200Error("tape.output: slot is already external")
        if self.slot_cols[slot] != self.buffers[buf][1]:
            raise L.MilB200Error("tape.output: column count differs from the buffer's")
        if self.slot_f32[slot] != (self.buffer_dtype(buf, None) == torch.float32):
            raise L.MilB200Error("tape.output: slot and buffer storage differ")
        self.slot_ext[slot] = 2
        self.outputs.append((slot, buf, row_offset_fn))

    # ---- frozen C image ------------------------------------------------------------------------------------
    def _freeze(self):
        if self._c is not None:
            return self._c
        n_ops, n_slots, n_params = len(self.ops), len(self.slot_cols), len(self.params)
        ops = (L.TapeOp * n_ops)()
        for i, o in enumerate(self.ops):
            ops[i] = L.TapeOp(*o)
        params = (L.TapeParam * max(n_params, 1))()
        off = 0
        offsets = []
        for i, p in enumerate(self.params):
            if off % 8:
                raise L.MilB200Error("tape: parameter sizes must be multiples of 8 elements (16-byte aligned ranges)")
            rows, cols = (p.shape[0], p.shape[1]) if p.dim() == 2 else (1, p.numel())
            params[i] = L.TapeParam(off, rows, cols)
            offsets.append(off)
            off += p.numel()
        self._c = dict(ops=ops, n_ops=n_ops, n_slots=n_slots, params=params, n_params=n_params, offsets=offsets, total=off,
                       sizes=[p.numel() for p in self.params],
                       shapes=[None if p.dim() == 1 else tuple(p.shape) for p in self.params])
        return self._c

    def _slots(self, rows):
        key = tuple(sorted(rows.items()))
        cache = self.__dict__.setdefault("_slots_cache", {})
        arr = cache.get(key)
        if arr is None or self._c is None:       # the slot table is final once the program is frozen
            n = len(self.slot_cols)
            arr = (L.TapeSlot * n)()
            for i in range(n):
                arr[i] = L.TapeSlot(int(rows[self.slot_rows[i]]), self.slot_cols[i],
                                    (L.SLOT_EXTERNAL if self.slot_ext[i] else 0) | (L.SLOT_F32 if self.slot_f32[i] else 0))
            if self._c is not None:
                if len(cache) > 64:
                    cache.clear()
                cache[key] = arr
        return arr

    # ---- flat parameter images (cached while the parameters are unchanged) ------------------------------------
    def _flat(self, compute_dtype):
        key = (compute_dtype,) + tuple((p.data_ptr(), p._version) for p in self.params)
        cache = getattr(self, "_flat_cache", None)
        if cache is not None and cache[0] == key:
            return cache[1], cache[2]
        native = torch.cat([p.detach().reshape(-1) for p in self.params])      # gather copy (data movement only)
        p32 = native if native.dtype == torch.float32 else F.cast(native, torch.float32)
        wc = native if native.dtype == compute_dtype else (p32 if compute_dtype == torch.float32 else F.cast(native, compute_dtype))
        self._flat_cache = (key, wc, p32)
        return wc, p32

    def invalidate(self):
        """Drop the cached flat parameter images.  The cache key is (data_ptr, _version) per parameter, which versioned
        in-place updates (optimiser steps, `p.copy_`, `load_state_dict`) change — but writes through `p.data` (EMA /
        weight-sync code, some checkpoint loaders) do not bump `_version`: call this after such a write."""
        self._flat_cache = None

    def _segments(self, segs):
        """segs: None or (((k_start, len, out_start, tok_row), ...), T) -> ctypes pointer for the C calls (cached)."""
        if segs is None:
            return None
        cache = self.__dict__.setdefault("_segs_cache", {})
        hit = cache.get(segs)
        if hit is None:
            if len(cache) > 64:
                cache.clear()
            st, arr = L.make_segments(segs[0], segs[1])
            hit = cache[segs] = (C.pointer(st), st, arr)
        return hit[0]

    def _pool(self, rows, slots, dtype, dev, code, segs=None):
        pools = getattr(self, "_pools", None)
        if pools is None:
            pools = self._pools = OrderedDict()
        key = (tuple(sorted(rows.items())), dtype, str(dev), segs)
        p = pools.get(key)
        if p is None:
            while len(pools) >= 3:                       # a few hundred MB each: keep the most recent shapes only
                k_old, p_old = next(iter(pools.items()))
                if p_old.busy():
                    break
                pools.pop(k_old)
            if len(pools) >= 3:
                return None
            p = pools[key] = _Pool(self, rows, slots, dtype, dev, code, self._segments(segs))
        else:
            pools.move_to_end(key)
        return p

    def run(self, rows, inputs, segs=None):
        """rows: {row key: int}; inputs: tensors for ``self.inputs`` (2-D, contiguous; fp32 for slots declared f32, the
        program dtype otherwise); segs: the segment table of programs with segment ops, as
        (((k_start, len, out_start, tok_row), ...), tokens_per_segment).  Returns the output buffers (one tensor per
        ``self.buffer``)."""
        outs = _TapeFn.apply(self, dict(rows), segs, len(inputs), *inputs, *self.params)
        return outs if isinstance(outs, tuple) else (outs,)

    def program_dtype(self, inputs):
        for s, t in zip(self.inputs, inputs):
            if not self.slot_f32[s]:
                return t.dtype
        return torch.float32


def _ptr_array(n):
    return (C.c_void_p * n)()


def _pooling_enabled():
    return os.environ.get("MILB200_TAPE_GRAPHS", "1") != "0"


class _Pool:
    """Pointer-stable buffers of one (tape, shapes, dtype) so that the native side can replay the program as a CUDA
    graph (csrc/tape.cu keys its graph cache on every pointer of a call): the activation arena, the output buffers,
    staging copies of inputs whose address changes from call to call, and the backward's gradient buffers.  Results
    are handed to autograd as fresh copies, so nothing the caller can hold aliases the pool."""

    def __init__(self, tape, rows, slots, dtype, dev, code, segp=None):
        c = tape._freeze()
        lib = L.lib()
        self.arena = torch.empty((lib.milb200_tape_arena_bytes(c["ops"], c["n_ops"], slots, c["n_slots"], code, segp),),
                                 dtype=torch.uint8, device=dev)
        self.out = [torch.empty((int(fn(rows)), cols), dtype=tape.buffer_dtype(i, dtype), device=dev)
                    for i, (fn, cols) in enumerate(tape.buffers)]
        self.in_stage = [None] * len(tape.inputs)
        self.last_ptr = [0] * len(tape.inputs)
        self.seed = [torch.empty_like(b) for b in self.out]
        self.gin = [None] * len(tape.inputs)
        self.g32 = torch.empty((c["total"],), dtype=torch.float32, device=dev)
        self.owner = None            # weakref to the autograd ctx whose backward still needs the arena

    def busy(self):
        return self.owner is not None and self.owner() is not None

    def stage_input(self, j, t):
        """Inputs whose address repeats (cached PE tables, static benchmark tensors) are used in place; inputs that
        arrive at a new address every call (data-loader batches) go through a fixed staging buffer."""
        ptr = t.data_ptr()
        if ptr == self.last_ptr[j] and self.in_stage[j] is None:
            return t
        if self.last_ptr[j] == 0:
            self.last_ptr[j] = ptr
            return t
        if self.in_stage[j] is None:
            self.in_stage[j] = torch.empty_like(t)
        self.in_stage[j].copy_(t)
        return self.in_stage[j]


class _TapeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, tape, rows, segs, n_in, *tensors):
        inputs = [t.contiguous() for t in tensors[:n_in]]
        c = tape._freeze()
        dtype = tape.program_dtype(inputs)
        dev = inputs[0].device
        for s, t in zip(tape.inputs, inputs):
            want = torch.float32 if tape.slot_f32[s] else dtype
            if t.dtype != want:
                raise L.MilB200Error(f"tape: input for slot {s} has dtype {t.dtype}, the program expects {want}")
        code = L.BF16 if dtype == torch.bfloat16 else L.F32
        if dtype not in (torch.float32, torch.bfloat16):
            raise L.MilB200Error(f"unsupported dtype {dtype}: mil_b200 kernels take float32 or bfloat16")
        slots = tape._slots(rows)
        for s, t in zip(tape.inputs, inputs):
            if tuple(t.shape) != (slots[s].rows, slots[s].cols):
                raise L.MilB200Error(f"tape: input for slot {s} has shape {tuple(t.shape)}, expected "
                                     f"{(slots[s].rows, slots[s].cols)}")
        segp = tape._segments(segs)
        wc, p32 = tape._flat(dtype)
        lib = L.lib()
        pool = tape._pool(rows, slots, dtype, dev, code, segs) if _pooling_enabled() else None
        if pool is not None and pool.busy():
            pool = None                      # a forward of the same shape is still waiting for its backward
        if pool is not None:
            inputs = [pool.stage_input(j, t) for j, t in enumerate(inputs)]
            bufs, arena = pool.out, pool.arena
        else:
            bufs = [torch.empty((int(fn(rows)), cols), dtype=tape.buffer_dtype(i, dtype), device=dev)
                    for i, (fn, cols) in enumerate(tape.buffers)]
            arena = torch.empty((lib.milb200_tape_arena_bytes(c["ops"], c["n_ops"], slots, c["n_slots"], code, segp),),
                                dtype=torch.uint8, device=dev)
        ext = _ptr_array(c["n_slots"])
        for s, t in zip(tape.inputs, inputs):
            ext[s] = t.data_ptr()
        for s, b, fn in tape.outputs:
            ext[s] = bufs[b].data_ptr() + int(fn(rows)) * tape.slot_cols[s] * bufs[b].element_size()
        ws = L.workspace(lib.milb200_tape_workspace_bytes(c["ops"], c["n_ops"], slots, c["n_slots"], code, 0, segp), dev)
        L.check(lib.milb200_tape_forward(c["ops"], c["n_ops"], slots, c["n_slots"], c["params"], c["n_params"], ext, L.ptr(wc),
                                         L.ptr(p32), L.ptr(arena), arena.numel(), L.ptr(ws), ws.numel(), code, segp,
                                         L.stream_ptr()), "tape_forward")
        ctx.tape, ctx.rows, ctx.n_in, ctx.code, ctx.segs = tape, rows, n_in, code, segs
        ctx.pool = pool
        ctx.save_for_backward(arena, wc, p32, *inputs, *bufs)
        if pool is not None:
            if any(ctx.needs_input_grad):
                pool.owner = weakref.ref(ctx)
            outs = [b.clone() for b in bufs]     # the caller gets its own copy; the pool keeps the values backward needs
        else:
            outs = bufs
        return tuple(outs) if len(outs) != 1 else outs[0]

    @staticmethod
    def backward(ctx, *gouts):
        tape, rows, n_in, code, pool = ctx.tape, ctx.rows, ctx.n_in, ctx.code, ctx.pool
        saved = ctx.saved_tensors
        arena, wc, p32 = saved[0], saved[1], saved[2]
        inputs = saved[3:3 + n_in]
        bufs = saved[3 + n_in:]
        c = tape._freeze()
        slots = tape._slots(rows)
        segp = tape._segments(ctx.segs)
        dev = bufs[0].device
        gouts = [(g.contiguous() if g is not None else torch.zeros_like(b)) for g, b in zip(gouts, bufs)]
        gouts = [g if g.dtype == b.dtype else F.cast(g, b.dtype) for g, b in zip(gouts, bufs)]
        # the upstream gradients are copied once into buffers this call owns; every output slot's gradient buffer is its
        # row range of that copy, seeded in place (seed pointer == gradient pointer: the C side skips its own copy)
        if pool is not None:
            for st, g in zip(pool.seed, gouts):
                st.copy_(g)
            gouts = pool.seed
        else:
            gouts = [g.clone() for g in gouts]
        n_slots = c["n_slots"]
        ext, gext, seeds = _ptr_array(n_slots), _ptr_array(n_slots), _ptr_array(n_slots)
        gin = []
        FIRST = 4                                  # tape, rows, segs, n_in precede the tensors in apply()
        for j, (s, t) in enumerate(zip(tape.inputs, inputs)):
            ext[s] = t.data_ptr()
            if ctx.needs_input_grad[FIRST + j]:
                if pool is not None:
                    if pool.gin[j] is None:
                        pool.gin[j] = torch.empty_like(t)
                    g = pool.gin[j]
                else:
                    g = torch.empty_like(t)
                gext[s] = g.data_ptr()
                gin.append(g)
            else:
                gin.append(None)
        for i, (s, b, fn) in enumerate(tape.outputs):
            off = int(fn(rows)) * tape.slot_cols[s] * bufs[b].element_size()
            ext[s] = bufs[b].data_ptr() + off
            seeds[s] = gext[s] = gouts[b].data_ptr() + off
        g32 = pool.g32 if pool is not None else torch.empty((c["total"],), dtype=torch.float32, device=dev)
        lib = L.lib()
        ws = L.workspace(lib.milb200_tape_workspace_bytes(c["ops"], c["n_ops"], slots, n_slots, code, 1, segp), dev)
        L.check(lib.milb200_tape_backward(c["ops"], c["n_ops"], slots, n_slots, c["params"], c["n_params"], ext, gext, seeds,
                                          L.ptr(wc), L.ptr(p32), L.ptr(g32), L.ptr(arena), arena.numel(), L.ptr(ws),
                                          ws.numel(), code, segp, L.stream_ptr()), "tape_backward")
        pdt = tape.params[0].dtype
        if pdt != torch.float32:
            gflat = F.cast(g32, pdt)
        else:
            gflat = g32.clone() if pool is not None else g32
        if pool is not None:
            gin = [g.clone() if g is not None else None for g in gin]
            pool.owner = None
        need = ctx.needs_input_grad
        base = FIRST + n_in
        pieces = gflat.split(c["sizes"])          # one call for the ~90 views; 2-D weights get their shape below
        gparams = [(pc if sh is None else pc.view(sh)) if need[base + j] else None
                   for j, (pc, sh) in enumerate(zip(pieces, c["shapes"]))]
        return (None, None, None, None, *gin, *gparams)
