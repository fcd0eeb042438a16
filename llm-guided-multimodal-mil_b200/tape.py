"""Static operator programs ("tapes"): host side of csrc/tape.cu.

A module describes its forward ONCE as ops over numbered tensor slots (``Tape.linear / attention / layernorm``); a call
then costs one ``milb200_tape_forward`` and, in backward, one ``milb200_tape_backward`` instead of one autograd node,
several allocations and a ctypes call per kernel.  Row counts are symbolic (``"T"``, ``"N"`` ...) and bound per call,
so one tape serves every bag size.  The whole program is a single ``torch.autograd.Function``; parameter gradients come
back as views of one flat fp32 buffer.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L
from . import functional as F

_ACT = {None: L.ACT_NONE, "none": L.ACT_NONE, "tanh": L.ACT_TANH, "relu": L.ACT_RELU, "sigmoid": L.ACT_SIGMOID}


class Tape:
    def __init__(self):
        self.ops = []            # (kind, in0, in1, in2, out, p0, p1, a0)
        self.slot_rows = []      # symbolic row key per slot
        self.slot_cols = []
        self.slot_ext = []       # 0 internal, 1 external input, 2 external output (lives in an output buffer)
        self.params = []         # nn.Parameter objects, in flat-buffer order
        self._pidx = {}
        self.inputs = []         # slot ids of external inputs, in call order
        self.outputs = []        # (slot, buffer index, row-offset function(rows) -> int)
        self.buffers = []        # (rows function(rows) -> int, cols)
        self._c = None

    # ---- building ------------------------------------------------------------------------------------------
    def slot(self, rows_key, cols, ext=0):
        self.slot_rows.append(rows_key)
        self.slot_cols.append(int(cols))
        self.slot_ext.append(ext)
        return len(self.slot_cols) - 1

    def input(self, rows_key, cols):
        s = self.slot(rows_key, cols, ext=1)
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

    def linear(self, x, lin, act=None, add=None):
        """out = act((x [+ add]) W^T + b) for an ``nn.Linear`` parameter container."""
        out = self.slot(self.slot_rows[x], lin.weight.shape[0])
        self.ops.append((L.OP_LINEAR, x, -1 if add is None else add, -1, out, self.param(lin.weight), self.param(lin.bias),
                         _ACT[act]))
        return out

    def attention(self, q, k, v, heads):
        out = self.slot(self.slot_rows[q], self.slot_cols[q])
        self.ops.append((L.OP_ATTENTION, q, k, v, out, -1, -1, int(heads)))
        return out

    def layernorm(self, x, ln, residual=None):
        if abs(ln.eps - 1e-5) > 1e-12:
            raise L.MilB200Error("layernorm kernels are built for eps = 1e-5 (nn.LayerNorm default)")
        out = self.slot(self.slot_rows[x], self.slot_cols[x])
        self.ops.append((L.OP_LAYERNORM, x, -1 if residual is None else residual, -1, out, self.param(ln.weight),
                         self.param(ln.bias), 0))
        return out

    def add(self, a, b):
        """out = a + b as a slot of its own, for sums that several ops consume (keys + key_pe feeds two projections
        per block: materialise it once instead of once per consumer)."""
        out = self.slot(self.slot_rows[a], self.slot_cols[a])
        self.ops.append((L.OP_ADD, a, b, -1, out, -1, -1, 0))
        return out

    def buffer(self, rows_fn, cols):
        self.buffers.append((rows_fn, int(cols)))
        return len(self.buffers) - 1

    def output(self, slot, buf, row_offset_fn):
        """Have `slot` written in place at row `row_offset_fn(rows)` of output buffer `buf`."""
        if self.slot_ext[slot] != 0:
            raise L.MilB200Error("tape.output: slot is already external")
        if self.slot_cols[slot] != self.buffers[buf][1]:
            raise L.MilB200Error("tape.output: column count differs from the buffer's")
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
        self._c = dict(ops=ops, n_ops=n_ops, n_slots=n_slots, params=params, n_params=n_params, offsets=offsets, total=off)
        return self._c

    def _slots(self, rows):
        n = len(self.slot_cols)
        arr = (L.TapeSlot * n)()
        for i in range(n):
            arr[i] = L.TapeSlot(int(rows[self.slot_rows[i]]), self.slot_cols[i], 1 if self.slot_ext[i] else 0)
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

    def run(self, rows, inputs):
        """rows: {row key: int}; inputs: tensors for ``self.inputs`` (2-D, contiguous, one dtype).  Returns the output
        buffers (one tensor per ``self.buffer``)."""
        outs = _TapeFn.apply(self, dict(rows), len(inputs), *inputs, *self.params)
        return outs if isinstance(outs, tuple) else (outs,)


def _ptr_array(n):
    return (C.c_void_p * n)()


class _TapeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, tape, rows, n_in, *tensors):
        inputs = [t.contiguous() for t in tensors[:n_in]]
        c = tape._freeze()
        dtype = inputs[0].dtype
        dev = inputs[0].device
        if any(t.dtype != dtype for t in inputs):
            raise L.MilB200Error("tape: all inputs must share one dtype")
        code = L.dtype_code(inputs[0])
        slots = tape._slots(rows)
        for s, t in zip(tape.inputs, inputs):
            if tuple(t.shape) != (slots[s].rows, slots[s].cols):
                raise L.MilB200Error(f"tape: input for slot {s} has shape {tuple(t.shape)}, expected "
                                     f"{(slots[s].rows, slots[s].cols)}")
        wc, p32 = tape._flat(dtype)
        bufs = [torch.empty((int(fn(rows)), cols), dtype=dtype, device=dev) for fn, cols in tape.buffers]
        ext = _ptr_array(c["n_slots"])
        for s, t in zip(tape.inputs, inputs):
            ext[s] = t.data_ptr()
        esz = inputs[0].element_size()
        for s, b, fn in tape.outputs:
            ext[s] = bufs[b].data_ptr() + int(fn(rows)) * tape.slot_cols[s] * esz
        lib = L.lib()
        arena_bytes = lib.milb200_tape_arena_bytes(c["ops"], c["n_ops"], slots, c["n_slots"], code)
        arena = torch.empty((arena_bytes,), dtype=torch.uint8, device=dev)
        ws = L.workspace(lib.milb200_tape_workspace_bytes(c["ops"], c["n_ops"], slots, c["n_slots"], code, 0), dev)
        L.check(lib.milb200_tape_forward(c["ops"], c["n_ops"], slots, c["n_slots"], c["params"], c["n_params"], ext, L.ptr(wc),
                                         L.ptr(p32), L.ptr(arena), arena.numel(), L.ptr(ws), ws.numel(), code,
                                         L.stream_ptr()), "tape_forward")
        ctx.tape, ctx.rows, ctx.n_in, ctx.code = tape, rows, n_in, code
        ctx.save_for_backward(arena, wc, p32, *inputs, *bufs)
        ctx.n_bufs = len(bufs)
        return tuple(bufs) if len(bufs) != 1 else bufs[0]

    @staticmethod
    def backward(ctx, *gouts):
        tape, rows, n_in, code = ctx.tape, ctx.rows, ctx.n_in, ctx.code
        saved = ctx.saved_tensors
        arena, wc, p32 = saved[0], saved[1], saved[2]
        inputs = saved[3:3 + n_in]
        bufs = saved[3 + n_in:]
        c = tape._freeze()
        slots = tape._slots(rows)
        dev, dtype = inputs[0].device, inputs[0].dtype
        esz = inputs[0].element_size()
        gouts = [(g.contiguous() if g is not None else torch.zeros_like(b)) for g, b in zip(gouts, bufs)]
        gouts = [g if g.dtype == dtype else F.cast(g, dtype) for g in gouts]
        n_slots = c["n_slots"]
        ext, gext, seeds = _ptr_array(n_slots), _ptr_array(n_slots), _ptr_array(n_slots)
        gin = []
        for j, (s, t) in enumerate(zip(tape.inputs, inputs)):
            ext[s] = t.data_ptr()
            if ctx.needs_input_grad[3 + j]:
                g = torch.empty_like(t)
                gext[s] = g.data_ptr()
                gin.append(g)
            else:
                gin.append(None)
        scratch_out = []
        for s, b, fn in tape.outputs:
            off = int(fn(rows)) * tape.slot_cols[s] * esz
            ext[s] = bufs[b].data_ptr() + off
            seeds[s] = gouts[b].data_ptr() + off
            # an output slot that also feeds later ops needs a writable gradient buffer of its own
            g = torch.empty((slots[s].rows, slots[s].cols), dtype=dtype, device=dev)
            gext[s] = g.data_ptr()
            scratch_out.append(g)
        g32 = torch.zeros((c["total"],), dtype=torch.float32, device=dev)
        lib = L.lib()
        ws = L.workspace(lib.milb200_tape_workspace_bytes(c["ops"], c["n_ops"], slots, n_slots, code, 1), dev)
        L.check(lib.milb200_tape_backward(c["ops"], c["n_ops"], slots, n_slots, c["params"], c["n_params"], ext, gext, seeds,
                                          L.ptr(wc), L.ptr(p32), L.ptr(g32), L.ptr(arena), arena.numel(), L.ptr(ws),
                                          ws.numel(), code, L.stream_ptr()), "tape_backward")
        pdt = tape.params[0].dtype
        gflat = g32 if pdt == torch.float32 else F.cast(g32, pdt)
        gparams = []
        for j, (p, off) in enumerate(zip(tape.params, c["offsets"])):
            gparams.append(gflat[off:off + p.numel()].view(p.shape) if ctx.needs_input_grad[3 + n_in + j] else None)
        return (None, None, None, *gin, *gparams)
