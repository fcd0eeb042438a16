"""Upstream feeder for the packed-CSR kernels (reference: dataset.py:366-393).

The reference's Dataset returns ONE slide per sample: `np.load(<pathology_path>/<hospital>/<Biopsy|Resection>/
<patient>.npy)` -> [n, 768] float, optionally thinned by the augmentation of dataset.py:375-381 (keep a sorted random
90 % of a biopsy's rows, 80 % of a resection's), zero-padded to 15 592 rows when batch_size > 1 (dataset.py:383-390)
and converted with `.float()`.  The DataLoader then collates, and `.cuda()` ships fp32.

`PackedBagFeeder` produces what the B200 kernels want instead: per step, the slides of the batch are gathered by native
threads (`milb200_pack_bags_host`) straight into a pinned staging buffer — rows back to back, already bf16 (RNE), with
the int32 CSR offsets — and copied to the device on a copy stream while the previous step computes.  Nothing is padded
and half the bytes cross PCIe.  Two staging/device buffer pairs rotate; a worker thread packs batch k+1 while batch k
is in flight.  The numpy arrays may be memory-mapped (`np.load(..., mmap_mode="r")`), in which case the page-cache reads
happen inside the packing threads.
"""
from __future__ import annotations

import ctypes as C
import queue
import random
import threading

import numpy as np
import torch

from . import _lib as L

_HOST_CODES = {np.dtype(np.float32): L.HOST_F32, np.dtype(np.float64): L.HOST_F64, np.dtype(np.float16): L.HOST_F16}

# dataset.py:375-381: fraction of rows KEPT in train mode with augmentation, by slide type
AUGMENT_KEEP = {"Biopsy": 0.9, "Resection": 0.8}


def augmentation_rows(n, slide_type, rng):
    """The reference's row thinning: `sorted(random.sample(range(n), int(n * keep)))` (dataset.py:377-381)."""
    keep = AUGMENT_KEEP.get(slide_type)
    if keep is None:
        return None
    return np.asarray(sorted(rng.sample(range(n), int(n * keep))), dtype=np.int32)


def pack_bags_host(bags, dst, offsets, keep_rows=None, n_threads=0):
    """Gather `bags` (list of 2-D numpy arrays [n_b, L], C-contiguous rows; or bf16 torch CPU tensors) into `dst`
    (CPU torch tensor [capacity, L], fp32 or bf16, normally pinned) and fill `offsets` (CPU int32 [len(bags)+1]).
    Returns the number of packed rows.  keep_rows: optional list (None or sorted int32 row indices per bag)."""
    nb = len(bags)
    if nb == 0:
        raise ValueError("pack_bags_host: empty batch")
    if dst.device.type != "cpu" or offsets.device.type != "cpu" or offsets.dtype != torch.int32 \
            or offsets.numel() < nb + 1 or not dst.is_contiguous() or dst.dim() != 2:
        raise ValueError("pack_bags_host: dst must be a contiguous CPU [rows, L] tensor and offsets CPU int32 [B+1]")
    Lf = dst.shape[1]
    ptrs = (C.c_void_p * nb)()
    rows = (C.c_int64 * nb)()
    pitch = (C.c_int64 * nb)()
    code = None
    for b, a in enumerate(bags):
        if isinstance(a, torch.Tensor):
            if a.dtype != torch.bfloat16 or a.device.type != "cpu" or a.dim() != 2 or a.stride(1) != 1:
                raise ValueError("pack_bags_host: torch sources must be CPU bf16 [n, L] with contiguous rows")
            c, ptr, n, lf, pb = L.HOST_BF16, a.data_ptr(), a.shape[0], a.shape[1], a.stride(0) * 2
        else:
            if a.ndim != 2 or a.dtype not in _HOST_CODES or (a.size and a.strides[1] != a.itemsize):
                raise ValueError("pack_bags_host: numpy sources must be [n, L] float16/32/64 with contiguous rows")
            c, ptr, n, lf, pb = _HOST_CODES[a.dtype], a.ctypes.data, a.shape[0], a.shape[1], a.strides[0]
        if lf != Lf:
            raise ValueError(f"pack_bags_host: bag {b} has {lf} features, destination {Lf}")
        if code is None:
            code = c
        elif code != c:
            raise ValueError("pack_bags_host: all bags of a batch must share one source dtype")
        ptrs[b], rows[b], pitch[b] = ptr, n, pb
    kp = kc = None
    keep_alive = []
    if keep_rows is not None and any(k is not None for k in keep_rows):
        kp = (C.c_void_p * nb)()
        kc = (C.c_int64 * nb)()
        for b, k in enumerate(keep_rows):
            if k is None:
                kp[b], kc[b] = None, rows[b]
            else:
                k = np.ascontiguousarray(k, dtype=np.int32)
                keep_alive.append(k)
                kp[b], kc[b] = k.ctypes.data, k.shape[0]
    L.check(L.lib().milb200_pack_bags_host(ptrs, rows, pitch, kp, kc, nb, Lf, code, dst.data_ptr(),
                                           L.dtype_code(dst), dst.shape[0], offsets.data_ptr(), int(n_threads)),
            "pack_bags_host")
    return int(offsets[nb])


class PackedBagFeeder:
    """Iterates packed-CSR batches on the device.

    sources: sequence of per-slide feature matrices (numpy arrays, possibly memory-mapped, or paths to `.npy` files).
    slide_types: optional sequence of "Biopsy"/"Resection" (dataset.py:376-381) enabling the train-mode augmentation.
    Each item is (X [total_n, L] on `device`, offsets int32 [B+1] on `device`, indices of the bags in the batch);
    X and offsets are views of a rotating device buffer: consume them before asking for the batch after next
    (the iterator waits on the consumer's stream before reusing a buffer)."""

    def __init__(self, sources, batch_bags, L_feat, device="cuda", dtype=torch.bfloat16, slide_types=None, augment=False,
                 shuffle=False, seed=1234, rank=0, world=1, n_threads=0, capacity_rows=None, drop_last=False):
        self.sources, self.B, self.L = list(sources), int(batch_bags), int(L_feat)
        self.device, self.dtype = torch.device(device), dtype
        self.slide_types, self.augment = slide_types, bool(augment)
        self.shuffle, self.seed, self.rank, self.world = shuffle, seed, rank, world
        self.n_threads, self.drop_last = n_threads, drop_last
        self.epoch = 0
        self._rows = [self._open(i).shape[0] for i in range(len(self.sources))]
        if capacity_rows is None:      # worst batch: the B largest bags
            capacity_rows = sum(sorted(self._rows, reverse=True)[:self.B])
        self.capacity = int(capacity_rows)
        pin = self.device.type == "cuda"
        self._stage = [torch.empty((self.capacity, self.L), dtype=dtype, pin_memory=pin) for _ in range(2)]
        self._off_h = [torch.empty(self.B + 1, dtype=torch.int32, pin_memory=pin) for _ in range(2)]
        self._dev = [torch.empty((self.capacity, self.L), dtype=dtype, device=self.device) for _ in range(2)]
        self._off_d = [torch.empty(self.B + 1, dtype=torch.int32, device=self.device) for _ in range(2)]
        if pin:
            self._copy_stream = torch.cuda.Stream(device=self.device)
            self._copied = [torch.cuda.Event() for _ in range(2)]
            self._consumed = [torch.cuda.Event() for _ in range(2)]
            self._staged_free = [torch.cuda.Event() for _ in range(2)]

    def _open(self, i):
        s = self.sources[i]
        if isinstance(s, (str, bytes)) or hasattr(s, "__fspath__"):
            return np.load(s, mmap_mode="r")
        return s

    def set_epoch(self, epoch):          # DistributedSampler's contract (train_ddp.py:201-202)
        self.epoch = int(epoch)

    def order(self):
        """This rank's bag indices for the epoch: DistributedSampler's permutation, padding and stride
        (train_ddp.py:191, drop_last=False): the index list is extended with its own head to a multiple of `world`, so
        every rank draws ceil(n / world) bags (and therefore the same number of batches — each step ends in one
        all-reduce, a rank that ran out of batches would leave the others waiting in it)."""
        n = len(self.sources)
        if self.shuffle:
            g = torch.Generator().manual_seed(self.seed + self.epoch)
            idx = torch.randperm(n, generator=g).tolist()
        else:
            idx = list(range(n))
        if self.world > 1 and n > 0:
            total = (n + self.world - 1) // self.world * self.world
            pad = total - n
            if pad:
                idx = idx + (idx * ((pad + n - 1) // n))[:pad]
        return idx[self.rank::self.world]

    def batches(self):
        idx = self.order()
        out = [idx[i:i + self.B] for i in range(0, len(idx), self.B)]
        if self.drop_last and out and len(out[-1]) < self.B:
            out.pop()
        return out

    def __len__(self):
        return len(self.batches())

    def _pack(self, slot, bag_ids, rng):
        arrays = [self._open(i) for i in bag_ids]
        keep = None
        if self.augment and self.slide_types is not None:
            keep = [augmentation_rows(a.shape[0], self.slide_types[i], rng) for a, i in zip(arrays, bag_ids)]
        n = pack_bags_host(arrays, self._stage[slot], self._off_h[slot], keep, self.n_threads)
        return n

    def __iter__(self):
        batches = self.batches()
        rng = random.Random(self.seed + 7919 * self.epoch + self.rank)
        cuda = self.device.type == "cuda"
        q: "queue.Queue" = queue.Queue(maxsize=1)
        free = [threading.Semaphore(1), threading.Semaphore(1)]      # staging slot may be overwritten

        stop = threading.Event()            # set when the consumer abandons the iteration: the worker must not block forever

        def put(item):
            while not stop.is_set():
                try:
                    q.put(item, timeout=0.2)
                    return True
                except queue.Full:
                    continue
            return False

        def worker():
            try:
                for k, ids in enumerate(batches):
                    slot = k & 1
                    while not free[slot].acquire(timeout=0.2):
                        if stop.is_set():
                            return
                    if cuda:
                        self._staged_free[slot].synchronize()      # the H2D copy that last read this staging slot is done
                    n = self._pack(slot, ids, rng)
                    if not put((slot, n, ids)):
                        return
                put(None)
            except BaseException as e:       # surface packing errors in the consumer
                put(e)

        th = threading.Thread(target=worker, daemon=True)
        th.start()
        if cuda:
            main = torch.cuda.current_stream(self.device)
            for ev in self._consumed:
                ev.record(main)
        try:
            yield from self._consume(q, free, cuda)
        finally:
            stop.set()
            th.join()

    def _consume(self, q, free, cuda):
        while True:
            item = q.get()
            if item is None:
                break
            if isinstance(item, BaseException):
                raise item
            slot, n, ids = item
            nb = len(ids)
            if cuda:
                main = torch.cuda.current_stream(self.device)
                with torch.cuda.stream(self._copy_stream):
                    self._copy_stream.wait_event(self._consumed[slot])     # kernels that read this device buffer are done
                    self._dev[slot][:n].copy_(self._stage[slot][:n], non_blocking=True)
                    self._off_d[slot][:nb + 1].copy_(self._off_h[slot][:nb + 1], non_blocking=True)
                    self._copied[slot].record(self._copy_stream)
                    self._staged_free[slot].record(self._copy_stream)
                free[slot].release()
                main.wait_event(self._copied[slot])
                yield self._dev[slot][:n], self._off_d[slot][:nb + 1], ids
                self._consumed[slot].record(torch.cuda.current_stream(self.device))
            else:
                self._dev[slot][:n].copy_(self._stage[slot][:n])
                self._off_d[slot][:nb + 1].copy_(self._off_h[slot][:nb + 1])
                free[slot].release()
                yield self._dev[slot][:n], self._off_d[slot][:nb + 1], ids
