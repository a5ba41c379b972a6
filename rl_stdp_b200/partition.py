"""Neuron partitioning of ONE network over several GPUs (BASELINE config 5).

Rank r integrates neurons [lo_r, hi_r) and stores every synapse whose POST is in that range, so
propagation, LTD, LTP and learn! are local; per simulated millisecond the ranks all-gather the
spike bitmap (N/8 bytes in total: 125 KB at N = 1M) and each rank replays detection for the
neurons it does not own (replicated pre-trace ring + spike list).  `da` and the noise generator
(Philox keyed by the global neuron id) are replicated, so the union of the ranks reproduces the
single-GPU run — bit for bit in exact mode.

In-library exchange (the production path, `connect_local` / `connect_dist`): the ranks swap IPC
handles of their bitmap rings once; from then on `neuron_kernel` stores its spike words into the
peers' rings over NVLink and raises a step counter there, the receiver waits on the device — no
host code per step, the window runs as one CUDA graph.

Exchange back-ends for `snn_set_partition`'s callback (host-driven, kept for comparison):
  TorchDistExchange  one process per GPU, `torch.distributed.all_gather_into_tensor` (NCCL over
                     NVLink) on the library's stream; also usable with CPU tensors over gloo.
  LocalExchange      several handles inside one process (threads), device-to-device copies — used
                     to test the partitioned path on a single GPU.
"""
from __future__ import annotations

import ctypes as C
import threading

import numpy as np

from . import _lib as L
from .network import SparseNetwork


def plan(n_neurons: int, world: int, align: int = 128):
    """Contiguous, `align`-aligned neuron ranges, as equal as possible (all equal when
    n_neurons % (align*world) == 0, which all_gather_into_tensor needs)."""
    blocks = (n_neurons + align - 1) // align
    out, start = [], 0
    for r in range(world):
        nb = blocks // world + (1 if r < blocks % world else 0)
        lo, hi = start * align, min(n_neurons, (start + nb) * align)
        out.append((lo, hi))
        start += nb
    return out


def split_by_post(rowptr, post, w, delay, lo: int, hi: int):
    """CSR (all presynaptic rows) restricted to synapses whose post id is in [lo, hi).
    Returns (rowptr, post, w, delay, index) with `index` = positions in the caller's arrays."""
    rowptr = np.asarray(rowptr, np.int64)
    post = np.asarray(post)
    keep = (post >= lo) & (post < hi)
    idx = np.flatnonzero(keep)
    pre = np.repeat(np.arange(len(rowptr) - 1), np.diff(rowptr))
    cnt = np.bincount(pre[idx], minlength=len(rowptr) - 1)
    rp = np.zeros(len(rowptr), np.int64)
    np.cumsum(cnt, out=rp[1:])
    return rp, post[idx].astype(np.int32), np.asarray(w, np.float64)[idx], (
        None if delay is None else np.asarray(delay, np.uint8)[idx]), idx


class _PtrTensor:
    """Minimal __cuda_array_interface__ holder so torch can alias library memory (no copy)."""

    def __init__(self, ptr: int, n_words: int):
        self.__cuda_array_interface__ = {"shape": (n_words,), "typestr": "<i4", "data": (int(ptr), False), "version": 3}


class TorchDistExchange:
    """all-gather of the per-step bitmap words with torch.distributed (NCCL on GPUs, gloo on CPU)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)

    def gather_words(self, local):
        """local: 1-D int32 tensor with this rank's words -> 1-D tensor with every rank's words."""
        import torch
        out = torch.empty(local.numel() * self.world, dtype=local.dtype, device=local.device)
        self.dist.all_gather_into_tensor(out, local.contiguous(), group=self.group)
        return out

    def callback(self):
        import torch
        state = {}

        def cb(user, words, word_lo, n_words, total_words, stream):
            try:
                full = torch.as_tensor(_PtrTensor(words, total_words), device="cuda")
                ext = torch.cuda.ExternalStream(stream)
                with torch.cuda.stream(ext):
                    if n_words * self.world == total_words:
                        # equal shards: in-place all-gather, rank r's words already sit at their place
                        self.dist.all_gather_into_tensor(full, full[word_lo:word_lo + n_words], group=self.group)
                    else:
                        if "shards" not in state:  # once: everybody's (word_lo, n_words)
                            mine = torch.tensor([word_lo, n_words], dtype=torch.int64, device="cuda")
                            allv = torch.empty(2 * self.world, dtype=torch.int64, device="cuda")
                            self.dist.all_gather_into_tensor(allv, mine, group=self.group)
                            state["shards"] = allv.view(self.world, 2).tolist()
                            mx = max(n for _, n in state["shards"])
                            state["pad"] = torch.zeros(mx, dtype=torch.int32, device="cuda")
                            state["tmp"] = torch.empty(mx * self.world, dtype=torch.int32, device="cuda")
                        pad, tmp = state["pad"], state["tmp"]
                        pad[:n_words].copy_(full[word_lo:word_lo + n_words])
                        self.dist.all_gather_into_tensor(tmp, pad, group=self.group)
                        mx = pad.numel()
                        for r, (lo, n) in enumerate(state["shards"]):
                            if r != self.rank:
                                full[lo:lo + n].copy_(tmp[r * mx:r * mx + n])
                return 0
            except Exception:  # never let an exception cross the C boundary
                import traceback
                traceback.print_exc()
                return 1

        self._cb = L.EXCHANGE_FN(cb)
        return self._cb


class LocalExchange:
    """Bitmap exchange between `world` handles living in one process on one GPU (one thread each)."""

    def __init__(self, world: int):
        self.world = world
        self.barrier = threading.Barrier(world)
        self.slots = [None] * world
        self.cudart = C.CDLL("libcudart.so")
        self.cudart.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
        self.cudart.cudaStreamSynchronize.argtypes = [C.c_void_p]
        self._cbs = []

    def callback(self, rank: int):
        def cb(user, words, word_lo, n_words, total_words, stream):
            try:
                self.cudart.cudaStreamSynchronize(stream)
                self.slots[rank] = (words, word_lo, n_words)
                self.barrier.wait()
                for r in range(self.world):
                    if r == rank:
                        continue
                    src, lo, n = self.slots[r]
                    rc = self.cudart.cudaMemcpy(words + 4 * lo, src + 4 * lo, 4 * n, 3)  # device to device
                    if rc != 0:
                        return 100 + rc
                self.cudart.cudaDeviceSynchronize()
                self.barrier.wait()  # nobody overwrites its bitmap before every peer has copied it
                return 0
            except Exception:
                import traceback
                traceback.print_exc()
                self.barrier.abort()
                return 1

        f = L.EXCHANGE_FN(cb)
        self._cbs.append(f)
        return f


def export_blob(net) -> bytes:
    buf = C.create_string_buffer(L.SNN_PEER_BLOB_BYTES)
    L.check(L.load().snn_partition_export(net._h, buf))
    return buf.raw


def connect(net, rank: int, blobs):
    """blobs: one export_blob() per rank, in rank order."""
    allb = b"".join(blobs)
    L.check(L.load().snn_partition_connect(net._h, int(rank), len(blobs), allb))


def connect_local(nets):
    """Handles of ONE process (same or different GPUs): plain device pointers, no IPC."""
    blobs = [export_blob(n) for n in nets]
    for r, n in enumerate(nets):
        connect(n, r, blobs)


def connect_dist(net, group=None, _export=None, _connect=None):
    """One process per GPU: all-gather the CUDA IPC blobs with torch.distributed, connect, barrier.
    (`_export` / `_connect` replace the library calls in the CPU test of this host logic.)"""
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    blobs = [None] * world
    dist.all_gather_object(blobs, (_export or export_blob)(net), group=group)
    (_connect or connect)(net, rank, blobs)
    dist.barrier(group)  # nobody steps before every rank has mapped its peers


class PartitionedRank(SparseNetwork):
    """The rank-local handle: same API as SparseNetwork, synapses restricted to local posts."""

    def __init__(self, n_neurons, n_exc, rowptr, post, w, delay, lo, hi, exchange_cb=None, **kw):
        rp, po, ww, dd, idx = split_by_post(rowptr, post, w, delay, lo, hi)
        self.lo, self.hi, self.local_index = int(lo), int(hi), idx
        if dd is not None and "max_delay" not in kw:
            kw["max_delay"] = int(np.asarray(delay).max()) if len(delay) else 1
        super().__init__(n_neurons, n_exc, rp, po, ww, dd, **kw)
        self._exchange_cb = exchange_cb  # keep the ctypes callback alive
        L.check(L.load().snn_set_partition(self._h, self.lo, self.hi, exchange_cb, None))


def run_local_ranks(nets, n_steps, noise, train, **kw):
    """Drives several PartitionedRank handles of ONE process in lock-step (one thread each)."""
    out = [None] * len(nets)
    errs = []

    def work(r):
        try:
            out[r] = nets[r].run(n_steps, noise, train, **kw)
        except Exception as e:  # pragma: no cover
            errs.append(e)

    ths = [threading.Thread(target=work, args=(r,)) for r in range(len(nets))]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    if errs:
        raise errs[0]
    return out
