"""Variant E — the layered Izhikevich `Network(arch, param)` of
Julia/Modules/Networks/net_arch.jl:72-129,135-137,239-281,316-379 — on the sparse CUDA engine.

The layered net is a special case of the CSR engine: dense exc->exc blocks e[l] (plastic), the
hidden layers' excitatory->partner synapses ei[l] = wei*I and lateral inhibition
ie[l] = wie*(1-I) (both fixed), homeostasis on excitatory neurons beyond layer 1, dopamine-gated
learning `e[l] += da*de[l]` with `de` never decayed.  Neuron ids follow `Ind` (:40-68): all
excitatory layers first, then one inhibitory partner per hidden-layer neuron.  Ids are 0-based.
"""
from __future__ import annotations

import numpy as np

from . import _lib as L
from .network import SparseNetwork


def layered_csr(arch, e0, wei, wie):
    """CSR (ascending post per row) of a layered net; returns (rowptr, post, w, e_index) where
    e_index[l] is an int array [arch[l], arch[l+1]] of CSR positions of e[l][i, j]."""
    arch = [int(a) for a in arch]
    Lh = len(arch)
    off = np.concatenate([[0], np.cumsum(arch)]).astype(int)
    n_exc = int(off[-1])
    ioff = {}
    acc = n_exc
    for l in range(1, Lh - 1):
        ioff[l] = acc
        acc += arch[l]
    n = acc
    rows = [[] for _ in range(n)]
    for l in range(Lh - 1):
        for i in range(arch[l]):
            pre = off[l] + i
            for j in range(arch[l + 1]):
                rows[pre].append((off[l + 1] + j, float(e0[l][i, j]), (l, i, j)))
    for l in ioff:
        for i in range(arch[l]):
            rows[off[l] + i].append((ioff[l] + i, float(wei), None))            # ei[l] = wei .* I   :30
            for j in range(arch[l]):
                if j != i:
                    rows[ioff[l] + i].append((off[l] + j, float(wie), None))     # ie[l] = wie .* (1-I) :29
    rowptr = np.zeros(n + 1, np.int64)
    post, w = [], []
    e_index = [np.zeros((arch[l], arch[l + 1]), np.int64) for l in range(Lh - 1)]
    for pre in range(n):
        for (po, ww, tag) in sorted(rows[pre], key=lambda x: x[0]):
            if tag is not None:
                e_index[tag[0]][tag[1], tag[2]] = len(post)
            post.append(po); w.append(ww)
        rowptr[pre + 1] = len(post)
    return rowptr, np.array(post, np.int32), np.array(w, np.float64), e_index, n_exc, n


class LayeredNetwork:
    """`Network(arch, param)` (Izhikevich).  `e0`: list of initial e[l] — the recorded
    `clamp.(randn(arch[l],arch[l+1]) .* std .+ m, 0, wmax)` of Weights (:21)."""

    def __init__(self, arch, param, e0, dtype="f64", mode="exact", device=0):
        self.arch = [int(a) for a in arch]
        self.param = dict(param)
        p = self.param
        rowptr, post, w, self._e_index, self.n_exc, self.n_neurons = layered_csr(arch, e0, p["wei"], p["wie"])
        self._sparse = SparseNetwork(
            self.n_neurons, self.n_exc, rowptr, post, w, None, dtype=dtype, mode=mode, noise="none", device=device,
            max_delay=1, vthresh=p["vthresh"], v_reset=p["c"], a_exc=p["ae"], a_inh=p["ai"], d_exc=p["de"], d_inh=p["di"],
            trace_set=0.1, trace_decay=p["ttrace"], apost=p["apost"], apre=p["apre"], da_decay=p["tda"],
            rate_mode=L.SNN_RATE_DA, sd_decay_mode=L.SNN_SD_DECAY_NEVER, wmin=0.0, wmax=p["wmax"],
            theta=p["theta"], ttheta=p["ttheta"], ho_from=self.arch[0], plastic_ei=0)

    # step!(net, input_spikes, train) :239-281 — Int ids force a spike (:135-137), (id, intensity)
    # tuples inject a current (:316-326)
    def step(self, input_spikes=None, train=False):
        force = currents = None
        if input_spikes is not None and len(input_spikes):
            if isinstance(input_spikes[0], tuple):
                p = self.param
                currents = [(0, int(i), (p["imax"] - p["imin"]) * float(x) + p["imin"]) for i, x in input_spikes]
            else:
                force = [(0, int(i)) for i in input_spikes]
        spikes, _ = self._sparse.run(1, None, [1 if train else 0], force=force, currents=currents)
        return spikes[0]

    def e(self, l):
        return self._sparse.get(L.FIELD_W)[self._e_index[l]]

    def de(self, l):
        return self._sparse.get(L.FIELD_SD)[self._e_index[l]]

    @property
    def da(self):
        return float(self._sparse.get(L.FIELD_DA)[0])

    @da.setter
    def da(self, x):
        self._sparse.set(L.FIELD_DA, [x])

    @property
    def v(self): return self._sparse.get(L.FIELD_V)
    @property
    def u(self): return self._sparse.get(L.FIELD_U)
    @property
    def ho(self): return self._sparse.get(L.FIELD_HO)
    @property
    def trace_e(self): return self._sparse.get(L.FIELD_TRACE)[:self.n_exc]
