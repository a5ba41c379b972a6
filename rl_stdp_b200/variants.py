"""Host-side mirrors of the reference's other Izhikevich `Network` variants on the sparse CUDA engine.

Every variant is a configuration of libsnn_b200 (snn_config switches) plus a constructor that turns
the reference's dense matrices into CSR synapses; nothing here computes on the CPU.  Ids are
0-based; every random quantity of the reference (connectivity, thalamic noise) is a recorded input.

  NetworkCol     variant B  Julia/Izhikevic_dastdp/Old_net/new_net_col.jl:4-163 (transposed storage)
  NetworkColSep  variant C  Julia/Izhikevic_dastdp/Old_net/new_net_col_sep.jl:4-214
  NetworkRes     variant D  Julia/Modules/Networks/net_res.jl:6-251
  NetworkG       variant G  Julia/Izhikevic_dastdp/Old_net/DASTDP.jl:24-106 + Julia_DA_STDPv2.jl:71-96
"""
from __future__ import annotations

import numpy as np

from . import _lib as L
from .network import SparseNetwork


def _csr_from_dense(connection, weight_of_pre):
    """connection[pre, post] -> (rowptr, post, w) with ascending post inside a row."""
    conn = np.asarray(connection) != 0
    n = conn.shape[0]
    pre, post = np.nonzero(conn)
    rowptr = np.zeros(n + 1, np.int64)
    np.add.at(rowptr, pre + 1, 1)
    return np.cumsum(rowptr), post.astype(np.int32), weight_of_pre(pre, post).astype(np.float64), pre


class _DenseView:
    def _dense(self, field):
        m = np.zeros((self.n_neurons, self.n_neurons))
        m[self._pre, self._post] = self._net.get(field)
        return m


class NetworkCol:
    """Variant B, `Network(n_neurons, connectivity, n_exc)` of new_net_col.jl:86-103: the network of variant A held
    TRANSPOSED — `connection`, `s`, `sd` are [post, pre] ("read by columns", :13-15).  The caller hands the matrices
    over exactly as that struct holds them (column-major [post, pre]) through snn_set_synapses_dense_col; `s[post, pre]`
    / `sd[post, pre]` read them back in the same orientation.  `connection_pre_post` is what connect() draws (:28),
    before the constructor transposes it (:94)."""

    def __init__(self, n_neurons, n_exc, connection_pre_post, dtype="f64", mode="exact", device=0):
        import ctypes as C
        from .network import default_config
        self.n_neurons, self.n_exc = int(n_neurons), int(n_exc)
        N = self.n_neurons
        conn_t = np.ascontiguousarray(np.asarray(connection_pre_post, np.int64).T)      # [post, pre], :94
        s_t = np.zeros((N, N))
        s_t[:, :n_exc] = 1.0                                                              # :97
        s_t[:, n_exc:] = -1.0                                                             # :98
        s_t *= conn_t                                                                     # :99
        self.connection = conn_t
        cfg = default_config(dtype=L.SNN_F64 if dtype == "f64" else L.SNN_F32,
                             mode=L.SNN_MODE_EXACT if mode == "exact" else L.SNN_MODE_FAST, device=int(device),
                             n_neurons=N, n_per_instance=N, n_exc=self.n_exc, max_delay=1, noise_mode=L.SNN_NOISE_REPLAY)
        # SparseNetwork without its constructor: the handle is created here and fed the dense, transposed matrices
        self._net = SparseNetwork.__new__(SparseNetwork)
        self._net.cfg, self._net.n_neurons, self._net.n_instances = cfg, N, 1
        self._net._h = C.c_void_p()
        lib = L.load()
        L.check(lib.snn_create(C.byref(cfg), C.byref(self._net._h)))
        # column-major [post, pre] == the C-order bytes of the [pre, post] transpose
        s_f = np.asfortranarray(s_t); c_f = np.asfortranarray(conn_t)
        L.check(lib.snn_set_synapses_dense_col(self._net._h, s_f.ctypes.data_as(C.c_void_p), c_f.ctypes.data_as(C.c_void_p)))
        self._net.n_synapses = int(lib.snn_n_synapses(self._net._h))
        # the library's CSR order: ascending pre, then post
        self._pre, self._post = np.nonzero(conn_t.T)

    def step(self, noise, train=False, input_spikes=None):            # step! :148-163
        force = None if input_spikes is None else [(0, int(j)) for j in input_spikes]
        sp, _ = self._net.run(1, np.asarray(noise, np.float64).reshape(1, -1), [1 if train else 0], force=force)
        return sp[0]

    def run(self, noise, train, da_events=None, **kw):
        return self._net.run(len(noise), noise, train, da_events=da_events, **kw)

    def _dense_t(self, field):
        m = np.zeros((self.n_neurons, self.n_neurons))
        m[self._post, self._pre] = self._net.get(field)
        return m

    @property
    def s(self): return self._dense_t(L.FIELD_W)       # [post, pre]
    @property
    def sd(self): return self._dense_t(L.FIELD_SD)     # [post, pre]; only plastic (exc pre) entries are maintained
    @property
    def v(self): return self._net.get(L.FIELD_V)
    @property
    def u(self): return self._net.get(L.FIELD_U)
    @property
    def STDP(self): return self._net.get(L.FIELD_TRACE)
    @property
    def da(self): return float(self._net.get(L.FIELD_DA)[0])
    @da.setter
    def da(self, x): self._net.set(L.FIELD_DA, [x])


class NetworkColSep(_DenseView):
    """`Network(n_neurons, connectivity, n_exc)` of new_net_col_sep.jl:112-134: excitatory weights 1.0
    (plastic among excitatory neurons only, :167-172), fixed inhibitory weight -0.5 (:104), homeostasis
    `ho[spiked] += 2.0` for every spiking neuron beyond arch[1] = 0, inhibitory ones included
    (:149-150), `ho *= 0.9999`.  `connection[pre, post]` replaces `rand(N,N) .< connectivity` (:32)."""

    def __init__(self, n_neurons, n_exc, connection, dtype="f64", mode="exact", device=0):
        self.n_neurons, self.n_exc = int(n_neurons), int(n_exc)
        rowptr, post, w, pre = _csr_from_dense(connection, lambda i, j: np.where(i < n_exc, 1.0, -0.5))
        self._pre, self._post = pre, post
        self._net = SparseNetwork(self.n_neurons, self.n_exc, rowptr, post, w, dtype=dtype, mode=mode, noise="replay",
                                  device=device, plastic_ei=0, theta=2.0, ttheta=0.9999, ho_from=0, ho_inh=1)

    def step(self, noise, train=False, input_spikes=None, input_current=None):
        """step!(net[, input], train) :191-214.  input_current = (indices, intensity, noise_in): the
        current method draws its own thalamic noise (:133), which is passed in as recorded."""
        force = None if input_spikes is None else [(0, int(j)) for j in input_spikes]
        cur = None
        if input_current is not None:
            idx, inten, nz = input_current
            cur = [(0, int(j), float(nz[j] + ((20.0 - 0.0) * x + 0.0))) for j, x in zip(idx, inten)]  # :134-138
        sp, _ = self._net.run(1, np.asarray(noise, np.float64).reshape(1, -1), [1 if train else 0], force=force,
                              currents=cur)
        return sp[0]

    @property
    def s(self): return self._dense(L.FIELD_W)      # s[pre, post] (the reference stores the transpose)
    @property
    def sd(self): return self._dense(L.FIELD_SD)
    @property
    def v(self): return self._net.get(L.FIELD_V)
    @property
    def u(self): return self._net.get(L.FIELD_U)
    @property
    def ho(self): return self._net.get(L.FIELD_HO)
    @property
    def STDP(self): return self._net.get(L.FIELD_TRACE)
    @property
    def da(self): return float(self._net.get(L.FIELD_DA)[0])
    @da.setter
    def da(self, x): self._net.set(L.FIELD_DA, [x])


class NetworkRes(_DenseView):
    """`Network(arch, param)` / `Network(n, connectivity, n_exc, param)` of net_res.jl:78-139: weights
    wee / wei / wie, no thalamic noise (:166), threshold 30 on v-ho (:167), homeostasis on excitatory
    spikers beyond arch[1] (:170-171), STDP on exc->exc always and on exc->inh only for the random
    graph arch == [0] (:190-201), learn! with the hard-coded 0.002 / 0.99 (:211-223).
    `param` = the YAML dict (Modules/cfg.yml keys); `connection[pre, post]` = the recorded connect()."""

    def __init__(self, n_neurons, n_exc, connection, arch, param, dtype="f64", mode="exact", device=0):
        self.n_neurons, self.n_exc, self.arch, self.param = int(n_neurons), int(n_exc), list(arch), dict(param)
        P = self.param

        def wt(i, j):
            return np.where(i < n_exc, np.where(j < n_exc, P["wee"], P["wei"]), P["wie"])

        rowptr, post, w, pre = _csr_from_dense(connection, wt)
        self._pre, self._post = pre, post
        self._net = SparseNetwork(
            self.n_neurons, self.n_exc, rowptr, post, w, dtype=dtype, mode=mode, noise="none", device=device,
            vthresh=30.0, v_reset=P["c"], a_exc=P["ae"], a_inh=P["ai"], d_exc=P["de"], d_inh=P["di"],
            theta=P["theta"], ttheta=P["ttheta"], ho_from=int(self.arch[0]), apost=P["apost"], apre=P["apre"],
            da_decay=P["tda"], wmax=P["wmax"], plastic_ei=1 if self.arch == [0] else 0)

    def step(self, train=False, input_spikes=None, input_current=None):
        """step!(net[, input], train) :227-251; input_current = (indices, intensity) :150-160."""
        P = self.param
        force = None if input_spikes is None else [(0, int(j)) for j in input_spikes]
        cur = None
        if input_current is not None:
            idx, inten = input_current
            cur = [(0, int(j), float((P["imax"] - P["imin"]) * x + P["imin"])) for j, x in zip(idx, inten)]
        sp, _ = self._net.run(1, None, [1 if train else 0], force=force, currents=cur)
        return sp[0]

    @property
    def s(self): return self._dense(L.FIELD_W)
    @property
    def sd(self): return self._dense(L.FIELD_SD)
    @property
    def v(self): return self._net.get(L.FIELD_V)
    @property
    def u(self): return self._net.get(L.FIELD_U)
    @property
    def ho(self): return self._net.get(L.FIELD_HO)
    @property
    def trace(self): return self._net.get(L.FIELD_TRACE)
    @property
    def da(self): return float(self._net.get(L.FIELD_DA)[0])
    @da.setter
    def da(self, x): self._net.set(L.FIELD_DA, [x])


class NetworkG:
    """The original sparse DA-STDP net (DASTDP.jl / Julia_DA_STDPv2.jl): fan-out table post[N, M],
    delay D = 1, threshold `>=` 30, LTP with the pre trace of the previous ms, all LTP before all LTD,
    currents accumulated in descending presynaptic id (snn_config.variant_g).  One documented
    difference: where a row of `post` names the same target twice, the reference's
    `I[ind] = I[ind] .+ s[...]` (DASTDP.jl:75) keeps only the last of the two weights (an artefact of
    indexed assignment with repeated indices); the library delivers every synapse."""

    def __init__(self, post, n_exc, s0=None, dtype="f64", mode="exact", device=0):
        post = np.asarray(post, np.int32)
        self.N, self.M = post.shape
        self.Ne = int(n_exc)
        if s0 is None:
            s0 = np.vstack([np.ones((self.Ne, self.M)), -np.ones((self.N - self.Ne, self.M))])  # Julia_DA_STDPv2.jl:53
        rowptr = np.arange(self.N + 1, dtype=np.int64) * self.M
        self._net = SparseNetwork(self.N, self.Ne, rowptr, post.reshape(-1), np.asarray(s0, np.float64).reshape(-1),
                                  dtype=dtype, mode=mode, noise="replay", device=device, threshold_ge=1, variant_g=1)

    def run(self, noise, msec0=1, da_events=None):
        """len(noise) passes of the ms loop body (Julia_DA_STDPv2.jl:73-86) starting at `msec0`;
        synweight_step fires when msec % 10 == 0."""
        n = len(noise)
        train = np.array([((msec0 + k) % 10) == 0 for k in range(n)], np.uint8)
        sp, _ = self._net.run(n, np.asarray(noise, np.float64), train, da_events=da_events)
        return sp

    @property
    def s(self): return self._net.get(L.FIELD_W).reshape(self.N, self.M)
    @property
    def sd(self): return self._net.get(L.FIELD_SD).reshape(self.N, self.M)
    @property
    def v(self): return self._net.get(L.FIELD_V)
    @property
    def u(self): return self._net.get(L.FIELD_U)
    @property
    def DA(self): return float(self._net.get(L.FIELD_DA)[0])
