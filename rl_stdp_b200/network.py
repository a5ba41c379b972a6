"""Host-side mirror of the reference's L1 `Network` API over libsnn_b200.so.

`SparseNetwork` is the general handle (CSR synapses, axonal delays, batched instances).
`Network(n_neurons, connectivity, n_exc)` mirrors variant A
(Julia/Izhikevic_dastdp/Old_net/new_net.jl:4-161): same constructor arguments, `step(train)`,
`step(input_spikes, train)`, `da`, `s[pre, post]`, `sd[pre, post]`, `v`, `u`, `STDP` — so that
parity tests read like the reference's own driver (newnet_DASTDP.jl).  Ids are 0-based here.

All arithmetic runs in the CUDA library; nothing here computes on the CPU.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib as L


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


@dataclass
class StepStats:
    n_spikes: int
    n_syn_events: int
    n_ltp_events: int
    device_ms: float
    learn_ms: float
    neuron_ms: float
    synapse_ms: float
    learn_launches: int
    neuron_launches: int
    synapse_launches: int
    pass_launches: int = 0  # kernels of the asynchronous learn passes (pass + control + replay)


def default_config(**over) -> L.SnnConfig:
    cfg = L.SnnConfig()
    L.check(L.load().snn_config_default(C.byref(cfg)))
    for k, v in over.items():
        if not hasattr(cfg, k):
            raise AttributeError(f"snn_config has no field {k!r}")
        setattr(cfg, k, v)
    return cfg


class SparseNetwork:
    """One (possibly batched) sparse Izhikevich DA-STDP network living on a B200."""

    def __init__(self, n_neurons, n_exc, rowptr, post, w, delay=None, *, n_instances=1, dtype="f32",
                 mode="fast", noise="replay", device=0, **params):
        lib = L.load()
        n_per = int(n_neurons)
        max_delay = int(params.pop("max_delay", int(delay.max()) if delay is not None and len(delay) else 1))
        # everything that is not the model goes through snn_set_param, by name (ABI 3)
        opts = {}
        if not bool(params.pop("graph", True)):
            opts["graph"] = 0          # plain stream-ordered launches instead of whole-window schedules
        persistent = params.pop("persistent", None)   # None: the library's automatic choice
        if persistent is not None:
            opts["persistent"] = 1 if persistent else 0   # the persistent chain of resident role kernels / the captured graph
        exp_flags = int(params.pop("exp_flags", 0))
        for bit, name in ((2, "timeline"), (4, "serialise_learn_pass"), (8, "neuron_blocks_256"), (512, "lsu_pass")):
            if exp_flags & bit:
                opts[name] = 1
        if exp_flags & 1024:
            opts["timeline"] = 2       # chain phases stamped by one block of the role, not the envelope over all blocks
        opts["learn_async"] = {"auto": 0, "off": 1, "force": 2}[params.pop("learn_async", "auto")]
        opts["learn_log_cap"] = int(params.pop("learn_log_cap", 0))
        # opt-in, production mode only: event-driven learning, one dense re-basing pass
        # every `learn_lazy` learn steps (and at every learn step while da > learn_lazy_da, if given)
        opts["learn_lazy"] = int(params.pop("learn_lazy", 0))
        opts["learn_lazy_da"] = float(params.pop("learn_lazy_da", 0.0))
        opts["learn_pass_tile_k"] = int(params.pop("learn_tile_k", 0))
        opts["learn_pass_stages"] = int(params.pop("learn_stages", 0))
        opts["learn_pass_blocks"] = int(params.pop("learn_blocks", 0))
        side_blocks = int(params.pop("side_blocks", 0))  # 10*synapse + ltp blocks per SM of the graph path
        opts["synapse_blocks_per_sm"], opts["ltp_blocks_per_sm"] = side_blocks // 10, side_blocks % 10
        opts["chain_lookahead"] = int(params.pop("chain_lookahead", 0))  # 0 = the library's default
        if bool(params.pop("wave_kernels", False)):
            opts["wave_kernels"] = 1   # captured graph: the chain's load-balanced waves instead of the group-per-item side kernels
        chain = params.pop("chain", None)  # "nthr,syn,ltp,ltp_thr,learn,neuron_regs,stages" (sweeps; 0 = default)
        if chain:
            names = ("chain_neuron_threads", "chain_synapse_blocks", "chain_ltp_blocks", "chain_ltp_threads",
                     "chain_learn_blocks", "chain_neuron_registers", "chain_learn_stages")
            for nme, v in zip(names, (chain.split(",") if isinstance(chain, str) else chain)):
                opts[nme] = int(v)
        cfg = default_config(
            dtype=L.SNN_F64 if dtype == "f64" else L.SNN_F32,
            mode=L.SNN_MODE_EXACT if mode == "exact" else L.SNN_MODE_FAST,
            device=int(device), n_neurons=n_per * int(n_instances), n_per_instance=n_per, n_exc=int(n_exc),
            max_delay=max_delay,
            noise_mode={"none": L.SNN_NOISE_NONE, "replay": L.SNN_NOISE_REPLAY, "philox": L.SNN_NOISE_PHILOX}[noise],
            **params)
        self.cfg = cfg
        self.n_neurons = int(cfg.n_neurons)
        self.n_instances = int(n_instances)
        self._h = C.c_void_p()
        L.check(lib.snn_create(C.byref(cfg), C.byref(self._h)))
        for name, v in opts.items():
            if v or name in ("graph", "persistent"):   # these two are only in `opts` when they were given
                self.set_param(name, v)
        rowptr = np.ascontiguousarray(rowptr, np.int64)
        post = np.ascontiguousarray(post, np.int32)
        w = np.ascontiguousarray(w, np.float64)
        delay = None if delay is None else np.ascontiguousarray(delay, np.uint8)
        if len(rowptr) != self.n_neurons + 1:
            raise ValueError("rowptr must have n_neurons+1 entries")
        L.check(lib.snn_set_synapses_csr(self._h, _ptr(rowptr), _ptr(post), _ptr(w), _ptr(delay)))
        self.n_synapses = int(lib.snn_n_synapses(self._h))

    def set_param(self, name: str, value: float):
        """snn_set_param: a model scalar by its snn_config field name, or a scheduling / diagnostic option."""
        L.check(L.load().snn_set_param(self._h, name.encode(), float(value)))

    def get_param(self, name: str) -> float:
        v = C.c_double()
        L.check(L.load().snn_get_param(self._h, name.encode(), C.byref(v)))
        return v.value

    def checkpoint(self) -> bytes:
        """Everything the next window depends on (delays in flight included), as one blob."""
        lib = L.load()
        n = C.c_int64()
        L.check(lib.snn_checkpoint_bytes(self._h, C.byref(n)))
        buf = C.create_string_buffer(n.value)
        L.check(lib.snn_checkpoint(self._h, buf, n.value))
        return buf.raw

    def restore(self, blob: bytes):
        L.check(L.load().snn_restore(self._h, blob, len(blob)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            L.load().snn_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- step!(net[, input_spikes], train) for a whole window
    def run(self, n_steps, noise=None, train=None, da_events=None, force=None, record=True,
            spikes_cap=None, profile=False, currents=None, probes=None, raw=False):
        """Runs n_steps milliseconds on the device.  noise: [n_steps, n_neurons] float64 or None;
        train: bool array [n_steps] (learn! after that step); da_events: [(step, instance,
        amount)]; force: [(step, neuron)]; currents: [(step, neuron, current)] (injected-current
        input!); probes: synapse indices (caller's CSR order) whose (w, sd) is recorded after every step
        into self.last_probes [n_steps, n_probes, 2].  Returns (list of ascending spike-id arrays, stats); with
        raw=True the record as the C ABI returns it, (ids [n_spikes], offsets [n_steps + 1]), instead of the list."""
        lib = L.load()
        n_steps = int(n_steps)
        if noise is not None:
            noise = np.ascontiguousarray(noise, np.float64)
            if noise.shape != (n_steps, self.n_neurons):
                raise ValueError("noise must be [n_steps, n_neurons]")
        tf = None if train is None else np.ascontiguousarray(train, np.uint8)
        ev = None
        n_da = 0
        if da_events:
            n_da = len(da_events)
            ev = (L.SnnDaEvent * n_da)(*[L.SnnDaEvent(int(s), int(i), float(a)) for s, i, a in da_events])
        fe = None
        n_force = 0
        if force:
            n_force = len(force)
            fe = (L.SnnForceEvent * n_force)(*[L.SnnForceEvent(int(s), int(j)) for s, j in force])
        out = offs = None
        cap = 0
        if record:
            cap = int(spikes_cap or max(1024, min(self.n_neurons * n_steps, 1 << 26)))
            out = np.empty(cap, np.int32)
            offs = np.zeros(n_steps + 1, np.int64)
        st = L.SnnStepStats()
        # profile: False | 'kernels' (events around every kernel class, serial launches) | 'learn'
        lib.snn_set_profiling(self._h, {False: 0, None: 0, True: 1, 'kernels': 1, 'learn': 2}[profile])
        probe_idx = probe_out = None
        if probes is not None and len(probes):
            probe_idx = np.ascontiguousarray(probes, np.int64)
            probe_out = np.zeros((n_steps, len(probe_idx), 2), np.float64)
        if currents or probe_idx is not None:
            currents = currents or []
            n_cur = len(currents)
            ce = (L.SnnCurrentEvent * max(1, n_cur))(*[L.SnnCurrentEvent(int(s_), int(j), float(c)) for s_, j, c in currents])
            io = L.SnnStepIO()
            io.struct_size = C.sizeof(L.SnnStepIO)
            io.n_steps = n_steps
            io.noise = None if noise is None else noise.ctypes.data
            io.train_flags = None if tf is None else tf.ctypes.data
            io.force = C.cast(fe, C.c_void_p) if fe is not None else None
            io.da = C.cast(ev, C.c_void_p) if ev is not None else None
            io.currents = C.cast(ce, C.c_void_p) if n_cur else None
            io.n_force, io.n_da, io.n_currents = n_force, n_da, n_cur
            io.spikes_out = None if out is None else out.ctypes.data
            io.spike_offsets = None if offs is None else offs.ctypes.data
            io.spikes_cap = cap
            io.stats = C.pointer(st)
            if probe_idx is not None:
                io.probe_syn, io.probe_out, io.n_probes = probe_idx.ctypes.data, probe_out.ctypes.data, len(probe_idx)
            rc = lib.snn_step_ex(self._h, C.byref(io))
        else:
            rc = lib.snn_step(self._h, n_steps, _ptr(noise), _ptr(tf), fe, n_force, ev, n_da, _ptr(out), _ptr(offs),
                              cap, C.byref(st))
        L.check(rc)
        stats = StepStats(st.n_spikes, st.n_syn_events, st.n_ltp_events, st.device_ms, st.learn_ms, st.neuron_ms,
                          st.synapse_ms, st.learn_launches, st.neuron_launches, st.synapse_launches, st.reserved)
        spikes = None
        if record and raw:
            spikes = (out[:offs[n_steps]], offs)
        elif record:
            spikes = [out[offs[k]:offs[k + 1]].copy() for k in range(n_steps)]
        if probe_out is not None:
            self.last_probes = probe_out  # [n_steps, n_probes, (w, sd)] after every step
        return spikes, stats

    def get(self, field) -> np.ndarray:
        n = {L.FIELD_W: self.n_synapses, L.FIELD_SD: self.n_synapses, L.FIELD_DA: self.n_instances}.get(
            field, self.n_neurons)
        out = np.zeros(n, np.float64)
        L.check(L.load().snn_get_array(self._h, field, _ptr(out), n))
        return out

    def get_synapses(self, idx):
        """(w, sd) of the synapses `idx` (positions in the CSR arrays the net was built from) — `net.s[n1, syn]`,
        `net.sd[n1, syn]` of newnet_DASTDP.jl:57-60 without copying all E of them."""
        idx = np.ascontiguousarray(idx, np.int64)
        w, sd = np.zeros(idx.size, np.float64), np.zeros(idx.size, np.float64)
        L.check(L.load().snn_get_synapses(self._h, _ptr(idx), idx.size, _ptr(w), _ptr(sd)))
        return w, sd

    def set(self, field, arr):
        a = np.ascontiguousarray(arr, np.float64)
        L.check(L.load().snn_set_array(self._h, field, _ptr(a), a.size))

    def add_dopamine(self, amount, instance=0):
        L.check(L.load().snn_add_dopamine(self._h, int(instance), float(amount)))

    def reset_v(self):
        L.check(L.load().snn_reset_v(self._h))

    @property
    def t(self):
        return int(L.load().snn_step_index(self._h))


class _MatView:
    """`net.s[pre, post]` / `net.sd[pre, post]` on top of the CSR arrays (newnet_DASTDP.jl:19,58)."""

    def __init__(self, net: "Network", field: int):
        self._net, self._field = net, field

    def _pos(self, pre, post):
        lo, hi = self._net._rowptr[pre], self._net._rowptr[pre + 1]
        k = np.searchsorted(self._net._post[lo:hi], post)
        if k < hi - lo and self._net._post[lo + k] == post:
            return lo + k
        return None

    def __getitem__(self, idx):
        pre, post = idx
        k = self._pos(int(pre), int(post))
        return 0.0 if k is None else float(self._net._sparse.get(self._field)[k])

    def __setitem__(self, idx, value):
        pre, post = idx
        k = self._pos(int(pre), int(post))
        if k is None:
            if value != 0.0:
                raise IndexError("no synapse there: connection[pre, post] == 0 (new_net.jl:139 masks it anyway)")
            return
        a = self._net._sparse.get(self._field)
        a[k] = value
        self._net._sparse.set(self._field, a)

    def dense(self) -> np.ndarray:
        n = self._net.n_neurons
        m = np.zeros((n, n))
        pre = np.repeat(np.arange(n), np.diff(self._net._rowptr))
        m[pre, self._net._post] = self._net._sparse.get(self._field)
        return m


class Network:
    """Variant A: `Network(n_neurons, connectivity, n_exc)` (Old_net/new_net.jl:80-98).

    `connection` replaces the unreplayable `rand(N,N) .< connectivity` (line 30): pass the
    recorded matrix, or leave None to draw it from numpy Philox(seed)."""

    def __init__(self, n_neurons, connectivity, n_exc, connection=None, seed=1, dtype="f64", mode="exact",
                 device=0):
        from . import synth
        self.n_neurons, self.n_exc = int(n_neurons), int(n_exc)
        if connection is None:
            connection = synth.dense_connection(self.n_neurons, float(connectivity), seed)
        self.connection = np.asarray(connection, np.int64)
        s0 = synth.variant_a_weights(self.n_neurons, self.n_exc, self.connection)
        self._rowptr, self._post, w = synth.dense_to_csr(s0, self.connection)
        self._sparse = SparseNetwork(self.n_neurons, self.n_exc, self._rowptr, self._post, w, dtype=dtype,
                                     mode=mode, noise="replay", device=device)
        self.exc = np.arange(0, self.n_exc)
        self.inh = np.arange(self.n_exc, self.n_neurons)
        self.s = _MatView(self, L.FIELD_W)
        self.sd = _MatView(self, L.FIELD_SD)

    def post(self, neuron):
        """post(neuron, connection) of newnet_DASTDP.jl:11-13."""
        return self._post[self._rowptr[neuron]:self._rowptr[neuron + 1]].copy()

    # step!(net, train) / step!(net, input_spikes, train)  new_net.jl:146-161
    def step(self, noise, train=False, input_spikes=None):
        force = None if input_spikes is None else [(0, int(j)) for j in input_spikes]
        spikes, _ = self._sparse.run(1, np.asarray(noise, np.float64).reshape(1, -1), [1 if train else 0],
                                     force=force)
        return spikes[0]

    def run(self, noise, train, da_events=None, **kw):
        return self._sparse.run(len(noise), noise, train, da_events=da_events, **kw)

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
    def STDP(self): return self._sparse.get(L.FIELD_TRACE)
    @property
    def I(self): return self._sparse.get(L.FIELD_I)
