"""ctypes front-end of the CPU oracle (oracle/snn_oracle.c, oracle/lif_oracle.c).

TEST INFRASTRUCTURE ONLY — imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs, never by the product package rl_stdp_b200/.
The oracle is a restatement of the Julia reference (file:line cited in the C sources), not the
Julia binary: "parity unpinned" by the reference's own tests (it has none) — except the LIF actor
episode path, which reproduces the fitness values the reference saved with its sNES elites (60 CartPole genomes, 10 MountainCar genomes)
(tests/test_oracle_cpu.py::test_reference_elites_known_answers).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# ORACLE_OMP=1 (set before the first use) loads the OpenMP build: same sources, same results, the
# independent loops of the sparse step split over the host cores (bench.py --impl reference)
_OMP = os.environ.get("ORACLE_OMP", "") == "1"
_SO = os.path.join(_HERE, "_build", "libsnn_oracle_omp.so" if _OMP else "libsnn_oracle.so")

# order of the double[] parameter block of ora_sparse_create
PARAM_ORDER = [
    "threshold_ge", "vthresh", "v_reset", "a_exc", "a_inh", "d_exc", "d_inh", "trace_set",
    "trace_decay", "apost", "apre", "da_decay", "learn_base", "eta", "wmin", "wmax", "sd_decay",
    "sd_decay_mode", "rate_mode", "plastic_ei", "theta", "ttheta", "ho_from", "ho_inh", "variant_g",
    "lazy", "lazy_da",   # event-driven learning of the production build (0 = the reference's rule)
]
VARIANT_A = dict(
    threshold_ge=0, vthresh=30.0, v_reset=-65.0, a_exc=0.02, a_inh=0.1, d_exc=8.0, d_inh=2.0,
    trace_set=0.1, trace_decay=0.95, apost=1.0, apre=1.5, da_decay=0.995, learn_base=0.002,
    eta=0.0, wmin=0.0, wmax=4.0, sd_decay=0.99, sd_decay_mode=0, rate_mode=0, plastic_ei=1,
    theta=0.0, ttheta=1.0, ho_from=0, ho_inh=0, variant_g=0, lazy=0, lazy_da=0.0,
)
LAY_PARAM_ORDER = [
    "vthresh", "c", "ae", "ai", "de", "di", "wei", "wie", "wmax", "theta", "ttheta", "imin", "imax",
    "apost", "apre", "tda", "tde", "ttrace", "taulif", "eta", "winh",
]
# Julia/Actors/cartpole_cfglif.yml:7-41
CARTPOLE_CFGLIF = dict(
    vthresh=1.0, c=0.0, ae=0.02, ai=0.1, de=8.0, di=2.0, wei=0.0, wie=0.0, wmax=1.1, theta=0.0,
    ttheta=0.998, imin=1.0, imax=0.0, apost=0.1, apre=0.15, tda=0.995, tde=0.99, ttrace=0.9,
    taulif=5.0, eta=0.001, winh=0.0,
)


def build(force: bool = False) -> str:
    global _SO, _OMP
    srcs = [os.path.join(_HERE, f) for f in ("snn_oracle.c", "lif_oracle.c")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    if _OMP and not os.path.exists(_SO):  # no OpenMP toolchain on this host: the serial build
        _OMP = False
        _SO = os.path.join(_HERE, "_build", "libsnn_oracle.so")
    return _SO


def omp_threads() -> int:
    """Threads the loaded oracle library really uses for ora_sparse_step, asked of the OpenMP
    runtime (1 for the serial build).  torchrun exports OMP_NUM_THREADS=1 to its workers, so the
    host's core count says nothing about it."""
    return int(lib().ora_omp_max_threads())


def use_all_cores() -> int:
    """bench.py --impl reference: give the OpenMP build every core this process may run on,
    whatever OMP_NUM_THREADS the launcher exported.  Returns the thread count in use."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:  # pragma: no cover
        n = os.cpu_count() or 1
    lib().ora_omp_set_threads(n)
    return omp_threads()


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        vp, i64, i32, dbl = C.c_void_p, C.c_int64, C.c_int32, C.c_double
        L.ora_omp_max_threads.restype = C.c_int
        L.ora_omp_set_threads.argtypes = [C.c_int]
        L.ora_dense_create.restype = vp
        L.ora_dense_create.argtypes = [i64, i64, vp]
        L.ora_dense_destroy.argtypes = [vp]
        L.ora_dense_field.restype = C.POINTER(dbl)
        L.ora_dense_field.argtypes = [vp, C.c_int]
        L.ora_dense_step.restype = i64
        L.ora_dense_step.argtypes = [vp, vp, vp, i64, C.c_int, vp]
        L.ora_sparse_create.restype = vp
        L.ora_sparse_create.argtypes = [i64, i64, i64, i32, vp, vp, vp, vp, vp]
        L.ora_sparse_destroy.argtypes = [vp]
        L.ora_sparse_get.argtypes = [vp, C.c_int, vp]
        L.ora_sparse_set.argtypes = [vp, C.c_int, vp]
        L.ora_sparse_add_da.argtypes = [vp, i64, dbl]
        L.ora_sparse_counters.argtypes = [vp, vp]
        L.ora_sparse_step.restype = i64
        L.ora_sparse_step.argtypes = [vp, vp, vp, i64, C.c_int, vp]
        L.ora_philox_noise.restype = dbl
        L.ora_philox_noise.argtypes = [C.c_uint64, i64, i64, dbl]
        L.ora_philox_noise_fill.argtypes = [C.c_uint64, i64, i64, dbl, vp]
        L.ora_philox4x32.argtypes = [C.c_uint32] * 6 + [vp]
        L.ora_lay_create.restype = vp
        L.ora_lay_create.argtypes = [i32, i32, vp, vp]
        L.ora_lay_destroy.argtypes = [vp]
        L.ora_lay_n.restype = i32
        L.ora_lay_n.argtypes = [vp]
        L.ora_lay_field.restype = C.POINTER(dbl)
        L.ora_lay_field.argtypes = [vp, C.c_int, C.c_int]
        L.ora_lay_step_lif.restype = C.c_int
        L.ora_lay_step_lif.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_int, vp]
        L.ora_lay_input_force.argtypes = [vp, vp, C.c_int]
        L.ora_lay_input_izh.argtypes = [vp, vp, vp, C.c_int]
        L.ora_lay_spike_izh.restype = C.c_int
        L.ora_lay_spike_izh.argtypes = [vp, vp]
        L.ora_lay_learn.argtypes = [vp]
        L.ora_lay_learn_critic.argtypes = [vp]
        L.ora_cartpole_step.restype = C.c_int
        L.ora_cartpole_step.argtypes = [vp, C.c_int]
        L.ora_norm_cartpole.argtypes = [vp, vp]
        L.ora_play_episode.restype = dbl
        L.ora_play_episode.argtypes = [vp, vp, vp, vp, vp, dbl, C.c_int, C.c_int, vp, vp, vp]
        L.ora_ksin.restype = dbl
        L.ora_ksin.argtypes = [dbl]
        L.ora_cos_reduced.restype = dbl
        L.ora_cos_reduced.argtypes = [dbl]
        L.ora_mountcar_step.restype = C.c_int
        L.ora_mountcar_step.argtypes = [vp, C.c_int]
        L.ora_play_episode_mountcar.restype = dbl
        L.ora_play_episode_mountcar.argtypes = [vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, vp, vp, vp]
        L.ora_kcos.restype = dbl
        L.ora_kcos.argtypes = [dbl]
        L.ora_exp.restype = dbl
        L.ora_exp.argtypes = [dbl]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def exp_fdlibm(x: float) -> float:
    """exp as the device encoders evaluate it (oracle/lif_oracle.c::ora_exp)"""
    return float(lib().ora_exp(float(x)))


class DenseA:
    """Variant A, literal (Old_net/new_net.jl:80-161)."""

    def __init__(self, N: int, Ne: int, connection: np.ndarray):
        assert connection.shape == (N, N)
        self.N, self.Ne = N, Ne
        conn = np.asfortranarray(connection.astype(np.int64))
        self._h = lib().ora_dense_create(N, Ne, _p(conn))
        self._spk = np.zeros(N, np.int32)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().ora_dense_destroy(self._h)
            self._h = None

    def _vec(self, which, n):
        return np.ctypeslib.as_array(lib().ora_dense_field(self._h, which), shape=(n,))

    @property
    def v(self): return self._vec(0, self.N)
    @property
    def u(self): return self._vec(1, self.N)
    @property
    def STDP(self): return self._vec(2, self.N)
    @property
    def I(self): return self._vec(7, self.N)
    @property
    def s(self):  # s[pre, post]
        return self._vec(4, self.N * self.N).reshape(self.N, self.N, order="F")
    @property
    def sd(self):
        return self._vec(5, self.N * self.N).reshape(self.N, self.N, order="F")
    @property
    def da(self): return float(self._vec(6, 1)[0])
    @da.setter
    def da(self, x): self._vec(6, 1)[0] = x

    def step(self, noise, train=False, force=None):
        noise = None if noise is None else np.ascontiguousarray(noise, np.float64)
        f = None if force is None else np.ascontiguousarray(force, np.int32)
        k = lib().ora_dense_step(self._h, _p(noise), _p(f), 0 if f is None else len(f), int(train), _p(self._spk))
        return self._spk[:k].copy()


class Sparse:
    """Sparse + axonal-delay generalisation (what the CUDA kernels implement)."""

    def __init__(self, N, Ne, rowptr, post, w, delay=None, max_delay=1, n_per_instance=None, **params):
        self.N = int(N)
        self.Nper = int(n_per_instance or N)
        self.Ne = int(Ne)
        self.rowptr = np.ascontiguousarray(rowptr, np.int64)
        self.post = np.ascontiguousarray(post, np.int32)
        self.w0 = np.ascontiguousarray(w, np.float64)
        self.delay = None if delay is None else np.ascontiguousarray(delay, np.uint8)
        self.E = int(self.rowptr[-1])
        p = dict(VARIANT_A)
        p.update(params)
        self.params = p
        pv = np.array([float(p[k]) for k in PARAM_ORDER], np.float64)
        self._h = lib().ora_sparse_create(self.N, self.Nper, self.Ne, int(max_delay), _p(self.rowptr),
                                          _p(self.post), _p(self.w0), _p(self.delay), _p(pv))
        self._spk = np.zeros(self.N, np.int32)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().ora_sparse_destroy(self._h)
            self._h = None

    def get(self, field: int) -> np.ndarray:
        n = {4: self.E, 5: self.E, 6: self.N // self.Nper}.get(field, self.N)
        out = np.zeros(n, np.float64)
        lib().ora_sparse_get(self._h, field, _p(out))
        return out

    def set(self, field: int, arr):
        a = np.ascontiguousarray(arr, np.float64)
        lib().ora_sparse_set(self._h, field, _p(a))

    def add_da(self, inst, x):
        lib().ora_sparse_add_da(self._h, int(inst), float(x))

    def counters(self):
        out = np.zeros(2, np.int64)
        lib().ora_sparse_counters(self._h, _p(out))
        return int(out[0]), int(out[1])

    def step(self, noise, train=False, force=None):
        noise = None if noise is None else np.ascontiguousarray(noise, np.float64)
        f = None if force is None else np.ascontiguousarray(force, np.int32)
        k = lib().ora_sparse_step(self._h, _p(noise), _p(f), 0 if f is None else len(f), int(train), _p(self._spk))
        return self._spk[:k].copy()


def philox_noise(seed: int, t: int, n: int, amp: float = 13.0) -> np.ndarray:
    out = np.zeros(n, np.float64)
    lib().ora_philox_noise_fill(int(seed), int(t), int(n), float(amp), _p(out))
    return out


class Layered:
    """Variants E/F (Modules/Networks/net_arch.jl)."""

    def __init__(self, arch, model="lif", **params):
        self.arch = [int(a) for a in arch]
        p = dict(CARTPOLE_CFGLIF)
        p.update(params)
        self.params = p
        pv = np.array([float(p[k]) for k in LAY_PARAM_ORDER], np.float64)
        a = np.array(self.arch, np.int32)
        self._h = lib().ora_lay_create(0 if model == "lif" else 1, len(self.arch), _p(a), _p(pv))
        self.n = lib().ora_lay_n(self._h)
        self.n_exc = sum(self.arch)
        self._spk = np.zeros(self.n, np.int32)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().ora_lay_destroy(self._h)
            self._h = None

    def _vec(self, which, n, l=0):
        return np.ctypeslib.as_array(lib().ora_lay_field(self._h, which, l), shape=(n,))

    @property
    def v(self): return self._vec(0, self.n)
    @property
    def u(self): return self._vec(1, self.n)
    @property
    def trace_e(self): return self._vec(2, self.n_exc)
    @property
    def ho(self): return self._vec(3, self.n)
    @property
    def I(self): return self._vec(7, self.n)
    def e(self, l):
        return self._vec(4, self.arch[l] * self.arch[l + 1], l).reshape(self.arch[l], self.arch[l + 1], order="F")
    def de(self, l):
        return self._vec(5, self.arch[l] * self.arch[l + 1], l).reshape(self.arch[l], self.arch[l + 1], order="F")
    @property
    def da(self): return float(self._vec(6, 1)[0])
    @da.setter
    def da(self, x): self._vec(6, 1)[0] = x

    def step_lif(self, idx, intensity, train=False, critic=False):
        idx = np.ascontiguousarray(idx, np.int32)
        it = np.ascontiguousarray(intensity, np.float64)
        k = lib().ora_lay_step_lif(self._h, _p(idx), _p(it), len(idx), int(train), int(critic), _p(self._spk))
        return self._spk[:k].copy()

    def learn(self, critic=False):
        """learn!(net) / learn_critic!(net) on their own (net_arch.jl:213-235)"""
        (lib().ora_lay_learn_critic if critic else lib().ora_lay_learn)(self._h)

    def step_izh(self, idx=None, intensity=None, train=False):
        if idx is not None:
            idx = np.ascontiguousarray(idx, np.int32)
            if intensity is None:
                lib().ora_lay_input_force(self._h, _p(idx), len(idx))
            else:
                it = np.ascontiguousarray(intensity, np.float64)
                lib().ora_lay_input_izh(self._h, _p(idx), _p(it), len(idx))
        k = lib().ora_lay_spike_izh(self._h, _p(self._spk))
        if train:
            lib().ora_lay_learn(self._h)
        return self._spk[:k].copy()

    def play_episode(self, matalpha, state0, tie_actions, n_steps=200, train=False, teach_actions=None, da_inc=0.0):
        ma = np.asfortranarray(matalpha, np.float64)
        s0 = np.ascontiguousarray(state0, np.float64)
        ta = np.ascontiguousarray(tie_actions, np.int32)
        te = None if teach_actions is None else np.ascontiguousarray(teach_actions, np.int32)
        nd, nr = C.c_int32(0), C.c_int32(0)
        so = np.zeros(4, np.float64)
        r = lib().ora_play_episode(self._h, _p(ma), _p(s0), _p(ta), _p(te), float(da_inc), int(n_steps), int(train),
                                   C.byref(nd), C.byref(nr), _p(so))
        return float(r), nd.value, nr.value, so

    def play_episode_mountcar(self, matalpha, state0, tie_picks, n_steps=200, train=False):
        ma = np.asfortranarray(matalpha, np.float64)
        s0 = np.ascontiguousarray(state0, np.float64)
        tp = np.ascontiguousarray(tie_picks, np.int32)
        nd, nr = C.c_int32(0), C.c_int32(0)
        so = np.zeros(2, np.float64)
        r = lib().ora_play_episode_mountcar(self._h, _p(ma), _p(s0), _p(tp), len(tp), int(n_steps), int(train),
                                            C.byref(nd), C.byref(nr), _p(so))
        return float(r), nd.value, nr.value, so


def cos_reduced(x):
    return float(lib().ora_cos_reduced(float(x)))


def mountcar_step(state, action):
    s = np.ascontiguousarray(state, np.float64).copy()
    done = lib().ora_mountcar_step(_p(s), int(action))
    return s, bool(done)


def cartpole_step(state, action):
    s = np.ascontiguousarray(state, np.float64).copy()
    done = lib().ora_cartpole_step(_p(s), int(action))
    return s, bool(done)
