"""Host-side mirror of the layered LIF `Network` of Julia/Modules/Networks/net_arch.jl for a whole
population of agents, over the snn_actor_* entry points of libsnn_b200.so.

`ActorPopulation(arch, n_agents, **param)` plays the role of `[Network(arch, param) for _ in
1:n_agents]`; `decision_window` is the body of play_episode's inner loop
(cartpole_fitness.jl:70-79 / teacher_sim.jl:61-80) and `play_episodes` is play_episode itself
(cartpole_fitness.jl:51-114) with the CartPole environment on the device.  `set_weight` /
`ei_idx` / `borne` restate cartpole_fitness.jl:43-46,119-154 on the host (genome decoding, tiny).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def borne(x, a, b):
    """cartpole_fitness.jl:43-46"""
    return a + 0.5 * (b - a) * (1 + np.cos(np.pi * x / 10))


def ei_idx(arch):
    """cartpole_fitness.jl:119-135 (0-based ids): the first div(arch[l],4) neurons of every
    non-final layer are 'inhibitory' (their outgoing weights share one negative value)."""
    inh, exc = [], []
    n_inh = [a // 4 for a in arch[:-1]] + [0]
    off = 0
    for l, a in enumerate(arch):
        for neuron in range(a):
            (inh if neuron < n_inh[l] else exc).append(neuron + off)
        off += a
    return exc, inh


def set_weight(arch, genes_w):
    """set_weight! cartpole_fitness.jl:137-154: genome -> list of e[l] (arch[l] x arch[l+1])."""
    exc, inh = ei_idx(arch)
    exc, inh = set(exc), set(inh)
    genes_w = np.asarray(genes_w, np.float64)
    out = []
    inh_count = 0
    done = 0
    for mi in range(len(arch) - 1):
        nr, nc = arch[mi], arch[mi + 1]
        m = np.zeros((nr, nc))
        offset_neuron = sum(arch[:mi])
        for idx in range(nr * nc):  # column-major linear index, 0-based
            offset = done - inh_count
            line = idx % nr
            cur = line + offset_neuron
            if cur in exc:
                m[line, idx // nr] = borne(genes_w[idx + offset], 0.0, 1.1)
            elif cur in inh:
                m[line, idx // nr] = borne(genes_w[-1], -2.0, 0.0)
                inh_count += 1
        done += nr * nc
        out.append(m)
    return out


class ActorPopulation:
    def __init__(self, arch, n_agents, device=0, **param):
        lib = L.load()
        cfg = L.SnnActorConfig()
        L.check(lib.snn_actor_config_default(C.byref(cfg)))
        cfg.device = int(device)
        cfg.n_agents = int(n_agents)
        cfg.n_layers = len(arch)
        for i, a in enumerate(arch):
            cfg.arch[i] = int(a)
        for k, v in param.items():
            if not hasattr(cfg, k):
                raise AttributeError(f"snn_actor_config has no field {k!r}")
            setattr(cfg, k, v)
        self.cfg = cfg
        self.arch = [int(a) for a in arch]
        self.n_agents = int(n_agents)
        self.n_exc = sum(self.arch)
        self.n_neurons = self.n_exc + sum(self.arch[1:-1])
        self._h = C.c_void_p()
        L.check(lib.snn_actor_create(C.byref(cfg), C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            L.load().snn_actor_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _count(self, field, layer):
        if field in (L.FIELD_V, L.FIELD_HO):
            return self.n_neurons
        if field == L.FIELD_TRACE:
            return self.n_exc
        if field == L.FIELD_DA:
            return 1
        return self.arch[layer] * self.arch[layer + 1]

    def get(self, field, layer=0):
        per = self._count(field, layer)
        out = np.zeros((self.n_agents, per), np.float64)
        L.check(L.load().snn_actor_get_array(self._h, field, layer, _ptr(out), out.size))
        return out

    def set(self, field, arr, layer=0):
        per = self._count(field, layer)
        a = np.ascontiguousarray(np.asarray(arr, np.float64).reshape(self.n_agents, per))
        L.check(L.load().snn_actor_set_array(self._h, field, layer, _ptr(a), a.size))

    def set_weights(self, layer, e):
        """e: [n_agents, arch[l], arch[l+1]] — stored column-major per agent like the Julia e[l]."""
        e = np.asarray(e, np.float64).reshape(self.n_agents, self.arch[layer], self.arch[layer + 1])
        self.set(L.FIELD_W, np.transpose(e, (0, 2, 1)).reshape(self.n_agents, -1), layer)

    def weights(self, layer, field=L.FIELD_W):
        flat = self.get(field, layer)
        return np.transpose(flat.reshape(self.n_agents, self.arch[layer + 1], self.arch[layer]), (0, 2, 1))

    def decision_window(self, intensity, n_steps=40, train=False, critic=False, reset_v=True, spikes=False):
        """n_steps x step_LIF!(net, int_spikes(intensity), train).  Returns (counts, masks, ms)."""
        it = np.ascontiguousarray(np.asarray(intensity, np.float64).reshape(self.n_agents, self.arch[0]))
        counts = np.zeros((self.n_agents, self.arch[-1]), np.int32)
        masks = np.zeros((self.n_agents, n_steps), np.int32) if spikes else None
        ms = C.c_double(0.0)
        L.check(L.load().snn_actor_window(self._h, int(n_steps), _ptr(it), int(train), int(critic), int(reset_v),
                                          _ptr(counts), _ptr(masks), C.byref(ms)))
        return counts, masks, ms.value

    def learn(self, critic=False):
        """learn!(net) / learn_critic!(net) on their own for every agent (net_arch.jl:213-235)."""
        L.check(L.load().snn_actor_learn(self._h, int(critic)))

    def play_episodes(self, matalpha, state0, tie_actions=None, n_steps=200, win_size=40, train=False,
                      teach_actions=None, env="cartpole"):
        """play_episode for every agent, environment on the device.

        env="cartpole" (cartpole_fitness.jl:51-114): matalpha [n_agents, arch[0], 4], state0 [n_agents, 4].
        env="mountaincar" (mountcar_fitness.jl:48-107): matalpha [n_agents, arch[0], 2], state0 [n_agents, 2].
        tie_actions [n_agents, k]: the replayed `rand(rng_action, ...)` stream (k >= n_steps; a MountainCar
        decision can consume three draws); None = the reference's default `rng_seed = true`, i.e. every
        episode draws from a fresh MersenneTwister(0) (julia_rng.py).
        Returns dict(rewards, n_decisions, n_rand, state, device_ms)."""
        envs = {"cartpole": (L.SNN_ENV_CARTPOLE, 4), "mountaincar": (L.SNN_ENV_MOUNTAINCAR, 2)}
        if env not in envs:
            raise ValueError(f"unknown environment {env!r}")
        env_id, nobs = envs[env]
        ma = np.asarray(matalpha, np.float64).reshape(self.n_agents, self.arch[0], nobs)
        ma = np.ascontiguousarray(np.transpose(ma, (0, 2, 1)).reshape(self.n_agents, -1))  # column-major
        s0 = np.zeros((self.n_agents, 4), np.float64)
        s0[:, :nobs] = np.asarray(state0, np.float64).reshape(self.n_agents, nobs)
        shared = False
        if tie_actions is None:
            # the reference's default: every episode of every agent draws from a fresh MersenneTwister(0) — one
            # stream, shared (uploaded once per call, not once per agent)
            from .julia_rng import julia_mt_pick_bits
            k = n_steps * (1 if env_id == L.SNN_ENV_CARTPOLE else 3)
            key = (env_id, k)
            if getattr(self, "_mt0_key", None) != key:
                self._mt0_key, self._mt0 = key, np.ascontiguousarray(julia_mt_pick_bits(0, k), np.int32).reshape(1, -1)
            tie_actions, shared = self._mt0, teach_actions is None
            if not shared:
                tie_actions = np.tile(tie_actions, (self.n_agents, 1))
        ta = np.ascontiguousarray(np.asarray(tie_actions, np.int32).reshape(1 if shared else self.n_agents, -1))
        if ta.shape[1] < n_steps:
            raise ValueError("tie_actions needs at least n_steps entries per agent")
        te = None if teach_actions is None else np.ascontiguousarray(
            np.asarray(teach_actions, np.int32).reshape(self.n_agents, n_steps))
        if te is not None:
            ta = np.ascontiguousarray(ta[:, :n_steps])   # snn_actor_episodes: stride n_steps
        rewards = np.zeros(self.n_agents, np.float64)
        nd = np.zeros(self.n_agents, np.int32)
        nr = np.zeros(self.n_agents, np.int32)
        so = np.zeros((self.n_agents, 4), np.float64)
        ms = C.c_double(0.0)
        if env_id == L.SNN_ENV_CARTPOLE and not shared and (te is not None or ta.shape[1] == n_steps):
            L.check(L.load().snn_actor_episodes(self._h, _ptr(ma), _ptr(s0), _ptr(ta), _ptr(te), int(n_steps),
                                                int(win_size), int(train), _ptr(rewards), _ptr(nd), _ptr(nr),
                                                _ptr(so), C.byref(ms)))
        else:
            if te is not None:
                raise ValueError("teach_actions: CartPole only")
            L.check(L.load().snn_actor_episodes_env(self._h, env_id, _ptr(ma), _ptr(s0), _ptr(ta),
                                                    -int(ta.shape[1]) if shared else int(ta.shape[1]),
                                                    int(n_steps), int(win_size), int(train), _ptr(rewards),
                                                    _ptr(nd), _ptr(nr), _ptr(so), C.byref(ms)))
        return dict(rewards=rewards, n_decisions=nd, n_rand=nr, state=so[:, :nobs], device_ms=ms.value)
