"""Independent numpy restatement of the reference hot path (second opinion for the C oracle).

TEST INFRASTRUCTURE ONLY.  Written from the Julia lines directly, in the reference's own
vectorised (broadcast) style, so that a transcription slip in oracle/snn_oracle.c or
oracle/lif_oracle.c shows up as a bit difference between the two.  numpy float64 elementwise
arithmetic is IEEE and unfused, evaluated in the order written — the same contract as Julia
broadcasts without @fastmath.

  NetA      Julia/Izhikevic_dastdp/Old_net/new_net.jl:4-161      (variant A)
  NetB      Julia/Izhikevic_dastdp/Old_net/new_net_col.jl:4-161  (variant B: transposed storage)
  NetArch   Julia/Modules/Networks/net_arch.jl:8-281,316-379      (variants E and F)
  NetColSep Julia/Izhikevic_dastdp/Old_net/new_net_col_sep.jl     (variant C)
  NetRes    Julia/Modules/Networks/net_res.jl                     (variant D)
  NetG      Julia/Izhikevic_dastdp/Old_net/DASTDP.jl + Julia_DA_STDPv2.jl (variant G)
"""
from __future__ import annotations

import numpy as np


class NetA:
    def __init__(self, n_neurons, n_exc, connection):
        N = n_neurons
        self.n_neurons, self.n_exc = N, n_exc
        self.v = -65.0 * np.ones(N)                                   # :81
        self.u = 0.2 * self.v                                         # :82
        self.d = np.where(np.arange(1, N + 1) <= n_exc, 8.0, 2.0)     # :83
        self.a = np.where(np.arange(1, N + 1) <= n_exc, 0.02, 0.1)    # :84
        self.connection = np.array(connection, dtype=np.int64)        # :86 (recorded rand(N,N) .< p)
        self.exc = np.arange(0, n_exc)
        self.inh = np.arange(n_exc, N)
        self.s = np.zeros((N, N))
        self.s[self.exc, :] = 1.0 * np.ones((n_exc, N))               # :89
        self.s[self.inh, :] = -1.0 * np.ones((N - n_exc, N))          # :90
        self.s *= self.connection                                     # :91
        self.sd = np.zeros((N, N))
        self.STDP = np.zeros(N)
        self.I = np.zeros(0)
        self.da = 0.0

    def input(self, input_spikes):                                    # :104-106
        self.v[input_spikes] = 40.0

    def spike(self, noise):                                           # :109-133
        self.I = np.array(noise, dtype=np.float64)                    # :111 recorded 13 .* (rand(N) .- 0.5)
        spiked = np.flatnonzero(self.v > 30)                          # :112
        self.v[spiked] = -65.0                                        # :113
        self.u[spiked] += self.d[spiked]                              # :114
        for neuron in spiked:                                         # :116-118
            self.I += self.s[neuron, :]
        v, u, I = self.v, self.u, self.I
        self.v = v + 0.5 * ((0.04 * v + 5) * v + 140 - u + I)         # :120
        v = self.v
        self.v = v + 0.5 * ((0.04 * v + 5) * v + 140 - u + I)         # :121
        self.u = u + self.a * (0.2 * self.v - u)                      # :122
        self.STDP[spiked] = 0.1                                       # :123
        for neuron in spiked:                                         # :125-128
            self.sd[:, neuron] += 1.0 * self.STDP
            self.sd[neuron, :] -= 1.5 * self.STDP
        self.STDP *= 0.95                                             # :130
        self.da *= 0.995                                              # :131
        return spiked

    def learn(self):                                                  # :135-142
        e = self.exc
        self.s[e, :] += (0.002 + self.da) * self.sd[e, :]
        x = self.s[e, :]
        self.s[e, :] = np.where(x > 4, 4.0, np.where(x < 0, 0.0, x))  # Base.clamp
        self.s *= self.connection
        self.sd *= 0.99

    def step(self, noise, train=False, input_spikes=None):            # :146-161
        if input_spikes is not None:
            self.input(input_spikes)
        spiked = self.spike(noise)
        if train:
            self.learn()
        return spiked


class NetB:
    """Variant B (new_net_col.jl): the same network stored transposed — connection, s and sd are indexed
    [post, pre] ("WARNING read by columns", :13-15), propagation adds COLUMNS (:116-118), the STDP loop
    updates the ROW of a spiker first (LTP of its incoming synapses) and then its COLUMN (LTD of its outgoing
    ones) (:127-130), learn! works on the exc COLUMNS (:138-140).  `connection` is given as [pre, post], as
    connect() draws it, and transposed here as :94 does."""

    def __init__(self, n_neurons, n_exc, connection):
        N = n_neurons
        self.n_neurons, self.n_exc = N, n_exc
        self.v = -65.0 * np.ones(N)                                   # :88
        self.u = 0.2 * self.v                                         # :89
        self.d = np.where(np.arange(1, N + 1) <= n_exc, 8.0, 2.0)     # :90
        self.a = np.where(np.arange(1, N + 1) <= n_exc, 0.02, 0.1)    # :91
        self.connection = np.array(connection, dtype=np.int64).T.copy()   # :93-94 transpose(connection)
        self.exc = np.arange(0, n_exc)
        self.inh = np.arange(n_exc, N)
        self.s = np.zeros((N, N))
        self.s[:, self.exc] = 1.0 * np.ones((N, n_exc))               # :97
        self.s[:, self.inh] = -1.0 * np.ones((N, N - n_exc))          # :98
        self.s *= self.connection                                     # :99
        self.sd = np.zeros((N, N))
        self.STDP = np.zeros(N)
        self.I = np.zeros(0)
        self.da = 0.0

    def input(self, input_spikes):                                    # :106-108
        self.v[input_spikes] = 40.0

    def spike(self, noise):                                           # :111-135
        self.I = np.array(noise, dtype=np.float64)                    # :113 recorded 13 .* (rand(N) .- 0.5)
        spiked = np.flatnonzero(self.v > 30)                          # :114
        self.v[spiked] = -65.0                                        # :115
        self.u[spiked] += self.d[spiked]                              # :116
        for neuron in spiked:                                         # :118-120
            self.I += self.s[:, neuron]
        v, u, I = self.v, self.u, self.I
        self.v = v + 0.5 * ((0.04 * v + 5) * v + 140 - u + I)         # :122
        v = self.v
        self.v = v + 0.5 * ((0.04 * v + 5) * v + 140 - u + I)         # :123
        self.u = u + self.a * (0.2 * self.v - u)                      # :124
        self.STDP[spiked] = 0.1                                       # :125
        for neuron in spiked:                                         # :127-130
            self.sd[neuron, :] += 1.0 * self.STDP
            self.sd[:, neuron] -= 1.5 * self.STDP
        self.STDP *= 0.95                                             # :132
        self.da *= 0.995                                              # :133
        return spiked

    def learn(self):                                                  # :137-144
        e = self.exc
        self.s[:, e] += (0.002 + self.da) * self.sd[:, e]
        x = self.s[:, e]
        self.s[:, e] = np.where(x > 4, 4.0, np.where(x < 0, 0.0, x))  # Base.clamp
        self.s *= self.connection
        self.sd *= 0.99

    def step(self, noise, train=False, input_spikes=None):            # :148-163
        if input_spikes is not None:
            self.input(input_spikes)
        spiked = self.spike(noise)
        if train:
            self.learn()
        return spiked


def _clamp(x, lo, hi):
    return np.where(x > hi, hi, np.where(x < lo, lo, x))


class NetArch:
    """net_arch.jl Network with 0-based ids.  `e0` = list of initial e[l] (the recorded
    clamp.(randn .* std .+ m, 0, wmax) of Weights :21)."""

    def __init__(self, arch, param, e0):
        self.arch = list(arch)
        self.param = dict(param)
        L = len(arch)
        self.n_exc = sum(arch)
        self.n_inh = sum(arch[1:-1])
        self.n_neurons = self.n_exc + self.n_inh
        # Ind :47-67
        architr = [0] + list(arch)
        self.el, self.il = [], []
        for l in range(2, len(architr) + 1):
            self.el.append(np.arange(sum(architr[:l - 1]), sum(architr[:l])))
            if l > 2 and l != len(architr):
                self.il.append(np.arange(sum(architr[1:l - 1]), sum(architr[1:l])) + self.n_exc - arch[0])
            else:
                self.il.append(np.zeros(0, dtype=np.int64))
        self.exc = np.arange(0, self.n_exc)
        # Weights :15-37
        self.e = [np.array(m, dtype=np.float64).copy() for m in e0]
        self.de = [np.zeros((arch[l], arch[l + 1])) for l in range(L - 1)]
        self.ie, self.ei = [], []
        for l in range(1, L + 1):
            if l > 1 and l != L:
                a = arch[l - 1]
                self.ie.append(param["wie"] * (np.ones((a, a)) - np.eye(a)))
                self.ei.append(param["wei"] * np.eye(a))
            else:
                self.ie.append(None)
                self.ei.append(None)
        N = self.n_neurons
        self.v = param["c"] * np.ones(N)                              # :105
        self.u = 0.2 * self.v                                         # :106
        ids1 = np.arange(1, N + 1)
        self.d = np.where(ids1 <= self.n_exc, param["de"], param["di"])
        self.a = np.where(ids1 <= self.n_exc, param["ae"], param["ai"])
        self.trace_e = np.zeros(self.n_exc)
        self.ho = np.zeros(N)
        self.I = np.zeros(N)
        self.da = 0.0
        self.arch_inh = [0 if (l == 0 or l == L - 1) else arch[l] // 4 for l in range(L)]  # :118-125

    # -- shared pieces
    def _propagate(self, spiked):                                     # :164-176 / :338-350
        for neuron in spiked:
            for l in range(len(self.arch) - 1):
                if neuron in self.el[l]:
                    neur = neuron - self.el[l][0]
                    self.I[self.el[l + 1]] += self.e[l][neur, :]
                    if len(self.il[l]):
                        self.I[self.il[l]] += self.ei[l][neur, :]
                if len(self.il[l]) and neuron in self.il[l]:
                    neur = neuron - self.il[l][0]
                    self.I[self.el[l]] += self.ie[l][neur, :]

    def _stdp(self, spiked_e):                                        # :188-200 / :359-371
        L = len(self.arch)
        for neuron in spiked_e:
            for l in range(L):
                if neuron in self.el[l]:
                    neur = neuron - self.el[l][0]
                    if l > 0:
                        self.de[l - 1][:, neur] += self.param["apost"] * self.trace_e[self.el[l - 1]]
                    if l < L - 1:
                        self.de[l][neur, :] -= self.param["apre"] * self.trace_e[self.el[l + 1]]

    # -- LIF (variant F)
    def input_LIF(self, input_spikes):                                # :142-152
        self.I = np.zeros(self.n_neurons)
        max_current = self.param["imin"] + 1.0
        min_current = self.param["imin"]
        indices = np.array([s[0] for s in input_spikes], dtype=np.int64)
        intensity = np.array([s[1] for s in input_spikes], dtype=np.float64)
        self.I[indices] += (max_current - min_current) * intensity + min_current
        l1 = np.arange(0, self.arch[0])
        self.v[l1] = self.v[l1] + (1 / self.param["taulif"]) * (-self.v[l1] + self.I[l1] + self.param["c"])

    def spike_LIF(self):                                              # :154-211
        p = self.param
        self.I = np.zeros(self.n_neurons)
        spiked = np.flatnonzero((self.v - self.ho) > p["vthresh"])
        spiked_e = spiked[spiked < self.n_exc]
        self.ho[spiked_e] += p["theta"]
        self.v[spiked] = p["c"]
        self._propagate(spiked)
        lr = np.arange(self.arch[0], self.n_exc)
        self.v[lr] += self.I[lr]
        self.v[lr] = self.v[lr] + (1 / p["taulif"]) * (-self.v[lr] + p["c"])
        self.v[lr] = np.maximum(self.v[lr], 0.)
        self.trace_e[spiked_e] = 1.0
        self._stdp(spiked_e)
        self.ho *= p["ttheta"]
        self.trace_e *= p["ttrace"]
        self.da *= p["tde"]
        for l in range(len(self.arch) - 1):
            self.de[l] *= p["tde"]
        return spiked

    def learn(self):                                                  # :213-220
        for l in range(len(self.arch) - 1):
            self.e[l] += self.da * self.de[l]
            self.e[l] = _clamp(self.e[l], 0, self.param["wmax"])

    def learn_critic(self):                                           # :222-235
        for l in range(len(self.arch) - 1):
            self.e[l] += self.param["eta"] * self.da * self.de[l]
            self.e[l] = _clamp(self.e[l], 0, self.param["wmax"])
        for l in range(len(self.arch_inh) - 1):
            for i in range(1, self.arch_inh[l] + 1):
                self.e[l][self.arch[l] - i, :] = self.param["winh"]

    def step_LIF(self, input_spikes, train=False):                    # :256-263
        self.input_LIF(input_spikes)
        spiked = self.spike_LIF()
        if train:
            self.learn()
        return spiked

    def step_LIF_critic(self, input_spikes, train=False):             # :265-272
        self.input_LIF(input_spikes)
        spiked = self.spike_LIF()
        if train:
            self.learn_critic()
        return spiked

    # -- Izhikevich layered (variant E)
    def input_force(self, input_spikes):                              # :135-137
        self.v[np.asarray(input_spikes, dtype=np.int64)] = 40.0

    def input_current(self, input_spikes):                            # :316-326
        self.I = np.zeros(self.n_neurons)
        indices = np.array([s[0] for s in input_spikes], dtype=np.int64)
        intensity = np.array([s[1] for s in input_spikes], dtype=np.float64)
        self.I[indices] += (self.param["imax"] - self.param["imin"]) * intensity + self.param["imin"]
        for _ in range(2):
            v = self.v[indices]
            self.v[indices] = v + 0.5 * ((0.04 * v + 5) * v + 140 - self.u[indices] + self.I[indices])

    def spike(self):                                                  # :328-379
        p = self.param
        self.I = np.zeros(self.n_neurons)
        spiked = np.flatnonzero((self.v - self.ho) > p["vthresh"])
        spiked_e = spiked[spiked < self.n_exc]
        spiked_ho = spiked_e[spiked_e >= self.arch[0]]
        self.ho[spiked_ho] += p["theta"]
        self.v[spiked] = p["c"]
        self.u[spiked] += self.d[spiked]
        self._propagate(spiked)
        for _ in range(2):
            v = self.v
            self.v = v + 0.5 * ((0.04 * v + 5) * v + 140 - self.u + self.I)
        self.u = self.u + self.a * (0.2 * self.v - self.u)
        self.trace_e[spiked_e] = 0.1
        self._stdp(spiked_e)
        self.ho *= p["ttheta"]
        self.trace_e *= p["ttrace"]
        self.da *= p["tda"]
        return spiked

    def step(self, input_spikes=None, train=False):                   # :239-281
        if input_spikes is not None and len(input_spikes):
            if isinstance(input_spikes[0], tuple):
                self.input_current(input_spikes)
            else:
                self.input_force(input_spikes)
        spiked = self.spike()
        if train:
            self.learn()
        return spiked


# cartpole_fitness.jl:14-46
def norm_cartpole(obs):
    pi = np.pi
    pos = obs[0]
    theta = np.mod(obs[2] + pi, 2 * pi) - pi
    cone = 12 * pi / 180
    return np.array([(pos + 2.4) / 4.8, (obs[1] + 7.5) / 15, (theta + cone) / (2 * cone), (obs[3] + 7.5) / 15])


def borne(x, a, b):
    return a + 0.5 * (b - a) * (1 + np.cos(np.pi * x / 10))


class NetColSep:
    """Variant C — Old_net/new_net_col_sep.jl:4-187 (storage s[post, pre], split s_exc / s_inh with a
    fixed -0.5 inhibitory weight, homeostasis, STDP among excitatory neurons only).  0-based ids.
    `connection[pre, post]` is the recorded `rand(N,N) .< connectivity` of connect() (:32); the
    constructor transposes it (:102).  Line 170 of the source indexes `STDP[1:net.exc]` with the
    Array `exc`, which raises a MethodError in Julia; it is restated as the evidently intended
    `1:net.n_exc` (the same range as the line before it)."""

    def __init__(self, n_neurons, n_exc, connection):
        N = n_neurons
        self.n_neurons, self.n_exc = N, n_exc
        self.v = -65.0 * np.ones(N)                                           # :93
        self.u = 0.2 * self.v                                                 # :94
        ids1 = np.arange(1, N + 1)
        self.d = np.where(ids1 <= n_exc, 8.0, 2.0)                            # :95
        self.a = np.where(ids1 <= n_exc, 0.02, 0.1)                           # :96
        conn = np.array(connection, dtype=np.int64).T                         # :99 transpose -> [post, pre]
        self.connection_exc = conn[:, :n_exc]                                 # :100
        self.connection_inh = conn[:, n_exc:]                                 # :101
        self.s_exc = 1.0 * np.ones((N, n_exc)) * self.connection_exc          # :103,105
        self.s_inh = -0.5 * np.ones((N, N - n_exc)) * self.connection_inh     # :104,106
        self.sd = np.zeros((N, n_exc))                                        # :107
        self.STDP = np.zeros(N)
        self.ho = np.zeros(N)
        self.I = np.zeros(0)
        self.da = 0.0
        self.arch = [0]

    def input(self, input_spikes):                                            # :127-129
        self.v[input_spikes] = 40.0

    def input_current(self, indices, intensity, noise):                       # :132-141
        """`noise` = the recorded 13 .* (rand(N) .- 0.5) this method draws for itself (:133)."""
        self.I = np.array(noise, dtype=np.float64)
        max_current, min_current = 20.0, 0.0
        idx = np.asarray(indices)
        self.I[idx] += (max_current - min_current) * np.asarray(intensity, dtype=np.float64) + min_current
        for _ in range(2):                                                    # :139-140
            v, u, I = self.v[idx], self.u[idx], self.I[idx]
            self.v[idx] = v + 0.5 * ((0.04 * v + 5) * v + 140 - u + I)

    def spike(self, noise):                                                   # :144-178
        n_exc = self.n_exc
        self.I = np.array(noise, dtype=np.float64)                            # :146
        spiked = np.flatnonzero(self.v - self.ho > 30)                        # :147
        spiked_ho = spiked[(spiked + 1) > self.arch[0]]                       # :149 (1-based id > arch[1])
        self.ho[spiked_ho] += 2.0                                             # :150
        self.v[spiked] = -65.0                                                # :151
        self.u[spiked] += self.d[spiked]                                      # :152
        for neuron in spiked:                                                 # :154-160
            if neuron < n_exc:
                self.I += self.s_exc[:, neuron]
            else:
                self.I += self.s_inh[:, neuron - n_exc]
        v, u, I = self.v, self.u, self.I
        self.v = v + 0.5 * ((0.04 * v + 5) * v + 140 - u + I)                 # :162
        v = self.v
        self.v = v + 0.5 * ((0.04 * v + 5) * v + 140 - u + I)                 # :163
        self.u = u + self.a * (0.2 * self.v - u)                              # :164
        self.STDP[spiked] = 0.1                                               # :165
        for neuron in spiked:                                                 # :167-172
            if neuron < n_exc:
                self.sd[neuron, :] += 1.0 * self.STDP[:n_exc]
                self.sd[:n_exc, neuron] -= 1.5 * self.STDP[:n_exc]
        self.ho *= 0.9999                                                     # :174
        self.STDP *= 0.95                                                     # :175
        self.da *= 0.995                                                      # :176
        return spiked

    def learn(self):                                                          # :180-187
        self.s_exc += (0.002 + self.da) * self.sd
        self.s_exc = _clamp(self.s_exc, 0.0, 4.0)
        self.s_exc *= self.connection_exc
        self.sd *= 0.99

    def step(self, noise, train=False, input_spikes=None, input_current=None):  # :191-214
        if input_spikes is not None:
            self.input(input_spikes)
        if input_current is not None:
            self.input_current(*input_current)
        spiked = self.spike(noise)
        if train:
            self.learn()
        return spiked


class NetRes:
    """Variant D — Modules/Networks/net_res.jl:6-251 (ee / ei / ie split, no thalamic noise,
    parameters from the YAML dict).  0-based ids; `connection[pre, post]` is the recorded matrix of
    connect() (:36-76: `rand(N,N) .< connectivity` with the inh->inh block zeroed for arch == [0],
    the layer blocks + recorded lateral `mat_ie` otherwise)."""

    def __init__(self, n_neurons, n_exc, connection, arch, param):
        N = n_neurons
        self.n_neurons, self.n_exc, self.arch, self.param = N, n_exc, list(arch), dict(param)
        self.v = param["c"] * np.ones(N)                                      # :80
        self.u = 0.2 * self.v                                                 # :81
        ids1 = np.arange(1, N + 1)
        self.d = np.where(ids1 <= n_exc, param["de"], param["di"])            # :82
        self.a = np.where(ids1 <= n_exc, param["ae"], param["ai"])            # :83
        conn = np.array(connection, dtype=np.int64)
        self.connection_ee = conn[:n_exc, :n_exc]                             # :86
        self.connection_ei = conn[:n_exc, n_exc:]                             # :87
        self.connection_ie = conn[n_exc:, :n_exc]                             # :88
        n_inh = N - n_exc
        self.s_ee = param["wee"] * np.ones((n_exc, n_exc)) * self.connection_ee   # :90,93
        self.s_ei = param["wei"] * np.ones((n_exc, n_inh)) * self.connection_ei   # :91,94
        self.s_ie = param["wie"] * np.ones((n_inh, n_exc)) * self.connection_ie   # :92,95
        self.sd_ee = np.zeros((n_exc, n_exc))
        self.sd_ei = np.zeros((n_exc, n_inh))
        self.trace_e = np.zeros(n_exc)
        self.trace_i = np.zeros(n_inh)
        self.ho = np.zeros(N)
        self.I = np.zeros(N)
        self.da = 0.0

    def input(self, input_spikes):                                            # :145-147
        self.v[input_spikes] = 40.0

    def input_current(self, indices, intensity):                              # :150-160
        self.I = np.zeros(self.n_neurons)
        idx = np.asarray(indices)
        self.I[idx] += (self.param["imax"] - self.param["imin"]) * np.asarray(intensity, dtype=np.float64) + self.param["imin"]
        for _ in range(2):
            v, u, I = self.v[idx], self.u[idx], self.I[idx]
            self.v[idx] = v + 0.5 * ((0.04 * v + 5) * v + 140 - u + I)

    def spike(self):                                                          # :163-209
        n_exc, P = self.n_exc, self.param
        random_graph = self.arch == [0]
        self.I = np.zeros(self.n_neurons)                                     # :166
        spiked = np.flatnonzero(self.v - self.ho > 30)                        # :167
        spiked_e = spiked[spiked < n_exc]
        spiked_i = spiked[spiked >= n_exc]
        spiked_ho = spiked_e[(spiked_e + 1) > self.arch[0]]                   # :170
        self.ho[spiked_ho] += P["theta"]                                      # :171
        self.v[spiked] = P["c"]                                               # :172
        self.u[spiked] += self.d[spiked]                                      # :173
        for neuron in spiked:                                                 # :175-182
            if neuron < n_exc:
                self.I[:n_exc] += self.s_ee[neuron, :]
                self.I[n_exc:] += self.s_ei[neuron, :]
            else:
                self.I[:n_exc] += self.s_ie[neuron - n_exc, :]
        v, u, I = self.v, self.u, self.I
        self.v = v + 0.5 * ((0.04 * v + 5) * v + 140 - u + I)                 # :184
        v = self.v
        self.v = v + 0.5 * ((0.04 * v + 5) * v + 140 - u + I)                 # :185
        self.u = u + self.a * (0.2 * self.v - u)                              # :186
        self.trace_e[spiked_e] = 0.1                                          # :187
        self.trace_i[spiked_i - n_exc] = 0.1                                  # :188
        for neuron in spiked_e:                                               # :190-196
            self.sd_ee[:, neuron] += P["apost"] * self.trace_e
            self.sd_ee[neuron, :] -= P["apre"] * self.trace_e
            if random_graph:
                self.sd_ei[neuron, :] -= P["apre"] * self.trace_i
        if random_graph:                                                      # :197-201
            for neuron in spiked_i:
                self.sd_ei[:, neuron - n_exc] += P["apost"] * self.trace_e
        self.ho *= P["ttheta"]                                                # :204
        self.trace_e *= 0.95
        self.trace_i *= 0.95
        self.da *= P["tda"]                                                   # :207
        return spiked

    def learn(self):                                                          # :211-223
        P = self.param
        self.s_ee += (0.002 + self.da) * self.sd_ee
        self.s_ee = _clamp(self.s_ee, 0.0, P["wmax"])
        self.s_ee *= self.connection_ee
        self.sd_ee *= 0.99
        if self.arch == [0]:
            self.s_ei += (0.002 + self.da) * self.sd_ei
            self.s_ei = _clamp(self.s_ei, 0.0, P["wmax"])
            self.s_ei *= self.connection_ei
            self.sd_ei *= 0.99

    def step(self, train=False, input_spikes=None, input_current=None):       # :227-251
        if input_spikes is not None:
            self.input(input_spikes)
        if input_current is not None:
            self.input_current(*input_current)
        spiked = self.spike()
        if train:
            self.learn()
        return spiked


class NetG:
    """Variant G — Old_net/DASTDP.jl:24-106 + NeuronModel.jl:7-26 driven as Julia_DA_STDPv2.jl:71-96
    (fan-out table `post[N, M]`, delay D = 1, threshold `>=`).  0-based ids.  `post` is the recorded
    `vcat(rand(1:N,Ne,M), rand(1:Ne,Ni,M))` (:32).  The trace is kept as the two live columns of
    STDP[N, 1001+D]: `prev` = column msec (after the previous ms's spikes), `cur` = column msec+D."""

    def __init__(self, post, n_exc, s0=None):
        post = np.asarray(post, dtype=np.int64)
        self.N, self.M = post.shape
        self.Ne = n_exc
        self.post = post
        self.v = -65.0 * np.ones(self.N)                                      # Julia_DA_STDPv2.jl:51
        self.u = 0.2 * self.v
        self.s = np.vstack([1.0 * np.ones((n_exc, self.M)), -1.0 * np.ones((self.N - n_exc, self.M))]) if s0 is None \
            else np.array(s0, dtype=np.float64)                               # :53
        self.sd = np.zeros((self.N, self.M))
        ids1 = np.arange(1, self.N + 1)
        self.a = np.where(ids1 <= n_exc, 0.02, 0.1)                           # NeuronModel.jl:9
        self.d = np.where(ids1 <= n_exc, 8.0, 2.0)                            # NeuronModel.jl:10
        self.prev = np.zeros(self.N)
        self.cur = np.zeros(self.N)
        self.DA = 0.0
        # pre[i] = excitatory synapses (row, col) onto i   DASTDP.jl:35-43
        self.pre = [np.argwhere((post == i) & (np.arange(self.N)[:, None] < n_exc)) for i in range(self.N)]

    def step(self, noise, msec):
        """One pass of the ms loop body (Julia_DA_STDPv2.jl:73-86); `msec` decides synweight_step."""
        I = np.array(noise, dtype=np.float64)                                 # :73
        fired = np.flatnonzero(self.v >= 30)                                  # :75  thresh = 30, >=
        if len(fired):                                                        # NeuronModel.jl:19-24
            self.v[fired] = -65.0
            self.u[fired] = self.u[fired] + self.d[fired]
            self.cur[fired] = 0.1                                             # DASTDP.jl:54-59 STDP[fired,msec+D]
        for f in fired:                                                       # DASTDP.jl:62-68 LTP
            rc = self.pre[f]
            if len(rc):
                self.sd[rc[:, 0], rc[:, 1]] = self.sd[rc[:, 0], rc[:, 1]] + 1.0 * self.prev[rc[:, 0]]
        for f in fired[::-1]:                                                 # DASTDP.jl:70-80 LTD, firings walked backwards
            ind = self.post[f, :]
            I[ind] = I[ind] + self.s[f, :]          # repeated targets inside one row: the last assignment wins, as in Julia
            self.sd[f, :] = self.sd[f, :] - (1.5 * self.cur[ind])
        v, u = self.v, self.u                                                 # NeuronModel.jl:12-17
        v = v + 0.5 * ((0.04 * v + 5) * v + 140 - u + I)
        v = v + 0.5 * ((0.04 * v + 5) * v + 140 - u + I)
        u = u + self.a * (0.2 * v - u)
        self.v, self.u = v, u
        self.prev = self.cur                                                  # DASTDP.jl:82-86
        self.cur = 0.95 * self.cur
        self.DA = self.DA * 0.995
        if msec % 10 == 0:                                                    # Julia_DA_STDPv2.jl:83-85, DASTDP.jl:88-92
            Ne = self.Ne
            self.s[:Ne, :] = _clamp(self.s[:Ne, :] + ((0.002 + self.DA) * self.sd[:Ne, :]), 0.0, 4.0)
            self.sd = 0.99 * self.sd
        return fired
