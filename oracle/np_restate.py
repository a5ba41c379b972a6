"""Independent numpy restatement of the reference hot path (second opinion for the C oracle).

TEST INFRASTRUCTURE ONLY.  Written from the Julia lines directly, in the reference's own
vectorised (broadcast) style, so that a transcription slip in oracle/snn_oracle.c or
oracle/lif_oracle.c shows up as a bit difference between the two.  numpy float64 elementwise
arithmetic is IEEE and unfused, evaluated in the order written — the same contract as Julia
broadcasts without @fastmath.

  NetA      Julia/Izhikevic_dastdp/Old_net/new_net.jl:4-161      (variant A)
  NetArch   Julia/Modules/Networks/net_arch.jl:8-281,316-379      (variants E and F)
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
