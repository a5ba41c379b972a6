"""Shared builders for the parity tests: the same recorded inputs go to the oracle and to CUDA."""
from __future__ import annotations

import numpy as np

from rl_stdp_b200 import synth


def c1_inputs(n=1000, n_exc=800, p=0.1, t=1000, seed_conn=1, seed_noise=2):
    """connection = Philox(1).uniform(N,N) < 0.1, noise = 13*(Philox(2).uniform-0.5),
    learn every 10th ms (newnet_DASTDP.jl:48-51)."""
    conn = synth.dense_connection(n, p, seed_conn)
    noise = synth.thalamic_noise(t, n, seed_noise)
    train = np.array([(k + 1) % 10 == 0 for k in range(t)], np.uint8)
    return conn, noise, train


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    den = np.maximum(np.abs(b), 1e-300)
    m = np.abs(a - b) / den
    m[(a == b)] = 0.0
    return float(m.max()) if m.size else 0.0


def run_oracle_sparse(ora, noise, train, da_events=(), force=()):
    """Drives oracle.py_oracle.Sparse like snn_step drives the device: da events after the step."""
    t = len(train)
    da_by = {}
    for s, i, a in da_events:
        da_by.setdefault(int(s), []).append((int(i), float(a)))
    f_by = {}
    for s, j in force:
        f_by.setdefault(int(s), []).append(int(j))
    out = []
    for k in range(t):
        out.append(ora.step(None if noise is None else noise[k], bool(train[k]), f_by.get(k)))
        for i, a in da_by.get(k, ()):
            ora.add_da(i, a)
    return out


def _gym_np_random(seed):
    """gym.utils.seeding.np_random of the gym releases that still had env.seed(): hash_seed = first 8 bytes
    of SHA-512(str(seed)) as a little-endian integer, split into 32-bit words for RandomState.seed."""
    import hashlib
    h = int.from_bytes(hashlib.sha512(str(seed).encode()).digest()[:8], "little")
    rs = np.random.RandomState()
    rs.seed([h % 2**32, h >> 32])
    return rs


def gym_seed0_reset_state():
    """The observation gym's classic CartPole returns from `env.seed(0); env.reset()`
    (cartpole_fitness.jl:53-62): gym.utils.seeding hashes the seed with SHA-512, keeps 8 bytes, seeds a
    numpy RandomState with the two little-endian 32-bit words, and CartPoleEnv.reset draws
    uniform(-0.05, 0.05, 4) from it.  Published value: [-0.04456399, 0.04653909, 0.01326909, -0.02099827]."""
    s = _gym_np_random(0).uniform(-0.05, 0.05, size=(4,))
    assert np.allclose(s, [-0.04456399, 0.04653909, 0.01326909, -0.02099827], atol=5e-9), s
    return s


def reference_elites(unique=True):
    """The genomes of Actors/cartpoleElites/sNESelites_cartpole_{1..17}.jld (fixture
    tests/golden/cartpole_elites.npz, made by tests/golden/make_elites.py: 170 saved individuals, 60
    distinct) decoded as fitness() does (cartpole_fitness.jl:158-176, cartpole_tryout.jl:20-38):
    [(matalpha 8x4, [e0, e1], saved fitness)]."""
    import os
    from rl_stdp_b200.lif import borne, set_weight
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cartpole_elites.npz"))
    genes, fit = d["genes"], d["fitness"]
    if unique:
        _, first = np.unique(genes, axis=0, return_index=True)
        first.sort()
        genes, fit = genes[first], fit[first]
    arch = [8, 8, 8]
    out = []
    for g, f in zip(genes, fit):
        m = borne(g[:arch[0] * 4], -1.0, 1.0).reshape((4, arch[0])).T   # reshape(genes_alpha, (arch[1], 4))
        out.append((m, set_weight(arch, g[arch[0] * 4:]), float(f)))
    return out


def mountcar_elites():
    """Actors/sNESelites_mountcar.jld decoded as mountcar_fitness.jl:158-178 does: matalpha = borne(genes[1:16])
    reshaped (8, 2); the weights start at gene 33 (`genes[arch[1]*4+1:end]`, genes 17..32 are unused)."""
    import os
    from rl_stdp_b200.lif import borne, set_weight
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mountcar_elites.npz"))
    arch = [8, 8, 8]
    return [(borne(g[:16], -1.0, 1.0).reshape((2, 8)).T, set_weight(arch, g[32:]), float(f))
            for g, f in zip(d["genes"], d["fitness"])]


def play_mountcar(window_fn, matalpha, tie_bits, n_steps=200):
    """play_episode of Actors/mountcar_fitness.jl:48-107 for A agents in lock-step, with MountainCar-v0 restated
    from gym's classic-control definition (NOT in the reference: PyCall -> gym) and gym's seed-0 reset
    position.  window_fn(intensity [A, 8]) -> last-layer spike counts [A, 8] of one 40-ms decision window
    (v reset to c first, counted from msec >= length(arch) - 1).  tie_bits = the stream of
    rand(MersenneTwister(0), pool) picks (rl_stdp_b200/julia_rng.py); every episode restarts it.
    Returns (reward [A], decisions [A], ties [A])."""
    import math
    A = len(matalpha)
    p0 = _gym_np_random(0).uniform(low=-0.6, high=-0.4)
    pos, vel = np.full(A, p0), np.zeros(A)
    tot, used, nd = np.zeros(A), np.zeros(A, np.int64), np.zeros(A, np.int64)
    live = np.ones(A, bool)
    for _ in range(n_steps):
        if not live.any():
            break
        obs = np.stack([(pos + 1.2) / 1.8, (vel + 0.07) / 0.14], axis=1)          # norm_mountcar :13-20
        cnt = np.asarray(window_fn(np.einsum("aij,aj->ai", np.asarray(matalpha), obs)), np.float64)
        for a in np.flatnonzero(live):
            sl, sm, sr = cnt[a, :2].sum(), cnt[a, 2:4].sum(), cnt[a, 4:].sum()    # div(8,3) = 2, 2, 4
            act, fac = int(np.argmax([sl, sm, sr])), 1                            # push_or_not :33-39
            for d, pool in zip((sl - sr, sl - sm, sm - sr), ((1, 3), (1, 2), (2, 3))):   # :84-93
                if d == 0.0:
                    act, fac = pool[tie_bits[used[a]]] - 1, 2
                    used[a] += 1
            v = vel[a] + ((act - 1) * 0.001 + math.cos(3 * pos[a]) * (-0.0025))   # gym: velocity += A + B
            v = min(max(v, -0.07), 0.07)
            x = min(max(pos[a] + v, -1.2), 0.6)
            if x == -1.2 and v < 0:
                v = 0.0
            pos[a], vel[a] = x, v
            tot[a] -= fac
            nd[a] += 1
            if x >= 0.5 and v >= 0:
                live[a] = False
    return tot, nd, used
