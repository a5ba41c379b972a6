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


def gym_seed0_reset_state():
    """The observation gym's classic CartPole returns from `env.seed(0); env.reset()`
    (cartpole_fitness.jl:53-62): gym.utils.seeding hashes the seed with SHA-512, keeps 8 bytes, seeds a
    numpy RandomState with the two little-endian 32-bit words, and CartPoleEnv.reset draws
    uniform(-0.05, 0.05, 4) from it.  Published value: [-0.04456399, 0.04653909, 0.01326909, -0.02099827]."""
    import hashlib
    h = int.from_bytes(hashlib.sha512(b"0").digest()[:8], "little")
    rs = np.random.RandomState()
    rs.seed([h % 2**32, h >> 32])
    s = rs.uniform(-0.05, 0.05, size=(4,))
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
