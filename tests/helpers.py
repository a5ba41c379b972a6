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
