"""Synthetic networks and recorded input streams for the configurations of SURVEY.md §8(d).

Every random quantity that enters the hot path (connectivity `rand(N,N) .< p`
new_net.jl:30, thalamic noise `13 .* (rand(N) .- 0.5)` new_net.jl:111, reward delays
`rand(1:2000)` newnet_DASTDP.jl:31) is an INPUT recorded on the host with a counter-based RNG
(numpy Philox) and replayed into both the oracle and the CUDA library.
"""
from __future__ import annotations

import numpy as np


def _rng(seed: int) -> np.random.Generator:
    return np.random.Generator(np.random.Philox(key=int(seed)))


def dense_connection(n: int, p: float, seed: int = 1) -> np.ndarray:
    """connection[pre, post] = rand(N,N) .< p  (new_net.jl:30); int64 N x N."""
    return (_rng(seed).random((n, n)) < p).astype(np.int64)


def dense_to_csr(s: np.ndarray, connection: np.ndarray):
    """CSR by presynaptic neuron of s[pre, post] masked by connection (ascending post)."""
    n = s.shape[0]
    pre, post = np.nonzero(connection)  # row-major: ascending pre, then post
    rowptr = np.zeros(n + 1, np.int64)
    np.add.at(rowptr, pre + 1, 1)
    rowptr = np.cumsum(rowptr)
    return rowptr, post.astype(np.int32), s[pre, post].astype(np.float64)


def variant_a_weights(n: int, n_exc: int, connection: np.ndarray) -> np.ndarray:
    """s of Network(n, p, n_exc): +1 on exc rows, -1 on inh rows, masked (new_net.jl:87-91)."""
    s = np.zeros((n, n))
    s[:n_exc, :] = 1.0
    s[n_exc:, :] = -1.0
    return s * connection


def thalamic_noise(t: int, n: int, seed: int = 2, amp: float = 13.0) -> np.ndarray:
    """[t][n] recorded 13 .* (rand(N) .- 0.5) (new_net.jl:111), final fp64 values."""
    return amp * (_rng(seed).random((t, n)) - 0.5)


def fixed_fanout(n: int, n_exc: int, fanout: int, max_delay: int = 1, seed: int = 9,
                 w_exc: float = 1.0, w_inh: float = -1.0, n_instances: int = 1):
    """Old_net/DASTDP.jl:32 style: every neuron has `fanout` targets; exc -> anyone,
    inh -> exc only; exc delays U{1..max_delay}, inh delay 1.
    Returns (rowptr, post, w, delay) with global ids; instances are block-diagonal."""
    rng = _rng(seed)
    ntot = n * n_instances
    rowptr = np.arange(ntot + 1, dtype=np.int64) * fanout
    post = np.empty(ntot * fanout, np.int32)
    w = np.empty(ntot * fanout, np.float64)
    delay = np.ones(ntot * fanout, np.uint8)
    chunk = max(1, (1 << 24) // fanout)
    for b in range(n_instances):
        base = b * n
        for lo in range(0, n, chunk):
            hi = min(n, lo + chunk)
            r = np.arange(lo, hi)
            exc_rows = r < n_exc
            tgt = rng.integers(0, n, size=(hi - lo, fanout), dtype=np.int32)
            tgt_e = rng.integers(0, n_exc, size=(hi - lo, fanout), dtype=np.int32)
            tgt = np.where(exc_rows[:, None], tgt, tgt_e)
            sl = slice((base + lo) * fanout, (base + hi) * fanout)
            post[sl] = (tgt + base).reshape(-1)
            w[sl] = np.repeat(np.where(exc_rows, w_exc, w_inh), fanout)
            if max_delay > 1:
                d = rng.integers(1, max_delay + 1, size=(hi - lo, fanout), dtype=np.uint8)
                d = np.where(exc_rows[:, None], d, np.uint8(1))
                delay[sl] = d.reshape(-1)
    return rowptr, post, w, delay


def clustered(n: int = 10000, n_exc: int = 8000, fanout: int = 100, n_clusters: int = 20,
              p_in: float = 0.8, seed: int = 7):
    """C3 'MatrixCluster' synthetic net: n_clusters equal blocks, a fraction p_in of
    every neuron's targets inside its own cluster, the rest uniform."""
    rng = _rng(seed)
    csize = n // n_clusters
    rowptr = np.arange(n + 1, dtype=np.int64) * fanout
    own = (np.arange(n) // csize)[:, None]
    inside = rng.random((n, fanout)) < p_in
    local = rng.integers(0, csize, size=(n, fanout), dtype=np.int32) + own * csize
    anywhere = rng.integers(0, n, size=(n, fanout), dtype=np.int32)
    post = np.where(inside, local, anywhere).astype(np.int32).reshape(-1)
    w = np.repeat(np.where(np.arange(n) < n_exc, 1.0, -1.0), fanout)
    return rowptr, post, w, np.ones(n * fanout, np.uint8)
