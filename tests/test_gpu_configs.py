"""The named configurations of BASELINE.json / SURVEY 8(d), each against the CPU oracle (not the kernels against
themselves): C3 on the block-clustered topology, the C5 topology at N = 20 000 for 1000 ms, the full-size
1M / 100M network against the OpenMP oracle, and the fp32 statistical gates over 10 simulated seconds."""
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

from tests.helpers import rel_err, run_oracle_sparse

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _L():
    from rl_stdp_b200 import _lib as L
    return L


def test_c3_clustered_10k_fp64_exact_and_fp32_gates():
    """Config 3 (MatrixCluster): 10 000 neurons in 20 blocks of 500, 80 % of every neuron's 100 targets inside its
    own block, DA-STDP on.  fp64 exact mode: spikes identical to the oracle for 300 ms, weights / sd to 1e-9.
    fp32 production mode (persistent chain): population and per-cluster rates within the stated tolerances."""
    from oracle import py_oracle as po
    from rl_stdp_b200 import synth
    from rl_stdp_b200.network import SparseNetwork
    L = _L()
    n, ne, t = 10_000, 8_000, 300
    rowptr, post, w, delay = synth.clustered(n, ne, 100, 20, 0.8, seed=7)
    # the topology is what it says: >= 75 % of the synapses stay inside their block of 500
    pre = np.repeat(np.arange(n), np.diff(rowptr))
    assert np.mean(pre // 500 == post // 500) > 0.75
    noise = synth.thalamic_noise(t, n, seed=8)
    train = np.array([(k + 1) % 10 == 0 for k in range(t)], np.uint8)
    da = [(99, 0, 0.5), (220, 0, 0.5)]
    ora = po.Sparse(n, ne, rowptr, post, w, delay, max_delay=1)
    ref = run_oracle_sparse(ora, noise, train, da)
    net = SparseNetwork(n, ne, rowptr, post, w, delay, dtype="f64", mode="exact", max_delay=1)
    got, st = net.run(t, noise, train, da_events=da)
    for k in range(t):
        assert np.array_equal(got[k], ref[k]), k
    assert st.n_spikes == sum(len(r) for r in ref) > 2000
    plastic = pre < ne
    assert rel_err(net.get(L.FIELD_W), ora.get(4)) <= 1e-9
    assert rel_err(net.get(L.FIELD_SD)[plastic], ora.get(5)[plastic]) <= 1e-9
    assert np.abs(ora.get(4)[plastic] - 1.0).max() > 1e-3
    fast = SparseNetwork(n, ne, rowptr, post, w, delay, dtype="f32", mode="fast", max_delay=1, persistent=True)
    fs, _ = fast.run(t, noise, train, da_events=da)
    assert fast.get_param("window_path") == 2          # the persistent chain
    nr, ng = sum(len(r) for r in ref), sum(len(s) for s in fs)
    assert abs(ng - nr) / nr < 0.02
    cr = np.bincount(np.concatenate(ref) // 500, minlength=20).astype(float)
    cg = np.bincount(np.concatenate(fs) // 500, minlength=20).astype(float)
    assert np.abs(cg - cr).max() / cr.mean() < 0.10     # per-cluster activity (20 clusters x ~150 spikes)


def test_c5_topology_n20000_fp64_exact_1000ms():
    """'parity for this topology at N = 20 000 with replayed noise' — the C5 network (fan-out 100,
    excitatory delays U{1..20}, inhibitory delay 1, learn! every 10 ms, dopamine twice) at 1/50 of its size, fp64
    exact mode: every spike of the first 1000 simulated ms identical to the oracle, weights and sd to 1e-9."""
    from oracle import py_oracle as po
    from rl_stdp_b200 import synth
    from rl_stdp_b200.network import SparseNetwork
    L = _L()
    n, ne, fan, dmax, t = 20_000, 16_000, 100, 20, 1000
    rowptr, post, w, delay = synth.fixed_fanout(n, ne, fan, dmax, seed=9)
    noise = synth.thalamic_noise(t, n, seed=10)
    train = np.array([(k + 1) % 10 == 0 for k in range(t)], np.uint8)
    da = [(333, 0, 0.5), (777, 0, 0.5)]
    ora = po.Sparse(n, ne, rowptr, post, w, delay, max_delay=dmax)
    ref = run_oracle_sparse(ora, noise, train, da)
    net = SparseNetwork(n, ne, rowptr, post, w, delay, dtype="f64", mode="exact", max_delay=dmax)
    got = []
    for lo in range(0, t, 250):
        sp, _ = net.run(250, noise[lo:lo + 250], train[lo:lo + 250],
                        da_events=[(s - lo, i, a) for s, i, a in da if lo <= s < lo + 250])
        got += sp
    for k in range(t):
        assert np.array_equal(got[k], ref[k]), k
    assert sum(len(r) for r in ref) > 15_000
    plastic = np.repeat(np.arange(n), np.diff(rowptr)) < ne
    assert rel_err(net.get(L.FIELD_W), ora.get(4)) <= 1e-9
    assert rel_err(net.get(L.FIELD_SD)[plastic], ora.get(5)[plastic]) <= 1e-9
    assert rel_err(net.get(L.FIELD_TRACE), ora.get(2)) <= 1e-9
    assert rel_err(net.get(L.FIELD_V), ora.get(0)) <= 1e-9


def test_full_size_c5_frozen_equals_openmp_oracle():
    """BASELINE config 5 at FULL size against the oracle itself: 1M neurons / 100M synapses, delays 1-20, on-device
    Philox noise (the oracle replays the same stream), weights frozen (+-1 weights sum exactly in any order, so
    the atomic production kernels are order-independent), the fp64 build of the production path (persistent
    chain): the spike record of the first 60 ms must be bit-identical to the OpenMP oracle's, and v afterwards."""
    from rl_stdp_b200 import synth
    from rl_stdp_b200.network import SparseNetwork
    L = _L()
    n, ne, fan, dmax, t = 1_000_000, 800_000, 100, 20, 60
    with tempfile.TemporaryDirectory() as td:
        out = os.path.join(td, "ora.npz")
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "tools", "full_size_oracle.py"), out,
                            str(n), str(ne), str(fan), str(dmax), str(t), "9", "10", "frozen"],
                           capture_output=True, text=True, timeout=1500)
        assert r.returncode == 0, r.stderr[-2000:]
        z = np.load(out)
        ids, offs, v_ref = z["ids"], z["offs"], z["v"]
    rowptr, post, w, delay = synth.fixed_fanout(n, ne, fan, dmax, seed=9)
    net = SparseNetwork(n, ne, rowptr, post, w, delay, dtype="f64", mode="fast", noise="philox", max_delay=dmax,
                        noise_seed=10, persistent=True)
    got, st = net.run(t, None, None, spikes_cap=1 << 22)
    assert net.get_param("window_path") == 2
    assert offs[-1] > 20_000                      # the network is out of its silent start-up phase
    for k in range(t):
        assert np.array_equal(got[k], ids[offs[k]:offs[k + 1]]), k
    assert st.n_spikes == offs[-1]
    assert np.array_equal(net.get(L.FIELD_V), v_ref)


def test_fp32_production_statistics_10s():
    """SURVEY 8(d) gates over 10 simulated seconds (the round-1 gate ran 1.5 s): fp32 production mode (persistent
    chain, asynchronous learn pass) against the fp64 oracle on the same recorded noise — mean rate +-2 %, exc / inh
    rates +-5 %, ISI CV +-5 %, mean plastic weight +-1 %, weight KS < 0.02."""
    from oracle import py_oracle as po
    from scipy.stats import ks_2samp
    from rl_stdp_b200 import synth
    from rl_stdp_b200.network import SparseNetwork
    L = _L()
    n, ne, fan, dmax, t, W = 2000, 1600, 100, 10, 10_000, 500
    rowptr, post, w, delay = synth.fixed_fanout(n, ne, fan, dmax, seed=71)
    train = np.array([(k + 1) % 10 == 0 for k in range(t)], np.uint8)
    da = [(k, 0, 0.5) for k in range(999, t, 1000)]
    ora = po.Sparse(n, ne, rowptr, post, w, delay, max_delay=dmax)
    net = SparseNetwork(n, ne, rowptr, post, w, delay, dtype="f32", mode="fast", max_delay=dmax, learn_async="force",
                        persistent=True)
    ref, got = [], []
    for lo in range(0, t, W):          # the noise is drawn window by window: 10 s x 2000 neurons would be 160 MB at once
        noise = synth.thalamic_noise(W, n, seed=1000 + lo)
        ev = [(s - lo, i, a) for s, i, a in da if lo <= s < lo + W]
        ref += run_oracle_sparse(ora, noise, train[lo:lo + W], ev)
        sp, _ = net.run(W, noise, train[lo:lo + W], da_events=ev)
        got += sp
    assert net.get_param("window_path") == 2
    nr, ng = sum(len(r) for r in ref), sum(len(s) for s in got)
    assert nr > 20_000 and abs(ng - nr) / nr < 0.02
    re_, ge = sum(int((r < ne).sum()) for r in ref), sum(int((s < ne).sum()) for s in got)
    assert abs(ge - re_) / re_ < 0.05 and abs((ng - ge) - (nr - re_)) / max(1, nr - re_) < 0.05

    def isi_cv(trains):
        last = -np.ones(n, np.int64)
        isis = []
        for k, ids in enumerate(trains):
            seen = ids[last[ids] >= 0]
            isis.append(k - last[seen])
            last[ids] = k
        x = np.concatenate(isis).astype(float)
        return x.std() / x.mean()
    cr, cg = isi_cv(ref), isi_cv(got)
    assert abs(cg - cr) / cr < 0.05, (cr, cg)
    plastic = np.repeat(np.arange(n), np.diff(rowptr)) < ne
    wr, wg = ora.get(4)[plastic], net.get(L.FIELD_W)[plastic]
    assert abs(wg.mean() - wr.mean()) / wr.mean() < 0.01
    assert ks_2samp(np.round(wr, 4), np.round(wg, 4)).statistic < 0.02
    assert wr.std() > 0.01                           # ten seconds of DA-STDP did shape the weights
