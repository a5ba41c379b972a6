"""GPU tests of the ABI-3 additions, all through the C ABI: variant B's transposed storage, named parameters,
checkpoint / resume with axonal delays in flight, and the three ways a production-mode window can be scheduled
(persistent chain, captured graph, plain launches) giving the same spikes."""
import numpy as np
import pytest

from tests.helpers import run_oracle_sparse

pytestmark = pytest.mark.gpu


def _L():
    from rl_stdp_b200 import _lib as L
    return L


def test_variant_b_transposed_storage_fp64_exact():
    """new_net_col.jl:111-144 against the engine fed the [post, pre] matrices (snn_set_synapses_dense_col)."""
    from oracle import np_restate as nr
    from rl_stdp_b200 import synth
    from rl_stdp_b200.variants import NetworkCol
    N, Ne, T = 300, 240, 400
    conn = synth.dense_connection(N, 0.1, 11)
    noise = synth.thalamic_noise(T, N, 12)
    train = np.array([(k + 1) % 10 == 0 for k in range(T)], np.uint8)
    ref = nr.NetB(N, Ne, conn)
    net = NetworkCol(N, Ne, conn, dtype="f64", mode="exact")
    assert np.array_equal(net.s, ref.s)
    want = []
    for k in range(T):
        want.append(ref.step(noise[k], bool(train[k])))
        if k in (120, 260):
            ref.da += 0.5
    got, _ = net.run(noise, train, da_events=[(120, 0, 0.5), (260, 0, 0.5)])
    for k in range(T):
        assert np.array_equal(got[k], want[k]), k
    assert sum(len(x) for x in want) > 100
    exc_cols = np.zeros((N, N), bool); exc_cols[:, :Ne] = True
    mask = (ref.connection != 0) & exc_cols
    assert np.abs(net.s - ref.s).max() <= 1e-9
    assert np.abs(net.sd[mask] - ref.sd[mask]).max() <= 1e-9 * max(1.0, np.abs(ref.sd[mask]).max())
    assert np.abs(net.v - ref.v).max() <= 1e-9 * 65 and np.abs(net.STDP - ref.STDP).max() <= 1e-12
    assert abs(net.da - ref.da) <= 1e-12
    assert np.abs(ref.s - nr.NetB(N, Ne, conn).s).max() > 1e-3   # learning moved the weights


def _delay_net(seed=31, n=1500, ne=1200, fan=40, dmax=12):
    from rl_stdp_b200 import synth
    return (n, ne, dmax) + tuple(synth.fixed_fanout(n, ne, fan, dmax, seed=seed))


@pytest.mark.parametrize("dtype,mode", [("f64", "exact"), ("f32", "fast")])
def test_checkpoint_resume_with_delays_in_flight(dtype, mode):
    """Stop in the middle of a run (spikes of the last 12 ms still travelling, a learn step 3 ms away, dopamine
    decaying), checkpoint, continue; a SECOND handle restored from the blob must produce the same spikes and the
    same weights as the uninterrupted one.  Exact mode with learning: bit for bit (and equal to the oracle).
    Production mode: frozen weights, so that the atomic sums are order-independent and the comparison is exact too."""
    from oracle import py_oracle as po
    from rl_stdp_b200 import synth
    from rl_stdp_b200.network import SparseNetwork
    L = _L()
    n, ne, dmax, rowptr, post, w, delay = _delay_net()
    t1, t2 = 137, 150
    noise = synth.thalamic_noise(t1 + t2, n, seed=32)
    learning = mode == "exact"
    train = np.array([learning and (k + 1) % 10 == 0 for k in range(t1 + t2)], np.uint8)
    da = [(60, 0, 0.5), (131, 0, 0.4), (200, 0, 0.3)]
    mk = lambda: SparseNetwork(n, ne, rowptr, post, w, delay, dtype=dtype, mode=mode, max_delay=dmax)
    a = mk()
    sp1, _ = a.run(t1, noise[:t1], train[:t1], da_events=[e for e in da if e[0] < t1])
    blob = a.checkpoint()
    assert a.t == t1
    spa, _ = a.run(t2, noise[t1:], train[t1:], da_events=[(s - t1, i, x) for s, i, x in da if s >= t1])
    b = mk()
    b.restore(blob)
    assert b.t == t1
    spb, _ = b.run(t2, noise[t1:], train[t1:], da_events=[(s - t1, i, x) for s, i, x in da if s >= t1])
    assert sum(len(x) for x in spa) > 200
    for k in range(t2):
        assert np.array_equal(spa[k], spb[k]), k
    for f in (L.FIELD_V, L.FIELD_U, L.FIELD_W, L.FIELD_SD, L.FIELD_DA, L.FIELD_TRACE):
        assert np.array_equal(a.get(f), b.get(f)), f
    if learning:
        ora = po.Sparse(n, ne, rowptr, post, w, delay, max_delay=dmax)
        ref = run_oracle_sparse(ora, noise, train, da)
        for k in range(t2):
            assert np.array_equal(spb[k], ref[t1 + k]), k
        assert np.abs(b.get(L.FIELD_W) - ora.get(4)).max() <= 1e-9
    # a blob does not restore into a different network
    other = SparseNetwork(n, ne, rowptr, post, w, delay, dtype=dtype, mode=mode, max_delay=dmax + 1)
    with pytest.raises(L.SnnError):
        other.restore(blob)
    with pytest.raises(L.SnnError):
        b.restore(blob[: len(blob) // 2])


def test_named_parameters():
    from rl_stdp_b200 import synth
    from rl_stdp_b200.network import SparseNetwork
    L = _L()
    n, ne, dmax, rowptr, post, w, delay = _delay_net(seed=41, n=800, ne=640, fan=30, dmax=4)
    net = SparseNetwork(n, ne, rowptr, post, w, delay, dtype="f64", mode="fast", max_delay=dmax)
    assert net.get_param("apre") == 1.5 and net.get_param("persistent") == -1 and net.get_param("graph") == 1
    with pytest.raises(L.SnnError):
        net.set_param("no_such_parameter", 1)
    with pytest.raises(L.SnnError):
        net.set_param("learn_log_cap", 1000)      # shapes the synapse storage: only before the synapses are set
    with pytest.raises(L.SnnError):
        net.set_param("wmin", 10.0)               # wmin > wmax
    assert net.get_param("wmin") == 0.0
    # a model scalar changed between two windows takes effect: no LTD -> sd can only grow
    noise = synth.thalamic_noise(300, n, seed=42)
    train = np.array([(k + 1) % 10 == 0 for k in range(300)], np.uint8)
    net.run(150, noise[:150], train[:150])
    net.set_param("apre", 0.0)
    assert net.get_param("apre") == 0.0
    net.set(L.FIELD_SD, np.zeros(net.n_synapses))
    net.run(150, noise[150:], train[150:])
    sd = net.get(L.FIELD_SD)
    assert sd.min() >= 0.0 and sd.max() > 0.0
    # reserved[] of the config struct is no longer a side channel
    import ctypes as C
    from rl_stdp_b200.network import default_config
    cfg = default_config(n_neurons=100, n_per_instance=100, n_exc=80)
    cfg.reserved[4] = 2
    h = C.c_void_p()
    assert L.load().snn_create(C.byref(cfg), C.byref(h)) == -1


def test_three_schedulers_same_spikes():
    """Production mode, frozen weights (order-independent sums): the persistent chain, the captured graph and
    plain launches are three schedules of the same kernels' arithmetic and must emit identical spikes — and the
    ones of the oracle."""
    from oracle import py_oracle as po
    from rl_stdp_b200 import synth
    from rl_stdp_b200.network import SparseNetwork
    n, ne, dmax, rowptr, post, w, delay = _delay_net(seed=51, n=3000, ne=2400, fan=50, dmax=9)
    T = 240
    noise = synth.thalamic_noise(T, n, seed=52)
    ora = po.Sparse(n, ne, rowptr, post, w, delay, max_delay=dmax)
    ref = run_oracle_sparse(ora, noise, np.zeros(T, np.uint8), [])
    for kw, path in ((dict(persistent=True), 2), (dict(persistent=False), 1), (dict(graph=False), 0), (dict(), 2)):
        net = SparseNetwork(n, ne, rowptr, post, w, delay, dtype="f64", mode="fast", max_delay=dmax, **kw)
        got = []
        for lo in range(0, T, 80):
            sp, _ = net.run(80, noise[lo:lo + 80], None)
            got += sp
            assert net.get_param("window_path") == path
        for k in range(T):
            assert np.array_equal(got[k], ref[k]), (path, k)
    assert sum(len(x) for x in ref) > 500


def test_spike_record_is_ascending_for_sparse_steps_and_bursts():
    """The returned record is ascending inside every step: ordinary steps are sorted by one block per step in shared
    memory (`record_sort_kernel`), a step with more spikes than that holds (a forced burst of 9 000 of 12 000 neurons)
    by the library's segmented sort; both against the oracle."""
    from oracle import py_oracle as po
    from rl_stdp_b200 import synth
    from rl_stdp_b200.network import SparseNetwork
    n, ne, dmax, rowptr, post, w, delay = _delay_net(seed=61, n=12000, ne=9600, fan=20, dmax=3)
    T = 12
    noise = synth.thalamic_noise(T, n, seed=62)
    rng = np.random.default_rng(63)
    burst = np.sort(rng.choice(n, 9000, replace=False))
    force = [(5, int(j)) for j in burst] + [(8, 11999), (8, 7), (8, 4242)]
    force.sort(key=lambda e: e[0])
    ora = po.Sparse(n, ne, rowptr, post, w, delay, max_delay=dmax)
    ref = run_oracle_sparse(ora, noise, np.zeros(T, np.uint8), [], force=force)
    net = SparseNetwork(n, ne, rowptr, post, w, delay, dtype="f64", mode="exact", max_delay=dmax)
    got, _ = net.run(T, noise, None, force=force)
    assert len(got[5]) >= 9000 and len(got[8]) >= 3
    for k in range(T):
        assert np.all(np.diff(got[k]) > 0), k
        assert np.array_equal(got[k], ref[k]), k


def test_get_synapses_matches_get_array():
    """snn_get_synapses (the reference's `net.s[n1, syn]`, `net.sd[n1, syn]` reads) against the full arrays, before and
    after learning, with a repeated and a changed index set; error paths."""
    from rl_stdp_b200 import synth
    from rl_stdp_b200.network import SparseNetwork
    L = _L()
    n, ne, dmax, rowptr, post, w, delay = _delay_net(seed=71, n=1500, ne=1200, fan=40, dmax=5)
    net = SparseNetwork(n, ne, rowptr, post, w, delay, dtype="f64", mode="fast", max_delay=dmax)
    E = len(post)
    rng = np.random.default_rng(72)
    idx = rng.choice(E, 300, replace=False)
    gw, gsd = net.get_synapses(idx)
    assert np.array_equal(gw, np.asarray(w, np.float64)[idx]) and not gsd.any()
    noise = synth.thalamic_noise(120, n, seed=73)
    train = np.array([(k + 1) % 10 == 0 for k in range(120)], np.uint8)
    net.run(120, noise, train, da_events=[(30, 0, 0.5)], record=False)
    W, SD = net.get(L.FIELD_W), net.get(L.FIELD_SD)
    for ix in (idx, idx, idx[::-1].copy(), np.array([0, E - 1])):   # same set twice (cached positions), then others
        gw, gsd = net.get_synapses(ix)
        assert np.array_equal(gw, W[ix]) and np.array_equal(gsd, SD[ix])
    assert np.abs(SD).sum() > 0
    with pytest.raises(RuntimeError):
        net.get_synapses([E])
    with pytest.raises(RuntimeError):
        net.get_synapses(np.zeros(5000, np.int64))
