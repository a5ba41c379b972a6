"""Property tests (hypothesis; SURVEY section 4 vi).  The same four properties are checked twice: on the CPU
oracle (no GPU needed: guards the checker itself) and, marked gpu, on the CUDA library through the C ABI.

  P1  delay-1 CSR  ==  the literal dense restatement (any connectivity)
  P2  relabelling the neurons inside each class (exc / inh) relabels the spikes and nothing else
  P3  a batch of one instance  ==  the single network; a batch of B  ==  B single networks
  P4  a G-way neuron partition  ==  the unpartitioned network
"""
import numpy as np
import pytest
from hypothesis import HealthCheck, Phase, given, settings
from hypothesis import strategies as st

from oracle import np_restate as nr
from oracle import py_oracle as po
from rl_stdp_b200 import synth

CPU = settings(max_examples=12, deadline=None, suppress_health_check=list(HealthCheck))
# no shrinking on the device: a failing example is already small, and every retry costs a kernel launch (or a time-out)
GPU = settings(max_examples=5, deadline=None, suppress_health_check=list(HealthCheck),
               phases=(Phase.explicit, Phase.reuse, Phase.generate), database=None)


def _random_net(rng, n, ne, fan, dmax):
    rowptr = np.arange(n + 1, dtype=np.int64) * fan
    post = np.empty(n * fan, np.int32)
    for i in range(n):   # no duplicate targets inside a row, exc -> anyone, inh -> exc (Old_net/DASTDP.jl:32)
        post[i * fan:(i + 1) * fan] = np.sort(rng.choice(n if i < ne else ne, fan, replace=False))
    w = np.where(np.repeat(np.arange(n), fan) < ne, 1.0, -1.0) * rng.uniform(0.5, 1.5, n * fan)
    delay = np.where(np.repeat(np.arange(n), fan) < ne, rng.integers(1, dmax + 1, n * fan), 1).astype(np.uint8)
    return rowptr, post, w, delay


def _run_cpu(n, ne, rowptr, post, w, delay, dmax, noise, train, n_per=None):
    o = po.Sparse(n, ne, rowptr, post, w, delay, max_delay=dmax, n_per_instance=n_per)
    return [o.step(noise[k], bool(train[k])) for k in range(len(noise))], o


@CPU
@given(seed=st.integers(0, 10_000), p=st.floats(0.03, 0.3))
def test_p1_delay1_csr_equals_dense_oracle(seed, p):
    n, ne, T = 60, 48, 60
    conn = synth.dense_connection(n, p, seed)
    noise = synth.thalamic_noise(T, n, seed + 1, amp=30.0)
    a = nr.NetA(n, ne, conn)
    s0 = synth.variant_a_weights(n, ne, conn)
    rowptr, post, w = synth.dense_to_csr(s0, conn)
    s = po.Sparse(n, ne, rowptr, post, w)
    for t in range(T):
        tr = (t + 1) % 5 == 0
        assert np.array_equal(a.step(noise[t], tr), s.step(noise[t], tr)), t
    pre, pst = np.nonzero(conn)
    assert np.array_equal(a.s[pre, pst], s.get(4))


def _permute(n, ne, rowptr, post, w, delay, perm):
    """the same network with neuron i renamed perm[i] (perm maps exc to exc, inh to inh)"""
    fan = int(rowptr[1])
    inv = np.argsort(perm)
    rp = rowptr.copy()
    po_, w_, d_ = np.empty_like(post), np.empty_like(w), np.empty_like(delay)
    for new in range(n):
        old = inv[new]
        sl_o, sl_n = slice(old * fan, (old + 1) * fan), slice(new * fan, (new + 1) * fan)
        order = np.argsort(perm[post[sl_o]], kind="stable")
        po_[sl_n] = perm[post[sl_o]][order]; w_[sl_n] = w[sl_o][order]; d_[sl_n] = delay[sl_o][order]
    return rp, po_.astype(np.int32), w_, d_


@CPU
@given(seed=st.integers(0, 10_000))
def test_p2_relabelling_inside_a_class_cpu(seed):
    rng = np.random.default_rng(seed)
    n, ne, fan, dmax, T = 50, 40, 8, 4, 50
    rowptr, post, w, delay = _random_net(rng, n, ne, fan, dmax)
    # frozen weights that sum exactly in any order (relabelling changes the order in which I accumulates)
    w = np.where(np.repeat(np.arange(n), fan) < ne, 1.0, -1.0)
    perm = np.concatenate([rng.permutation(ne), ne + rng.permutation(n - ne)])
    noise = synth.thalamic_noise(T, n, seed + 3, amp=30.0)
    noise_p = np.empty_like(noise); noise_p[:, perm] = noise
    a, _ = _run_cpu(n, ne, rowptr, post, w, delay, dmax, noise, np.zeros(T))
    b, _ = _run_cpu(n, ne, *_permute(n, ne, rowptr, post, w, delay, perm), dmax, noise_p, np.zeros(T))
    for t in range(T):
        assert np.array_equal(np.sort(perm[a[t]]), b[t]), t


@CPU
@given(seed=st.integers(0, 10_000), B=st.integers(1, 3))
def test_p3_batch_equals_singles_cpu(seed, B):
    rng = np.random.default_rng(seed)
    n, ne, fan, dmax, T = 40, 32, 6, 3, 40
    nets = [_random_net(rng, n, ne, fan, dmax) for _ in range(B)]
    noise = synth.thalamic_noise(T, n * B, seed + 5, amp=30.0)
    train = np.array([(t + 1) % 5 == 0 for t in range(T)])
    rowptr = np.arange(n * B + 1, dtype=np.int64) * fan
    post = np.concatenate([p + b * n for b, (_, p, _, _) in enumerate(nets)]).astype(np.int32)
    w = np.concatenate([x[2] for x in nets]); delay = np.concatenate([x[3] for x in nets])
    big, ob = _run_cpu(n * B, ne, rowptr, post, w, delay, dmax, noise, train, n_per=n)
    for b, (rp, p, ww, dd) in enumerate(nets):
        one, oo = _run_cpu(n, ne, rp, p, ww, dd, dmax, noise[:, b * n:(b + 1) * n], train)
        for t in range(T):
            mine = big[t][(big[t] >= b * n) & (big[t] < (b + 1) * n)] - b * n
            assert np.array_equal(mine, one[t]), (b, t)
        assert np.array_equal(ob.get(4)[b * n * fan:(b + 1) * n * fan], oo.get(4))


# ----------------------------------------------------------------------------------------- on the device
def _run_gpu(n, ne, rowptr, post, w, delay, dmax, noise, train, **kw):
    from rl_stdp_b200.network import SparseNetwork
    net = SparseNetwork(n, ne, rowptr, post, w, delay, dtype="f64", max_delay=dmax, **kw)
    sp, _ = net.run(len(noise), noise, np.asarray(train, np.uint8))
    return sp, net


@pytest.mark.gpu
@GPU
@given(seed=st.integers(0, 10_000), p=st.floats(0.03, 0.3), mode=st.sampled_from(["exact", "fast"]))
def test_p1_delay1_csr_equals_dense_gpu(seed, p, mode):
    n, ne, T = 96, 72, 80
    conn = synth.dense_connection(n, p, seed)
    noise = synth.thalamic_noise(T, n, seed + 1, amp=30.0)
    train = np.array([mode == "exact" and (t + 1) % 5 == 0 for t in range(T)])
    a = nr.NetA(n, ne, conn)
    ref = [a.step(noise[t], bool(train[t])) for t in range(T)]
    s0 = synth.variant_a_weights(n, ne, conn)
    rowptr, post, w = synth.dense_to_csr(s0, conn)
    got, net = _run_gpu(n, ne, rowptr, post, w, None, 1, noise, train, mode=mode)
    for t in range(T):
        assert np.array_equal(got[t], ref[t]), t
    from rl_stdp_b200 import _lib as L
    pre, pst = np.nonzero(conn)
    assert np.abs(net.get(L.FIELD_W) - a.s[pre, pst]).max() <= 1e-9


@pytest.mark.gpu
@GPU
@given(seed=st.integers(0, 10_000), mode=st.sampled_from(["exact", "fast"]))
def test_p2_relabelling_inside_a_class_gpu(seed, mode):
    rng = np.random.default_rng(seed)
    n, ne, fan, dmax, T = 128, 96, 10, 5, 60
    rowptr, post, w, delay = _random_net(rng, n, ne, fan, dmax)
    w = np.where(np.repeat(np.arange(n), fan) < ne, 1.0, -1.0)
    perm = np.concatenate([rng.permutation(ne), ne + rng.permutation(n - ne)])
    noise = synth.thalamic_noise(T, n, seed + 3, amp=30.0)
    noise_p = np.empty_like(noise); noise_p[:, perm] = noise
    a, _ = _run_gpu(n, ne, rowptr, post, w, delay, dmax, noise, np.zeros(T), mode=mode)
    b, _ = _run_gpu(n, ne, *_permute(n, ne, rowptr, post, w, delay, perm), dmax, noise_p, np.zeros(T), mode=mode)
    for t in range(T):
        assert np.array_equal(np.sort(perm[a[t]]), b[t]), t


@pytest.mark.gpu
@GPU
@given(seed=st.integers(0, 10_000), B=st.integers(1, 4))
def test_p3_batch_equals_singles_gpu(seed, B):
    from rl_stdp_b200 import _lib as L
    rng = np.random.default_rng(seed)
    n, ne, fan, dmax, T = 128, 96, 8, 4, 50
    nets = [_random_net(rng, n, ne, fan, dmax) for _ in range(B)]
    noise = synth.thalamic_noise(T, n * B, seed + 5, amp=30.0)
    train = np.array([(t + 1) % 5 == 0 for t in range(T)])
    rowptr = np.arange(n * B + 1, dtype=np.int64) * fan
    post = np.concatenate([p + b * n for b, (_, p, _, _) in enumerate(nets)]).astype(np.int32)
    w = np.concatenate([x[2] for x in nets]); delay = np.concatenate([x[3] for x in nets])
    big, nb = _run_gpu(n, ne, rowptr, post, w, delay, dmax, noise, train, mode="exact", n_instances=B)
    wb = nb.get(L.FIELD_W)
    for b, (rp, p, ww, dd) in enumerate(nets):
        one, no = _run_gpu(n, ne, rp, p, ww, dd, dmax, noise[:, b * n:(b + 1) * n], train, mode="exact")
        for t in range(T):
            mine = big[t][(big[t] >= b * n) & (big[t] < (b + 1) * n)] - b * n
            assert np.array_equal(mine, one[t]), (b, t)
        assert np.array_equal(wb[b * n * fan:(b + 1) * n * fan], no.get(L.FIELD_W))


@pytest.mark.gpu
@GPU
@given(seed=st.integers(0, 10_000), world=st.sampled_from([2, 3, 4]), mode=st.sampled_from(["exact", "chain"]))
def test_p4_partition_equals_single_gpu(seed, world, mode):
    """G handles of one process on one GPU, peer-to-peer exchange (plain device pointers) — the same kernels the
    multi-GPU runs use; production mode (persistent chain with its remote role) with frozen weights, exact mode with
    learning.  (The captured-graph scheduler is not run this way: a rank that instantiates its graph while the
    kernels of its peers in the SAME process spin for it stalls them; one process per rank — the deployment, and
    test_partition_p2p_ipc_two_processes_one_gpu — has no such coupling.)"""
    from rl_stdp_b200.partition import PartitionedRank, connect_local, plan, run_local_ranks
    kw = dict(persistent=True) if mode == "chain" else {}   # "chain": production mode on the persistent chain (remote role)
    mode = "fast" if mode == "chain" else mode
    rng = np.random.default_rng(seed)
    n, ne, fan, dmax, T = 128 * world, 96 * world, 12, 5, 60
    rowptr, post, w, delay = _random_net(rng, n, ne, fan, dmax)
    if mode == "fast":
        w = np.where(np.repeat(np.arange(n), fan) < ne, 1.0, -1.0)
    noise = synth.thalamic_noise(T, n, seed + 7, amp=30.0)
    train = np.array([mode == "exact" and (t + 1) % 5 == 0 for t in range(T)], np.uint8)
    one, _ = _run_gpu(n, ne, rowptr, post, w, delay, dmax, noise, train, mode=mode)
    nets = [PartitionedRank(n, ne, rowptr, post, w, delay, lo, hi, None, dtype="f64", mode=mode, max_delay=dmax, **kw)
            for lo, hi in plan(n, world)]
    connect_local(nets)
    outs = run_local_ranks(nets, T, noise, train)
    for sp, _ in outs:
        for t in range(T):
            assert np.array_equal(sp[t], one[t]), t
