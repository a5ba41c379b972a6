"""The reference's other Izhikevich `Network` variants on the CUDA
engine, through the host mirrors of rl_stdp_b200.variants, against the literal numpy restatements
of the Julia sources (oracle/np_restate.py): fp64 exact mode, identical spikes, weights and traces."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_variant_c_col_sep_exact():
    """new_net_col_sep.jl: fixed -0.5 inhibition, STDP among excitatory neurons only, homeostasis for
    every spiker (inhibitory too), forced-spike and injected-current inputs (the latter with its own
    recorded noise draw, :133)."""
    from oracle import np_restate as nr
    from rl_stdp_b200 import synth
    from rl_stdp_b200.variants import NetworkColSep
    n, ne, t = 160, 128, 400
    conn = synth.dense_connection(n, 0.15, seed=61)
    noise = synth.thalamic_noise(t, n, seed=62, amp=40.0)
    noise_in = synth.thalamic_noise(t, n, seed=63)
    rng = np.random.Generator(np.random.Philox(key=64))
    ref = nr.NetColSep(n, ne, conn)
    net = NetworkColSep(n, ne, conn)
    total = 0
    for k in range(t):
        train = (k + 1) % 10 == 0
        kw_ref, kw = {}, {}
        if k % 7 == 3:
            idx = rng.choice(n, size=3, replace=False)
            kw_ref = kw = dict(input_spikes=idx)
        elif k % 7 == 5:
            idx, inten = rng.choice(ne, size=5, replace=False), rng.random(5)
            kw_ref = kw = dict(input_current=(idx, inten, noise_in[k]))
        a = ref.step(noise[k], train, **kw_ref)
        b = net.step(noise[k], train, **kw)
        assert np.array_equal(a, b), k
        total += len(a)
        if k == 120:
            ref.da += 0.5
            net.da = net.da + 0.5
    assert total > 400 and np.any(ref.ho[ne:] > 0)
    s_ref = np.concatenate([ref.s_exc, ref.s_inh], axis=1).T
    assert np.abs(net.s - s_ref).max() <= 1e-9 * 4
    ee = (conn != 0) & (np.arange(n)[:, None] < ne) & (np.arange(n)[None, :] < ne)
    sd_ref = np.zeros((n, n)); sd_ref[:ne, :ne] = ref.sd.T[:ne, :ne]
    assert np.abs(net.sd[ee] - sd_ref[ee]).max() <= 1e-9
    assert np.array_equal(net.v, ref.v) and np.array_equal(net.u, ref.u) and np.array_equal(net.ho, ref.ho)
    assert np.array_equal(net.STDP, ref.STDP) and net.da == ref.da
    assert np.abs(s_ref[ee] - 1.0).max() > 1e-3


@pytest.mark.parametrize("random_graph", [True, False])
def test_variant_d_net_res_exact(random_graph):
    """net_res.jl: parameters from the YAML dict, no noise, ee (+ei for the random graph) plasticity,
    homeostasis beyond the input layer, forced-spike and injected-current inputs."""
    from oracle import np_restate as nr
    from rl_stdp_b200.variants import NetworkRes
    rng = np.random.Generator(np.random.Philox(key=71))
    P = dict(c=-65.0, de=8.0, di=2.0, ae=0.02, ai=0.1, theta=1.5, ttheta=0.998, apost=1.2, apre=1.7, tda=0.99,
             wmax=3.0, imin=10.0, imax=40.0, wee=1.0, wei=2.0, wie=-1.5, i_per=0.3)
    if random_graph:
        n, ne, arch = 150, 120, [0]
        conn = (rng.random((n, n)) < 0.12).astype(np.int64)
        conn[ne:, ne:] = 0
        n_in = n
    else:
        P.update(wee=9.0, wmax=12.0, wei=25.0, wie=-4.0)
        arch = [16, 12, 8]
        ne, n = sum(arch), sum(arch) + arch[1]
        conn = np.zeros((n, n), np.int64)
        conn[0:16, 16:28] = 1; conn[16:28, 28:36] = 1
        mat_ie = (rng.random((12, 12)) < 0.3).astype(np.int64); np.fill_diagonal(mat_ie, 0)
        conn[ne:ne + 12, 16:28] = mat_ie
        conn[16:28, ne:ne + 12] = np.eye(12, dtype=np.int64)
        n_in = 16
    ref = nr.NetRes(n, ne, conn, arch, P)
    net = NetworkRes(n, ne, conn, arch, P)
    total = 0
    for k in range(300):
        train = (k + 1) % 5 == 0
        idx = rng.choice(n_in, size=4, replace=False)
        if k % 3 == 0:
            kw = dict(input_spikes=idx)
        else:
            kw = dict(input_current=(idx, rng.random(4)))
        a = ref.step(train, **kw)
        b = net.step(train, **kw)
        assert np.array_equal(a, b), k
        total += len(a)
        if k % 40 == 7:
            ref.da += 0.3
            net.da = net.da + 0.3
    assert total > 200
    s_ref = np.zeros((n, n)); s_ref[:ne, :ne] = ref.s_ee; s_ref[:ne, ne:] = ref.s_ei; s_ref[ne:, :ne] = ref.s_ie
    assert np.abs(net.s - s_ref).max() <= 1e-9 * P["wmax"]
    c = conn != 0
    sd = net.sd
    assert np.abs(sd[:ne, :ne][c[:ne, :ne]] - ref.sd_ee[c[:ne, :ne]]).max() <= 1e-9
    if random_graph:
        assert np.abs(sd[:ne, ne:][c[:ne, ne:]] - ref.sd_ei[c[:ne, ne:]]).max() <= 1e-9
        assert np.abs(ref.sd_ei).max() > 0
    assert np.array_equal(net.v, ref.v) and np.array_equal(net.ho, ref.ho) and net.da == ref.da
    assert np.abs(s_ref[:ne, :ne][c[:ne, :ne]] - P["wee"]).max() > 1e-3


@pytest.mark.parametrize("mode", ["exact", "fast"])
def test_variant_g_dastdp_exact(mode):
    """Old_net/DASTDP.jl + Julia_DA_STDPv2.jl: fan-out table, `>=` threshold, LTP from the previous
    ms's pre trace, LTP before LTD, descending current accumulation (snn_config.variant_g).  Rows of
    `post` are duplicate-free (see NetworkG's note on the reference's indexed-assignment artefact).
    The production mode (atomic scatter) is checked over the span where its +-1 weights are still
    exactly summable (before the first synweight_step changes them) and statistically afterwards."""
    from oracle import np_restate as nr
    from rl_stdp_b200 import synth
    from rl_stdp_b200.variants import NetworkG
    from tests.test_oracle_cpu import _g_post_table
    n, ne, m, t = 400, 320, 40, 500
    post = _g_post_table(n, ne, m, 81)
    noise = synth.thalamic_noise(t, n, seed=82, amp=26.0)
    ref = nr.NetG(post, ne)
    want = [ref.step(noise[k], k + 1) for k in range(t)]
    net = NetworkG(post, ne, mode=mode)
    got = net.run(noise, msec0=1)
    assert sum(len(w) for w in want) > 800
    if mode == "exact":
        for k in range(t):
            assert np.array_equal(got[k], want[k]), k
        assert np.abs(net.s - ref.s).max() <= 1e-9 * 4
        assert np.abs(net.sd[:ne] - ref.sd[:ne]).max() <= 1e-9
        assert np.array_equal(net.v, ref.v)
        assert np.abs(ref.s[:ne] - 1.0).max() > 1e-3
    else:
        for k in range(10):
            assert np.array_equal(got[k], want[k]), k
        ng, nw = sum(len(g) for g in got), sum(len(w) for w in want)
        assert abs(ng - nw) / nw < 0.05
        assert abs(net.s[:ne].mean() - ref.s[:ne].mean()) < 0.01
