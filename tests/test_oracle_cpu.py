"""CPU tests of the oracle itself: the C restatement, the independent numpy restatement and the
sparse generalisation must agree BIT FOR BIT (the reference has no golden vectors, so the two
restatements written from the cited Julia lines pin each other), and both must reproduce the
committed fixtures under tests/golden/."""
import os

import numpy as np
import pytest

from oracle import np_restate as nr
from oracle import py_oracle as po
from rl_stdp_b200 import synth
from tests.helpers import run_oracle_sparse

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _drive_a(N, Ne, T, seed):
    conn = synth.dense_connection(N, 0.1, seed)
    noise = synth.thalamic_noise(T, N, seed + 1)
    return conn, noise


def test_dense_c_equals_numpy_equals_sparse():
    N, Ne, T = 300, 240, 500
    conn, noise = _drive_a(N, Ne, T, 1)
    A = po.DenseA(N, Ne, conn)
    B = nr.NetA(N, Ne, conn)
    s0 = synth.variant_a_weights(N, Ne, conn)
    rowptr, post, w = synth.dense_to_csr(s0, conn)
    S = po.Sparse(N, Ne, rowptr, post, w)
    total = 0
    for t in range(T):
        train = (t + 1) % 10 == 0
        force = [5, 17] if t == 250 else None
        a = A.step(noise[t], train, force)
        b = B.step(noise[t], train, force)
        c = S.step(noise[t], train, force)
        assert np.array_equal(a, b) and np.array_equal(a, c), t
        total += len(a)
        if t in (120, 300):
            A.da += 0.5; B.da += 0.5; S.add_da(0, 0.5)
    assert total > 100
    assert np.array_equal(A.s, B.s) and np.array_equal(A.sd, B.sd)
    assert np.array_equal(A.v, B.v) and np.array_equal(A.u, B.u) and np.array_equal(A.STDP, B.STDP)
    assert A.da == B.da
    pre, pst = np.nonzero(conn)
    exc = pre < Ne
    assert np.array_equal(A.s[pre, pst], S.get(4))
    assert np.array_equal(A.sd[pre, pst][exc], S.get(5)[exc])
    assert np.array_equal(A.v, S.get(0)) and np.array_equal(A.STDP, S.get(2))
    assert A.da == S.get(6)[0]
    assert np.abs(A.sd).max() > 0 and np.abs(A.s - s0).max() > 0  # learning happened


def test_variant_b_restatement_equals_variant_a_and_sparse_oracle():
    """Variant B (new_net_col.jl:111-144) stores s / sd / connection transposed and walks the row of a spiker before
    its column (:127-130).  Its literal restatement must give, bit for bit, the transpose of variant A's matrices and
    the sparse oracle's values — the evidence (not just the argument) that the CSR engine covers B."""
    N, Ne, T = 260, 200, 400
    conn, noise = _drive_a(N, Ne, T, 5)
    A = nr.NetA(N, Ne, conn)
    B = nr.NetB(N, Ne, conn)
    s0 = synth.variant_a_weights(N, Ne, conn)
    rowptr, post, w = synth.dense_to_csr(s0, conn)
    S = po.Sparse(N, Ne, rowptr, post, w)
    assert np.array_equal(B.s, A.s.T) and np.array_equal(B.connection, A.connection.T)
    total = 0
    for t in range(T):
        train = (t + 1) % 10 == 0
        force = [3, 41] if t == 200 else None
        a = A.step(noise[t], train, force)
        b = B.step(noise[t], train, force)
        c = S.step(noise[t], train, force)
        assert np.array_equal(a, b) and np.array_equal(b, c), t
        total += len(b)
        if t in (90, 250):
            A.da += 0.5; B.da += 0.5; S.add_da(0, 0.5)
    assert total > 100
    assert np.array_equal(B.s, A.s.T) and np.array_equal(B.sd, A.sd.T)
    assert np.array_equal(B.v, A.v) and np.array_equal(B.u, A.u) and np.array_equal(B.STDP, A.STDP) and B.da == A.da
    pre, pst = np.nonzero(conn)
    exc = pre < Ne
    assert np.array_equal(B.s[pst, pre], S.get(4))
    assert np.array_equal(B.sd[pst, pre][exc], S.get(5)[exc])
    assert np.abs(B.sd).max() > 0 and np.abs(B.s - s0.T).max() > 0  # learning happened


def test_golden_variant_a():
    g = np.load(os.path.join(GOLD, "variant_a_n150_t300.npz"))
    N, Ne, T = int(g["N"]), int(g["Ne"]), int(g["T"])
    conn, noise = g["connection"], g["noise"]
    A = po.DenseA(N, Ne, conn)
    ids, offs = [], [0]
    for t in range(T):
        s = A.step(noise[t], (t + 1) % 10 == 0)
        ids.extend(s.tolist()); offs.append(len(ids))
        if t in g["da_steps"]:
            A.da += 0.5
    assert np.array_equal(np.array(ids, np.int32), g["spike_ids"])
    assert np.array_equal(np.array(offs, np.int64), g["spike_offsets"])
    assert np.array_equal(A.s, g["s_final"]) and np.array_equal(A.sd, g["sd_final"])
    assert np.array_equal(A.v, g["v_final"]) and A.da == float(g["da_final"])


def test_sparse_delay1_equals_delayD_with_shifted_semantics():
    """Sanity of the delay ring: with every delay = d the network is the delay-1 network whose
    arrivals (current + LTD) come d-1 calls later; with no recurrent spikes inside d calls the
    first spikes are identical."""
    N, Ne, T = 200, 160, 60
    rowptr, post, w, _ = synth.fixed_fanout(N, Ne, 30, 1, seed=3)
    noise = synth.thalamic_noise(T, N, seed=4)
    a = po.Sparse(N, Ne, rowptr, post, w, np.ones(len(post), np.uint8), max_delay=1)
    b = po.Sparse(N, Ne, rowptr, post, w, np.full(len(post), 4, np.uint8), max_delay=4)
    first = None
    for t in range(T):
        sa, sb = a.step(noise[t]), b.step(noise[t])
        if first is None and (len(sa) or len(sb)):
            first = t
            assert np.array_equal(sa, sb)
    assert first is not None


def test_philox_known_answer():
    """Philox4x32-10 known-answer vectors (Random123 kat_vectors): pins the shared noise generator."""
    out = np.zeros(4, np.uint32)
    L = po.lib()
    L.ora_philox4x32(0, 0, 0, 0, 0, 0, out.ctypes.data)
    assert [hex(x) for x in out] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    L.ora_philox4x32(0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, out.ctypes.data)
    assert [hex(x) for x in out] == ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]
    L.ora_philox4x32(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0, out.ctypes.data)
    assert [hex(x) for x in out] == ["0xd16cfe09", "0x94fdcceb", "0x5001e420", "0x24126ea1"]
    x = po.philox_noise(7, 3, 1000)
    assert x.min() >= -6.5 and x.max() <= 6.5 and abs(x.mean()) < 0.5


def test_run_helper_da_after_step():
    N, Ne = 100, 80
    rowptr, post, w, d = synth.fixed_fanout(N, Ne, 10, 2, seed=5)
    o = po.Sparse(N, Ne, rowptr, post, w, d, max_delay=2)
    run_oracle_sparse(o, synth.thalamic_noise(3, N, seed=6), [0, 0, 0], [(0, 0, 1.0)])
    assert o.get(6)[0] == 1.0 * 0.995 * 0.995


def test_layered_c_equals_numpy_lif_and_izh():
    """Variants F (LIF) and E (Izhikevich layered): C oracle == numpy restatement, bit for bit."""
    rng = np.random.Generator(np.random.Philox(key=77))
    arch = [6, 8, 4]
    param = dict(po.CARTPOLE_CFGLIF)
    param.update(theta=0.05, wei=0.3, wie=-0.2, apost=0.014, apre=0.019, ttrace=0.95, imin=0.9, imax=1.5, eta=0.5,
                 winh=-0.1)
    e0 = [np.clip(rng.standard_normal((arch[l], arch[l + 1])) * 0.3 + 0.8, 0, 1.1) for l in range(2)]
    for critic in (False, True):
        c = po.Layered(arch, "lif", **param)
        n = nr.NetArch(arch, param, e0)
        for l in range(2):
            c.e(l)[:, :] = e0[l]
        tot = 0
        for t in range(120):
            inten = rng.random(arch[0]) * 1.3
            if t % 7 == 0:
                c.da = c.da + 0.01; n.da += 0.01
            a = c.step_lif(np.arange(arch[0]), inten, True, critic)
            spikes = [(i, inten[i]) for i in range(arch[0])]
            b = n.step_LIF_critic(spikes, True) if critic else n.step_LIF(spikes, True)
            assert np.array_equal(a, b), (critic, t)
            tot += len(a)
        assert tot > 50
        assert np.array_equal(c.v, n.v) and np.array_equal(c.ho, n.ho) and np.array_equal(c.trace_e, n.trace_e)
        for l in range(2):
            assert np.array_equal(c.e(l), n.e[l]) and np.array_equal(c.de(l), n.de[l])
        assert c.da == n.da
    # Izhikevich layered (variant E)
    pe = dict(param)
    pe.update(vthresh=30.0, c=-65.0, theta=2.0, ttheta=0.9999, wei=3.0, wie=-2.0, wmax=20.0, imin=0.0, imax=20.0,
              apost=1.0, apre=1.5, ttrace=0.95, tda=0.995)
    e0 = [np.clip(rng.standard_normal((arch[l], arch[l + 1])) * 2 + 8, 0, 20.0) for l in range(2)]
    c = po.Layered(arch, "izh", **pe)
    n = nr.NetArch(arch, pe, e0)
    for l in range(2):
        c.e(l)[:, :] = e0[l]
    tot = 0
    for t in range(300):
        if t % 3 == 0:
            idx = rng.choice(arch[0], size=2, replace=False)
            inten = rng.random(2)
            a = c.step_izh(idx, inten, train=(t % 10 == 9))
            b = n.step([(int(i), float(x)) for i, x in zip(idx, inten)], train=(t % 10 == 9))
        elif t % 3 == 1:
            idx = rng.choice(arch[0], size=1)
            a = c.step_izh(idx, None, train=False)
            b = n.step([int(idx[0])], train=False)
        else:
            a = c.step_izh(None, None, train=False)
            b = n.step(None, train=False)
        if t % 50 == 0:
            c.da = c.da + 0.1; n.da += 0.1
        assert np.array_equal(a, b), t
        tot += len(a)
    assert tot > 30
    assert np.array_equal(c.v, n.v) and np.array_equal(c.u, n.u) and np.array_equal(c.ho, n.ho)
    for l in range(2):
        assert np.array_equal(c.e(l), n.e[l]) and np.array_equal(c.de(l), n.de[l])


def test_cartpole_restated_physics():
    """The restated CartPole-v1 step: polynomial sin/cos within 1 ulp of libm on the reachable angle
    range, a falling pole terminates at |theta| > 12 deg, and play_episode caps at n_steps."""
    import math
    L = po.lib()
    for x in np.linspace(-0.78, 0.78, 401):
        assert abs(L.ora_ksin(float(x)) - math.sin(x)) <= 1.2e-16
        assert abs(L.ora_kcos(float(x)) - math.cos(x)) <= 1.2e-16
    s = np.array([0.0, 0.0, 0.02, 0.0])
    steps, done = 0, False
    while not done and steps < 500:
        s, done = po.cartpole_step(s, 1)
        steps += 1
    assert done and steps < 200 and (abs(s[2]) > 12 * 2 * math.pi / 360 or abs(s[0]) > 2.4)
    s2, d2 = po.cartpole_step(np.array([0.0, 0.0, 0.0, 0.0]), 1)
    # one Euler step from rest with +10 N: xacc = temp - pml*thetaacc/total ; theta_dot < 0
    assert s2[1] > 0 and s2[3] < 0 and not d2


def _dense_sd_of(ora, rowptr, post, n):
    pre = np.repeat(np.arange(n), np.diff(rowptr))
    w, sd = np.zeros((n, n)), np.zeros((n, n))
    w[pre, post] = ora.get(4)
    sd[pre, post] = ora.get(5)
    return w, sd


def test_variant_c_restatement_equals_sparse_oracle():
    """new_net_col_sep.jl restated literally (numpy) == the sparse C oracle configured as variant C
    (plastic exc->exc only, fixed -0.5 inhibition, homeostasis for every spiker incl. inhibitory)."""
    from oracle import np_restate as nr, py_oracle as po
    from rl_stdp_b200 import synth
    n, ne, t = 120, 96, 400
    conn = synth.dense_connection(n, 0.15, seed=31)
    noise = synth.thalamic_noise(t, n, seed=32, amp=40.0)
    ref = nr.NetColSep(n, ne, conn)
    pre, post = np.nonzero(conn)
    rowptr = np.zeros(n + 1, np.int64); np.add.at(rowptr, pre + 1, 1); rowptr = np.cumsum(rowptr)
    w = np.where(pre < ne, 1.0, -0.5)
    ora = po.Sparse(n, ne, rowptr, post.astype(np.int32), w, plastic_ei=0, theta=2.0, ttheta=0.9999, ho_inh=1)
    total = 0
    for k in range(t):
        train = (k + 1) % 10 == 0
        a = ref.step(noise[k], train)
        b = ora.step(noise[k], train)
        assert np.array_equal(a, b), k
        total += len(a)
        if k == 150:
            ref.da += 0.5; ora.add_da(0, 0.5)
    assert total > 300 and np.any(ref.ho[ne:] > 0)  # inhibitory neurons did get homeostasis
    wd, sdd = _dense_sd_of(ora, rowptr, post, n)
    s_ref = np.concatenate([ref.s_exc, ref.s_inh], axis=1).T  # [pre, post]
    assert np.array_equal(wd, s_ref)
    ee = (conn != 0) & (np.arange(n)[:, None] < ne) & (np.arange(n)[None, :] < ne)
    assert np.array_equal(sdd[ee], ref.sd.T[:ne][ee[:ne]])
    assert np.array_equal(ora.get(0), ref.v) and np.array_equal(ora.get(3), ref.ho)


@pytest.mark.parametrize("random_graph", [True, False])
def test_variant_d_restatement_equals_sparse_oracle(random_graph):
    """net_res.jl restated literally == the sparse oracle configured from the YAML parameter dict."""
    from oracle import np_restate as nr, py_oracle as po
    rng = np.random.Generator(np.random.Philox(key=41))
    P = dict(c=-65.0, de=8.0, di=2.0, ae=0.02, ai=0.1, theta=1.5, ttheta=0.998, apost=1.2, apre=1.7, tda=0.99,
             wmax=3.0, imin=10.0, imax=40.0, wee=1.0, wei=2.0, wie=-1.5, i_per=0.3)
    if random_graph:
        n, ne, arch = 100, 80, [0]
        conn = (rng.random((n, n)) < 0.12).astype(np.int64)
        conn[ne:, ne:] = 0                                   # net_res.jl:42
    else:
        P.update(wee=9.0, wmax=12.0, wei=25.0, wie=-4.0)     # no thalamic noise in this variant: strong feed-forward drive
        arch = [12, 10, 8]
        ne, n = sum(arch), sum(arch) + arch[1]
        conn = np.zeros((n, n), np.int64)
        conn[0:12, 12:22] = 1; conn[12:22, 22:30] = 1        # :47-55 layer blocks
        mat_ie = (rng.random((10, 10)) < 0.3).astype(np.int64); np.fill_diagonal(mat_ie, 0)
        conn[ne:ne + 10, 12:22] = mat_ie                     # :68 recorded lateral inhibition
        conn[12:22, ne:ne + 10] = np.eye(10, dtype=np.int64)  # :69
    ref = nr.NetRes(n, ne, conn, arch, P)
    pre, post = np.nonzero(conn)
    rowptr = np.zeros(n + 1, np.int64); np.add.at(rowptr, pre + 1, 1); rowptr = np.cumsum(rowptr)
    w = np.where(pre < ne, np.where(post < ne, P["wee"], P["wei"]), P["wie"])
    ora = po.Sparse(n, ne, rowptr, post.astype(np.int32), w, v_reset=P["c"], a_exc=P["ae"], a_inh=P["ai"],
                    d_exc=P["de"], d_inh=P["di"], theta=P["theta"], ttheta=P["ttheta"], ho_from=arch[0],
                    apost=P["apost"], apre=P["apre"], da_decay=P["tda"], wmax=P["wmax"],
                    plastic_ei=1 if random_graph else 0)
    total = 0
    n_in = 12 if not random_graph else n
    for k in range(300):
        train = (k + 1) % 5 == 0
        idx = rng.choice(n_in, size=4, replace=False)
        if k % 3 == 0:   # forced spikes
            a = ref.step(train, input_spikes=idx)
            b = ora.step(None, train, idx.astype(np.int32))
        else:            # injected currents: two half-steps on v[idx] before spike! (:150-160)
            inten = rng.random(4)
            a = ref.step(train, input_current=(idx, inten))
            cur = (P["imax"] - P["imin"]) * inten + P["imin"]
            v = ora.get(0); u = ora.get(1)
            for _ in range(2):
                v[idx] = v[idx] + 0.5 * ((0.04 * v[idx] + 5) * v[idx] + 140 - u[idx] + cur)
            ora.set(0, v)
            b = ora.step(None, train)
        assert np.array_equal(a, b), k
        total += len(a)
        if k % 40 == 7:
            ref.da += 0.3; ora.add_da(0, 0.3)
    assert total > 200
    wd, sdd = _dense_sd_of(ora, rowptr, post, n)
    s_ref = np.zeros((n, n)); s_ref[:ne, :ne] = ref.s_ee; s_ref[:ne, ne:] = ref.s_ei; s_ref[ne:, :ne] = ref.s_ie
    assert np.array_equal(wd, s_ref)
    c = conn != 0
    assert np.array_equal(sdd[:ne, :ne][c[:ne, :ne]], ref.sd_ee[c[:ne, :ne]])
    if random_graph:
        assert np.array_equal(sdd[:ne, ne:][c[:ne, ne:]], ref.sd_ei[c[:ne, ne:]])
        assert np.abs(ref.sd_ei).max() > 0
    assert np.array_equal(ora.get(0), ref.v) and np.array_equal(ora.get(3), ref.ho)
    assert np.abs(s_ref[:ne, :ne][c[:ne, :ne]] - P["wee"]).max() > 1e-3  # learning moved the weights


def _g_post_table(n, ne, m, seed):
    """vcat(rand(1:N,Ne,M), rand(1:Ne,Ni,M)) (DASTDP.jl:32) without repeated targets inside a row."""
    rng = np.random.Generator(np.random.Philox(key=seed))
    post = np.zeros((n, m), np.int64)
    for i in range(n):
        post[i] = rng.choice(n if i < ne else ne, size=m, replace=False)
    return post


def test_variant_g_restatement_equals_sparse_oracle():
    """DASTDP.jl + Julia_DA_STDPv2.jl restated literally == the sparse oracle with variant_g = 1
    (LTP from the previous ms's pre trace, LTP before LTD, descending accumulation) and `>=`."""
    from oracle import np_restate as nr, py_oracle as po
    from rl_stdp_b200 import synth
    n, ne, m, t = 200, 160, 25, 600
    post = _g_post_table(n, ne, m, 51)
    noise = synth.thalamic_noise(t, n, seed=52, amp=30.0)
    ref = nr.NetG(post, ne)
    rowptr = np.arange(n + 1, dtype=np.int64) * m
    w = np.repeat(np.where(np.arange(n) < ne, 1.0, -1.0), m)
    ora = po.Sparse(n, ne, rowptr, post.reshape(-1).astype(np.int32), w, threshold_ge=1, variant_g=1)
    total = both = 0
    for k in range(t):
        msec = k + 1
        a = ref.step(noise[k], msec)
        b = ora.step(noise[k], msec % 10 == 0)
        assert np.array_equal(a, b), k
        total += len(a)
        if k == 200:
            ref.DA += 0.5; ora.add_da(0, 0.5)                # DA_inc after the step (:90)
    assert total > 500
    assert np.array_equal(ora.get(4).reshape(n, m), ref.s)
    assert np.array_equal(ora.get(5).reshape(n, m)[:ne], ref.sd[:ne])  # sd of inhibitory rows is dead state
    assert np.array_equal(ora.get(0), ref.v)
    assert np.abs(ref.s[:ne] - 1.0).max() > 1e-3


def test_reference_elites_known_answers():
    """The one place the reference pins this path with its OWN outputs: the sNES elites saved in
    Actors/cartpoleElites/sNESelites_cartpole_{1..17}.jld (170 individuals, 60 distinct genomes) carry the
    fitness play_episode (cartpole_fitness.jl:51-114) computed for them from gym's seed-0 reset state with
    the stock cartpole_cfglif.yml: 200.0 (200 decisions, no tie) for 59 of them and 199.1 (200 decisions,
    ONE tie, which costs 0.9) for one.  The oracle (LIF step, encoder, read-out, restated CartPole-v1) must
    reproduce every one — for the 199.1 genome with either tie action, so Julia's MersenneTwister draw does
    not matter.  Controls: vthresh 1.0 -> 1.05 leaves 2 of the 60, taulif 5 -> 5.2
    leaves 5, another start state 33-40; the 20 genomes of Actors/results_sNES_cartpole.jld, saved by an
    earlier state of the scripts, score 8-60 instead of their recorded 200."""
    from tests.helpers import gym_seed0_reset_state, reference_elites
    s0 = gym_seed0_reset_state()
    el = reference_elites()
    assert len(el) == 60 and sorted({f for _, _, f in el}) == [199.1, 200.0]
    assert len(reference_elites(unique=False)) == 170

    def play(m, w, tie, **par):
        n = po.Layered([8, 8, 8], **par)               # defaults = Julia/Actors/cartpole_cfglif.yml
        for l in range(2):
            n.e(l)[:, :] = w[l]
        return n.play_episode(m, s0, np.full(200, tie, np.int32))[:3]

    for m, w, fit in el:
        n_ties = 0 if fit == 200.0 else 1
        for tie in (0, 1):
            assert play(m, w, tie) == (fit, 200, n_ties)
    # sensitivity: with the threshold 5 % higher almost none of the answers survives
    assert sum(play(m, w, 0, vthresh=1.05)[0] == fit for m, w, fit in el) < 10


def test_julia_mersenne_twister_restatement():
    """rl_stdp_b200/julia_rng.py against the published head of `rand(MersenneTwister(0), 10)`."""
    from rl_stdp_b200.julia_rng import DSFMT, julia_mt_pick_bits
    g = DSFMT([0])
    assert [g.rand() for _ in range(10)] == [
        0.8236475079774124, 0.9103565379264364, 0.16456579813368521, 0.17732884646626457, 0.278880109331201,
        0.20347655804192266, 0.042301665932029664, 0.06826925550564478, 0.3618283907762174, 0.9732164043865108]
    assert julia_mt_pick_bits(0, 10) == [0, 0, 1, 1, 0, 0, 0, 1, 1, 0]
    g = DSFMT([0])
    for _ in range(1000):   # crosses two state regenerations
        assert 0.0 <= g.rand() < 1.0


def test_reference_mountcar_elites_known_answers():
    """Ten NON-saturated known answers: Actors/sNESelites_mountcar.jld stores, for ten genomes, the fitness
    play_episode (mountcar_fitness.jl:48-107) returned — minus one per decision until the flag is reached,
    minus two for a decision with a tie, the tie resolved by `rand(MersenneTwister(0), pool)`.  With gym's
    seed-0 reset position, Julia's tie stream (rl_stdp_b200/julia_rng.py) and a restated MountainCar-v0, the oracle's
    LIF step reproduces all ten, e.g. -129 = 121 decisions with 8 ties."""
    from tests.helpers import mountcar_elites, play_mountcar
    from rl_stdp_b200.julia_rng import julia_mt_pick_bits
    el = mountcar_elites()
    nets = []
    for _, w, _ in el:
        n = po.Layered([8, 8, 8])          # mountcar_cfglif.yml == the cartpole defaults
        for l in range(2):
            n.e(l)[:, :] = w[l]
        nets.append(n)
    idx = np.arange(8, dtype=np.int32)

    def window(intensity):
        out = np.zeros((len(nets), 8))
        for a, n in enumerate(nets):
            n.v[:] = n.params["c"]
            for ms in range(1, 41):
                for x in n.step_lif(idx, intensity[a], False):
                    if ms >= 2 and 16 <= x < 24:
                        out[a, x - 16] += 1
        return out

    tot, nd, ties = play_mountcar(window, [m for m, _, _ in el], julia_mt_pick_bits(0, 64))
    assert tot.tolist() == [f for _, _, f in el] == [-129.0, -129.0, -127.0, -127.0, -127.0, -125.0, -125.0,
                                                     -123.0, -121.0, -117.0]
    assert nd.tolist() == [121, 122, 118, 126, 111, 113, 121, 118, 117, 116]
    assert ties.tolist() == [8, 7, 9, 1, 18, 12, 4, 5, 4, 1]


def test_mountcar_oracle_episode_function():
    """ora_play_episode_mountcar (the C restatement the device is compared with) gives the same ten known
    answers as the Python-level replay above, which uses libm's cos; its reduced-argument cosine stays
    within one ulp of libm over the track."""
    import math
    from tests.helpers import _gym_np_random, mountcar_elites
    from rl_stdp_b200.julia_rng import julia_mt_pick_bits
    for x in np.linspace(-3.6, 1.8, 20001):
        c = po.cos_reduced(float(x))
        assert abs(c - math.cos(x)) <= np.spacing(abs(c)) * 1.0001, x
    p0 = _gym_np_random(0).uniform(low=-0.6, high=-0.4)
    assert abs(p0 - -0.5891279887498433) < 1e-15
    bits = julia_mt_pick_bits(0, 600)
    got = []
    for m, w, fit in mountcar_elites():
        n = po.Layered([8, 8, 8])
        for l in range(2):
            n.e(l)[:, :] = w[l]
        r, nd, nr, _ = n.play_episode_mountcar(m, [p0, 0.0], bits)
        assert r == fit
        got.append((nd, nr))
    assert got == list(zip([121, 122, 118, 126, 111, 113, 121, 118, 117, 116], [8, 7, 9, 1, 18, 12, 4, 5, 4, 1]))
