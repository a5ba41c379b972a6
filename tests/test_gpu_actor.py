"""GPU parity of the batched LIF actor kernels (variant F, configs 2 and 4) against the oracle
(oracle/lif_oracle.c), through the C ABI.  fp64, bit-exact: spike masks, counts, weights, traces,
episode returns and final CartPole states."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _mk(n_agents, arch, seed, **param):
    from oracle import py_oracle as po
    from rl_stdp_b200.lif import ActorPopulation
    rng = np.random.Generator(np.random.Philox(key=seed))
    pop = ActorPopulation(arch, n_agents, **param)
    oparam = {("de" if k == "de_" else k): v for k, v in param.items() if k != "da_inc"}
    oras = [po.Layered(arch, "lif", **oparam) for _ in range(n_agents)]
    # Weights(arch, param) :21 — clamp.(randn .* std .+ m, 0, wmax), recorded
    wmax = param.get("wmax", 1.1)
    for l in range(len(arch) - 1):
        e = np.clip(rng.standard_normal((n_agents, arch[l], arch[l + 1])) * 0.3 + 0.8, 0, wmax)
        pop.set_weights(l, e)
        for a, o in enumerate(oras):
            o.e(l)[:, :] = e[a]
    return pop, oras, rng


@pytest.mark.parametrize("train,critic", [(False, False), (True, False), (True, True)])
def test_decision_window_bitexact(train, critic):
    from rl_stdp_b200 import _lib as L
    arch, A, T = [8, 8, 8], 37, 40
    param = dict(theta=0.05, wei=0.3, wie=-0.2, apost=0.014, apre=0.019, ttrace=0.95, imin=0.9, eta=0.5, winh=-0.1)
    pop, oras, rng = _mk(A, arch, 5, **param)
    idx = np.arange(arch[0], dtype=np.int32)
    for rep in range(3):  # three consecutive decisions: ho / trace / de / da carry over, v is reset
        inten = rng.random((A, arch[0])) * 1.2
        da0 = rng.random(A) * 0.02 - 0.005
        pop.set(L.FIELD_DA, pop.get(L.FIELD_DA)[:, 0] + da0)
        counts, masks, _ = pop.decision_window(inten, T, train=train, critic=critic, reset_v=True, spikes=True)
        for a, o in enumerate(oras):
            o.da = o.da + da0[a]
            o.v[:] = o.params["c"]
            cnt = np.zeros(arch[-1], np.int32)
            for msec in range(1, T + 1):
                sp = o.step_lif(idx, inten[a], train, critic)
                m = 0
                for j in sp:
                    m |= 1 << int(j)
                assert (int(masks[a, msec - 1]) & 0xFFFFFFFF) == m, (rep, a, msec)
                if msec >= len(arch) - 1:
                    for j in sp:
                        if j >= sum(arch[:-1]) and j < sum(arch):
                            cnt[j - sum(arch[:-1])] += 1
            assert np.array_equal(counts[a], cnt)
    v, ho, tr, da = pop.get(L.FIELD_V), pop.get(L.FIELD_HO), pop.get(L.FIELD_TRACE), pop.get(L.FIELD_DA)
    tot = 0
    for a, o in enumerate(oras):
        assert np.array_equal(v[a], o.v) and np.array_equal(ho[a], o.ho) and np.array_equal(tr[a], o.trace_e)
        assert da[a, 0] == o.da
        for l in range(len(arch) - 1):
            assert np.array_equal(pop.weights(l)[a], o.e(l)), (a, l)
            assert np.array_equal(pop.weights(l, L.FIELD_SD)[a], o.de(l)), (a, l)
            tot += np.abs(o.de(l)).sum()
    assert tot > 0


@pytest.mark.parametrize("critic", [False, True])
def test_learn_on_its_own_bitexact(critic):
    """snn_actor_learn: learn!(net) / learn_critic!(net) called outside step_LIF!, as critic_sim.jl:107 does after a
    window run with train = false — eligibility accumulates over the window, one weight update at its end."""
    from rl_stdp_b200 import _lib as L
    arch, A, T = [8, 8, 8], 9, 40
    param = dict(theta=0.05, wei=0.3, wie=-0.2, apost=0.014, apre=0.019, ttrace=0.95, imin=0.9, eta=0.5, winh=-0.1)
    pop, oras, rng = _mk(A, arch, 15, **param)
    idx = np.arange(arch[0], dtype=np.int32)
    inten = rng.random((A, arch[0])) * 1.2
    da0 = rng.random(A) * 0.05
    pop.set(L.FIELD_DA, da0)
    pop.decision_window(inten, T, train=False, reset_v=True)
    pop.learn(critic=critic)
    moved = 0.0
    for a, o in enumerate(oras):
        o.da = da0[a]
        o.v[:] = o.params["c"]
        w0 = [o.e(l).copy() for l in range(len(arch) - 1)]
        for _ in range(T):
            o.step_lif(idx, inten[a], False, False)
        o.learn(critic)
        for l in range(len(arch) - 1):
            assert np.array_equal(pop.weights(l)[a], o.e(l)), (a, l)
            moved += np.abs(o.e(l) - w0[l]).sum()
    assert moved > 0


@pytest.mark.parametrize("arch", [[8, 8, 8, 8, 8], [12, 10, 6], [16, 16, 16]])
def test_larger_actors_two_neurons_per_lane(arch):
    """Nets of 33..64 neurons (e.g. the [8,8,8,8,8] actors of CartPole/LIF_base): two neurons per lane.
    Decision windows with learning and whole episodes, bit-exact against the oracle."""
    from rl_stdp_b200 import _lib as L
    from rl_stdp_b200.lif import borne
    A, T = 19, 40
    param = dict(theta=0.05, wei=0.3, wie=-0.2, apost=0.014, apre=0.019, ttrace=0.95, imin=0.9)
    pop, oras, rng = _mk(A, arch, 21, **param)
    idx = np.arange(arch[0], dtype=np.int32)
    n_out0 = sum(arch[:-1])
    for rep in range(2):
        inten = rng.random((A, arch[0])) * 1.5
        da0 = rng.random(A) * 0.02 - 0.005
        pop.set(L.FIELD_DA, pop.get(L.FIELD_DA)[:, 0] + da0)
        counts, _, _ = pop.decision_window(inten, T, train=True, reset_v=True)
        for a, o in enumerate(oras):
            o.da = o.da + da0[a]
            o.v[:] = o.params["c"]
            cnt = np.zeros(arch[-1], np.int32)
            for msec in range(1, T + 1):
                sp = o.step_lif(idx, inten[a], True, False)
                if msec >= len(arch) - 1:
                    for j in sp:
                        if n_out0 <= j < sum(arch):
                            cnt[j - n_out0] += 1
            assert np.array_equal(counts[a], cnt), (rep, a)
    v, ho, tr, da = pop.get(L.FIELD_V), pop.get(L.FIELD_HO), pop.get(L.FIELD_TRACE), pop.get(L.FIELD_DA)
    moved = 0
    for a, o in enumerate(oras):
        assert np.array_equal(v[a], o.v) and np.array_equal(ho[a], o.ho) and np.array_equal(tr[a], o.trace_e)
        assert da[a, 0] == o.da
        for l in range(len(arch) - 1):
            assert np.array_equal(pop.weights(l)[a], o.e(l)), (a, l)
            assert np.array_equal(pop.weights(l, L.FIELD_SD)[a], o.de(l)), (a, l)
            moved += np.abs(o.de(l)).sum()
    assert moved > 0 and counts.sum() > 0
    # whole episodes
    n_steps = 120
    matalpha = borne(rng.random((A, arch[0], 4)) * 20 - 10, -1.0, 1.0)
    state0 = rng.random((A, 4)) * 0.1 - 0.05
    ties = rng.integers(0, 2, size=(A, n_steps), dtype=np.int32)
    out = pop.play_episodes(matalpha, state0, ties, n_steps=n_steps)
    for a, o in enumerate(oras):
        r, nd, nr, so = o.play_episode(matalpha[a], state0[a], ties[a], n_steps, train=False)
        assert out["rewards"][a] == r and out["n_decisions"][a] == nd and out["n_rand"][a] == nr
        assert np.array_equal(out["state"][a], so)


def test_episodes_bitexact_and_ragged():
    """play_episode for 64 agents (different weights, M_alpha, reset states, tie streams): returns,
    decision counts, tie counts and final env states identical to the oracle; episodes end at
    different lengths (ragged), some hit the 200-decision cap."""
    from oracle import py_oracle as po
    from rl_stdp_b200.lif import borne
    arch, A, n_steps = [8, 8, 8], 64, 200
    pop, oras, rng = _mk(A, arch, 9)
    matalpha = borne(rng.random((A, arch[0], 4)) * 20 - 10, -1.0, 1.0)
    state0 = rng.random((A, 4)) * 0.1 - 0.05  # env.reset(): U(-0.05, 0.05)^4
    ties = rng.integers(0, 2, size=(A, n_steps), dtype=np.int32)
    out = pop.play_episodes(matalpha, state0, ties, n_steps=n_steps)
    lens = set()
    for a, o in enumerate(oras):
        r, nd, nr, so = o.play_episode(matalpha[a], state0[a], ties[a], n_steps, train=False)
        assert out["rewards"][a] == r, a
        assert out["n_decisions"][a] == nd and out["n_rand"][a] == nr
        assert np.array_equal(out["state"][a], so)
        lens.add(nd)
    assert len(lens) > 3


def test_teacher_style_learning_episode():
    """train=true episodes with dopamine +-da_inc per decision (teacher_sim.jl:61-80,132-133)."""
    from rl_stdp_b200 import _lib as L
    arch, A, n_steps = [8, 8, 8], 16, 60
    param = dict(apost=0.014, apre=0.019, ttrace=0.95, imin=0.5, da_inc=0.0001)
    pop, oras, rng = _mk(A, arch, 13, **param)
    matalpha = rng.random((A, arch[0], 4)) * 2 - 1
    state0 = rng.random((A, 4)) * 0.1 - 0.05
    ties = rng.integers(0, 2, size=(A, n_steps), dtype=np.int32)
    teach = rng.integers(0, 2, size=(A, n_steps), dtype=np.int32)
    out = pop.play_episodes(matalpha, state0, ties, n_steps=n_steps, train=True, teach_actions=teach)
    changed = 0
    for a, o in enumerate(oras):
        e_before = [o.e(l).copy() for l in range(2)]
        r, nd, nr, so = o.play_episode(matalpha[a], state0[a], ties[a], n_steps, train=True, teach_actions=teach[a],
                                       da_inc=param["da_inc"])
        assert out["rewards"][a] == r and out["n_decisions"][a] == nd
        for l in range(2):
            assert np.array_equal(pop.weights(l)[a], o.e(l))
            changed += int(np.any(o.e(l) != e_before[l]))
        assert pop.get(L.FIELD_DA)[a, 0] == o.da
    assert changed > 0


def test_actor_config_errors():
    from rl_stdp_b200 import _lib as L
    from rl_stdp_b200.lif import ActorPopulation
    with pytest.raises(L.SnnError):
        ActorPopulation([24, 24, 24], 4)  # 72 exc + 24 inh > 64 neurons (two per lane)
    with pytest.raises(L.SnnError):
        ActorPopulation([8, 40, 8], 4)    # a layer wider than 32
    with pytest.raises(L.SnnError):
        ActorPopulation([8, 8, 8], 0)


def test_reference_elites_balance_the_pole():
    """End to end against the reference's OWN results: the ten sNES elites it saved in
    Actors/cartpoleElites/sNESelites_cartpole_1.jld (all with fitness 200 = the episode cap; loaded by
    cartpole_tryout.jl:15) are decoded as fitness() does (cartpole_fitness.jl:158-176) and played on the
    device — LIF actor kernels + restated CartPole-v1 — from 30 random reset states each, like
    test_random (cartpole_tryout.jl:53-64).  A controller evolved on the reference implementation only
    keeps the pole up here if the network semantics, the input encoding, the read-out and the physics
    agree with it."""
    import os
    from rl_stdp_b200 import jld
    from rl_stdp_b200.lif import ActorPopulation, borne, set_weight
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    elites = jld.load(os.path.join(root, "tests", "golden", "sNESelites_cartpole_1.jld"), "elites")
    arch, reps = [8, 8, 8], 30
    A = len(elites) * reps
    rng = np.random.Generator(np.random.Philox(key=77))
    pop = ActorPopulation(arch, A)  # defaults = Julia/Actors/cartpole_cfglif.yml
    e0, e1, ma = [], [], []
    for el in elites:
        g = el["genes"]
        m = borne(g[:arch[0] * 4], -1.0, 1.0).reshape((4, arch[0])).T   # reshape(genes_alpha, (arch[1], 4))
        w = set_weight(arch, g[arch[0] * 4:])
        for _ in range(reps):
            ma.append(m); e0.append(w[0]); e1.append(w[1])
    pop.set_weights(0, np.array(e0)); pop.set_weights(1, np.array(e1))
    state0 = rng.random((A, 4)) * 0.1 - 0.05
    ties = rng.integers(0, 2, size=(A, 200), dtype=np.int32)
    out = pop.play_episodes(np.array(ma), state0, ties, n_steps=200)
    rew = out["rewards"].reshape(len(elites), reps)
    print("mean reward per elite:", np.round(rew.mean(axis=1), 1), " overall", rew.mean())
    # measured: per-elite means 200, 200, 139, 163, 115, 98, 187, 89, 200, 130 (overall 152); the elites were
    # selected on ONE seeded start (fitness 200 there), so some are less robust to random starts
    assert rew.mean() > 120.0, rew.mean(axis=1)
    assert (rew.mean(axis=1) >= 195.0).sum() >= 2
    assert (out["n_decisions"] == 200).mean() > 0.4
    # contrast: random genomes drawn like the optimiser's initial population do not balance the pole
    rnd = ActorPopulation(arch, A)
    g = rng.standard_normal((A, 160)) * 3.0
    rnd.set_weights(0, np.array([set_weight(arch, x[32:])[0] for x in g]))
    rnd.set_weights(1, np.array([set_weight(arch, x[32:])[1] for x in g]))
    ma_r = np.array([borne(x[:32], -1.0, 1.0).reshape((4, 8)).T for x in g])
    out_r = rnd.play_episodes(ma_r, state0, ties, n_steps=200)
    assert out_r["rewards"].mean() < 0.5 * rew.mean(), out_r["rewards"].mean()


def test_reference_elites_known_answers_on_device():
    """The reference's saved fitness values (59 x 200.0 and one 199.1 over the 60 distinct sNES elites, gym
    seed-0 start; tests/test_oracle_cpu.py::test_reference_elites_known_answers) reproduced by the CUDA actor
    through snn_actor_episodes, with tie counts and final CartPole states bit-equal to the oracle's."""
    from oracle import py_oracle as po
    from rl_stdp_b200.lif import ActorPopulation
    from tests.helpers import gym_seed0_reset_state, reference_elites
    s0 = gym_seed0_reset_state()
    el = reference_elites()
    pop = ActorPopulation([8, 8, 8], len(el))
    for l in range(2):
        pop.set_weights(l, np.array([w[l] for _, w, _ in el]))
    for tie in (0, 1):
        out = pop.play_episodes(np.array([m for m, _, _ in el]), np.tile(s0, (len(el), 1)),
                                np.full((len(el), 200), tie, np.int32), n_steps=200)
        fit = np.array([f for _, _, f in el])
        assert np.array_equal(out["rewards"], fit)
        assert np.all(out["n_decisions"] == 200) and np.array_equal(out["n_rand"], (fit != 200.0).astype(np.int32))
        for a, (m, w, _) in enumerate(el):
            o = po.Layered([8, 8, 8])
            for l in range(2):
                o.e(l)[:, :] = w[l]
            _, _, _, so = o.play_episode(m, s0, np.full(200, tie, np.int32))
            assert np.array_equal(out["state"][a], so), a


def test_reference_mountcar_elites_known_answers_on_device():
    """The ten non-saturated MountainCar answers (tests/test_oracle_cpu.py::
    test_reference_mountcar_elites_known_answers) with every 40-ms decision window run by the CUDA actor
    (snn_actor_window); environment, read-out and Julia's tie stream on the host."""
    from rl_stdp_b200.lif import ActorPopulation
    from tests.helpers import mountcar_elites, play_mountcar
    from rl_stdp_b200.julia_rng import julia_mt_pick_bits
    el = mountcar_elites()
    pop = ActorPopulation([8, 8, 8], len(el))
    for l in range(2):
        pop.set_weights(l, np.array([w[l] for _, w, _ in el]))
    tot, nd, ties = play_mountcar(lambda it: pop.decision_window(it, 40, reset_v=True)[0],
                                  [m for m, _, _ in el], julia_mt_pick_bits(0, 64))
    assert tot.tolist() == [f for _, _, f in el]
    assert nd.tolist() == [121, 122, 118, 126, 111, 113, 121, 118, 117, 116]
    assert ties.tolist() == [8, 7, 9, 1, 18, 12, 4, 5, 4, 1]


def test_mountcar_episodes_on_device():
    """snn_actor_episodes_env(SNN_ENV_MOUNTAINCAR): the whole mountcar_fitness.jl play_episode on the device.
    (1) the reference's ten saved results with Julia's default tie stream; (2) random genomes, random starts
    and random tie streams bit-equal to the oracle (reward, decisions, ties, final position and velocity)."""
    from oracle import py_oracle as po
    from rl_stdp_b200.lif import ActorPopulation, borne, set_weight
    from tests.helpers import _gym_np_random, mountcar_elites
    el = mountcar_elites()
    p0 = _gym_np_random(0).uniform(low=-0.6, high=-0.4)
    pop = ActorPopulation([8, 8, 8], len(el))
    for l in range(2):
        pop.set_weights(l, np.array([w[l] for _, w, _ in el]))
    out = pop.play_episodes(np.array([m for m, _, _ in el]), np.tile([p0, 0.0], (len(el), 1)), env="mountaincar")
    assert out["rewards"].tolist() == [f for _, _, f in el]
    assert out["n_decisions"].tolist() == [121, 122, 118, 126, 111, 113, 121, 118, 117, 116]
    assert out["n_rand"].tolist() == [8, 7, 9, 1, 18, 12, 4, 5, 4, 1]

    arch, A = [8, 8, 9], 64   # 9 outputs: pools of 3, 3, 3
    rng = np.random.Generator(np.random.Philox(key=91))
    genes = rng.standard_normal((A, 16 + 8 * 8 + 8 * 9 + 16)) * 4.0
    ma = np.array([borne(g[:16], -1.0, 1.0).reshape((2, 8)).T for g in genes])
    ws = [set_weight(arch, g[32:]) for g in genes]
    s0 = np.stack([rng.uniform(-0.6, -0.4, A), rng.uniform(-0.01, 0.01, A)], axis=1)
    ties = rng.integers(0, 2, size=(A, 600), dtype=np.int32)
    pop = ActorPopulation(arch, A)
    for l in range(2):
        pop.set_weights(l, np.array([w[l] for w in ws]))
    out = pop.play_episodes(ma, s0, ties, n_steps=200, env="mountaincar")
    n_tied = 0
    for a in range(A):
        o = po.Layered(arch)
        for l in range(2):
            o.e(l)[:, :] = ws[a][l]
        r, nd, nr, so = o.play_episode_mountcar(ma[a], s0[a], ties[a])
        assert (out["rewards"][a], out["n_decisions"][a], out["n_rand"][a]) == (r, nd, nr), a
        assert np.array_equal(out["state"][a], so), a
        n_tied += nr
    assert n_tied > 0 and out["n_decisions"].min() < 200   # ties and early finishes both exercised
