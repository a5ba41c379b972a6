"""Device-side grid-cell encoders (snn_gridcode_*, csrc/gridcode.cu) against the checker oracle/gridcells.py, which
restates Julia/Modules/Gridcells/StateSpikeGaussGrid.jl: 64 random observations x the 60 grids of grid_gen()
(GridGenerator.jl:11-27, seeded orientations) — ids, intensities (bit for bit) and latency slots."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _obs(n, seed):
    rng = np.random.default_rng(seed)
    # CartPole observation ranges with margins: x, x_dot, theta, theta_dot
    return np.stack([rng.uniform(-3, 3, n), rng.uniform(-4, 4, n), rng.uniform(-0.6, 0.6, n), rng.uniform(-5, 5, n)], axis=1)


def test_input_and_latency_spikes_bitexact_on_grid_gen():
    from oracle import gridcells as gc
    from rl_stdp_b200.gridcode import GridCode
    grids = gc.grid_gen()
    assert len(grids) == 60
    code = GridCode([(g.grid, g.node_size) for g in grids])
    obs = _obs(64, 5)
    obs[0] = [0.523324424, 0.323242424243, 0.4242424225, 0.224253525253]   # the scratch observation of :157
    got = code.input_spikes(obs)
    lat = code.latency_spikes(obs, 10)
    n_hits = 0
    for o, (ids, inten), slots in zip(obs, got, lat):
        want_ids, want_int = gc.input_spikes(o, grids)
        assert np.array_equal(ids, want_ids)
        assert np.array_equal(inten, want_int)            # bit for bit: same IEEE operations on both sides
        n_hits += len(ids)
        want_lat = gc.latency_spikes(o, grids, 10)
        for ms in range(10):
            assert np.array_equal(slots[ms], want_lat[ms]), ms
    assert n_hits > 64 * 10                                # the grids do cover the state space


def test_reach_and_empty_grid_and_errors():
    from oracle import gridcells as gc
    from rl_stdp_b200 import _lib as L
    from rl_stdp_b200.gridcode import GridCode
    g = gc.HexGrid(0.08, theta=0.3, node_size=0.05)
    empty = (np.zeros((0, 2)), 0.05)
    code = GridCode([(g.grid, g.node_size), empty, (g.grid[::-1].copy(), g.node_size)])
    obs = _obs(16, 9)
    for reach in (1.0, 1.25, 1.5, 3.0):
        inten, _ = code.encode(obs, reach)
        assert np.all(inten[:, 1] == 0) and np.all(inten[:, 4] == 0)        # the empty grid never fires
        for o, row in zip(obs, inten):
            ids, val = gc.input_spikes(o, [g], reach=reach)
            want = np.zeros(2)
            want[ids] = val
            assert row[0] == want[0] and row[3] == want[1]
    # the FIRST node inside the radius counts: the reversed grid may answer with another node when two are in reach
    wide, _ = code.encode(obs, 3.0)
    assert np.any(wide[:, 0] != wide[:, 2])
    with pytest.raises(L.SnnError):
        GridCode([(g.grid, 0.0)])
    with pytest.raises(L.SnnError):
        code.encode(obs, reach=-1.0)
