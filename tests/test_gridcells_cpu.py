"""The CHECKER of the device-side grid-cell encoders (oracle/gridcells.py, restatement of Julia/Modules/Gridcells/*.jl).
No reference output pins them (DESIGN.md section 0, row f): structural properties, the reproducible part of
grid_gen (Julia's seeded generator), and agreement with a literal loop-by-loop evaluation."""
import math

import numpy as np

from oracle import gridcells as gc

OBS = [0.523324424, 0.323242424243, 0.4242424225, 0.224253525253]   # the scratch observation of StateSpikeGaussGrid.jl:157


def test_ref_grid_is_hexagonal():
    g = gc.ref_grid(0.05)
    w, h = math.sqrt(3) * 0.05, 0.1
    assert g.min() >= -1 and g.max() <= 2 + 1e-12
    # nearest-neighbour distance of a hexagonal tiling with these steps: sqrt((w/2)^2 + (0.75 h)^2) = w
    d = np.sqrt(((g[:, None, :] - g[None, :400, :]) ** 2).sum(-1))
    d[d == 0] = 9
    assert abs(d.min() - w) < 1e-12
    assert np.allclose(math.hypot(w / 2, 0.75 * h), w)


def test_grid_gen_reproduces_julia_seeded_orientations():
    grids = gc.grid_gen()
    assert len(grids) == 4 + 8 + 16 + 32
    dil = [2 * math.sqrt(2), 2, math.sqrt(2), 1]
    k = 0
    for mod in range(4):
        for _ in range(2 ** (mod + 2)):
            assert grids[k].node_size == 0.25 * math.sqrt(3) * (0.03 * dil[mod]) * dil[mod]
            assert grids[k].grid.min() >= 0.0 and grids[k].grid.max() <= 1.0      # cropped to the unit square
            k += 1
    # the first orientation comes from rand(MersenneTwister(0)) = 0.8236475079774124
    theta0 = 0.8236475079774124 * (2 * math.pi / 6) - math.pi / 6
    want = gc.HexGrid(0.03 * dil[0], theta=theta0, node_size=1.0).grid
    assert np.array_equal(grids[0].grid, want)
    # coarser modules have fewer nodes
    n = [len(g.grid) for g in grids]
    assert np.mean(n[:4]) < np.mean(n[4:12]) < np.mean(n[12:28]) < np.mean(n[28:])


def _literal_input_spikes(obs, grid_cells, reach):
    """StateSpikeGaussGrid.jl:83-109 line by line (1-based ids)."""
    pos = [obs[0], math.fmod(obs[2], math.pi)]
    vel = [obs[1], obs[3]]
    ps = [(x + 15.0) / 30.0 for x in pos]
    vs = [(x + 15.0) / 30.0 for x in vel]
    n = len(grid_cells)
    out = []
    for i in range(1, n + 1):
        g = grid_cells[i - 1]
        d = g.node_size
        close_pos = [p for p in g.grid if math.sqrt((p[0] - ps[0]) ** 2 + (p[1] - ps[1]) ** 2) < reach * d]
        close_vel = [p for p in g.grid if math.sqrt((p[0] - vs[0]) ** 2 + (p[1] - vs[1]) ** 2) < reach * d]
        if close_pos:
            dd = math.sqrt((ps[0] - close_pos[0][0]) ** 2 + (ps[1] - close_pos[0][1]) ** 2)
            out.append((i, 0.999999 * math.exp(-2 * (dd / d))))
        if close_vel:
            dd = math.sqrt((vs[0] - close_vel[0][0]) ** 2 + (vs[1] - close_vel[0][1]) ** 2)
            out.append((n + i, 0.999999 * math.exp(-2 * (dd / d))))
    return out


def test_input_spikes_equals_literal_loops():
    grids = gc.grid_gen()
    rng = np.random.default_rng(5)
    n_hits = 0
    for obs in [OBS] + [list(rng.uniform([-2.4, -3, -0.2, -3], [2.4, 3, 0.2, 3])) for _ in range(6)]:
        ids, inten = gc.input_spikes(obs, grids)
        lit = _literal_input_spikes(obs, grids, 1.5)
        assert [i - 1 for i, _ in lit] == ids.tolist()
        assert np.allclose([x for _, x in lit], inten, rtol=0, atol=1e-15)
        assert np.all(inten > 0) and np.all(inten < 1) and np.all(ids < 2 * len(grids))
        n_hits += len(ids)
        # latency code: same neurons within the tighter radius, earlier for higher intensity
        win = 10
        lat = gc.latency_spikes(obs, grids, win)
        ids2, int2 = gc.input_spikes(obs, grids, reach=1.25)
        assert sorted(np.concatenate(lat).tolist()) == sorted(ids2.tolist())
        when = {int(i): ms for ms, a in enumerate(lat) for i in a}
        for i, x in zip(ids2, int2):
            assert when[int(i)] == win - int(math.floor(win * x)) - 1
        assert set(ids2.tolist()) <= set(ids.tolist())
    assert n_hits > 50


def test_hex_modules_over_the_state_box():
    borpos, bortheta, borv = (-4.8, 4.8), (-math.pi, math.pi), (-15.0, 15.0)
    mp = gc.HexModule(0.07, dilat=1.1, theta=0.2, bor_x=borpos, bor_y=bortheta)
    mv = gc.HexModule(0.07, dilat=1.1, theta=0.2, bor_x=borv, bor_y=borv)
    assert mp.n_grid == 4 and len(mp.grids) == 4
    for g in mp.grids:
        assert g.grid[:, 0].min() >= -4.8 and g.grid[:, 0].max() <= 4.8
        assert g.grid[:, 1].min() >= -4.8 * (2 * math.pi / 9.6) - 1e-9     # cropped to the box's aspect ratio
        assert g.node_size == mp.grids[0].node_size
    ids, inten = gc.input_spikes(OBS, [mp, mp], [mv, mv])
    assert np.all(ids < 16) and len(set(ids.tolist())) == len(ids)
    b = gc.binary_spikes(OBS, [mp, mp], [mv, mv])
    assert set(b.tolist()) <= set(ids.tolist())
    # the two identical modules answer identically, 4 ids apart
    first = [i for i in ids.tolist() if i < 4]
    second = [i - 4 for i in ids.tolist() if 4 <= i < 8]
    assert first == second


def test_restated_exp_is_within_one_ulp_of_libm():
    """gauss() uses the fdlibm exp written out in IEEE operations (oracle/lif_oracle.c::ora_exp; the device evaluates
    the same ones): against libm it may differ in the last bit, never more, over the range the encoders use."""
    from oracle import py_oracle as po
    rng = np.random.default_rng(3)
    xs = np.concatenate([-rng.random(20000) * 4.0, rng.random(2000) * 4.0, [0.0, -1e-10, -0.3465, -0.3466, -1.0397, -1.0398]])
    worst = 0.0
    for x in xs:
        a, b = po.exp_fdlibm(float(x)), math.exp(float(x))
        worst = max(worst, abs(a - b) / math.ulp(b))
    assert worst <= 1.0
    assert po.exp_fdlibm(0.0) == 1.0
