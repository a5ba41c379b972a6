"""CHECKER of the device-side grid-cell encoders (rl_stdp_b200/csrc/gridcode.cu): a restatement of
Julia/Modules/Gridcells/{HexGrid,GridGenerator,StateSpikeGaussGrid,StateSpikeGrid}.jl.

TEST INFRASTRUCTURE ONLY — nothing under rl_stdp_b200/ imports this.  It builds the grids the tests feed to the
device (the product takes grids from its caller: in the reference's workflow GridGenerator.jl stays on the Julia
side) and restates `input_spikes` / `latency_spikes` observation by observation.  `gauss` uses the oracle's
written-out fdlibm exp (oracle/lif_oracle.c::ora_exp), the same IEEE operations as the device, so intensities
compare bit for bit; against libm's exp it differs by at most one ulp (tests/test_gridcells_cpu.py).

Not pinned by reference outputs: the only saved encoder output (CartPole/Heatmaps/disp_spikesGrid.jld) was made
with grids drawn from an unseeded `rand()`.  `grid_gen()` itself IS reproducible: it calls `Random.seed!(0)`, and
Julia's generator is restated in julia_rng.py.  Ids are 0-based here (Julia's minus one).
"""
from __future__ import annotations

import math

import numpy as np

from rl_stdp_b200.julia_rng import DSFMT


def ref_grid(hex_size: float) -> np.ndarray:
    """HexGrid.jl ref_grid: the centres of a hexagonal tiling of [-1, 2]^2, in the reference's push order."""
    w = math.sqrt(3) * hex_size
    h = 2 * hex_size
    hi, lo = 2, -1
    def jdiv(x, y):                                    # Julia div for floats: round((x - rem(x, y)) / y)
        return int(round((x - math.fmod(x, y)) / y))
    nx = jdiv(float(hi - lo), w / 2)                   # 0:div(hi-lo, w/2)
    ny = jdiv(float(hi - lo), 0.75 * h)
    cx = [i * w / 2 + lo for i in range(nx + 1)]
    cy = [j * 0.75 * h + lo for j in range(ny + 1)]
    pts = [(cx[i], cy[j]) for i in range(len(cx)) for j in range(len(cy)) if (i + 1 + j + 1) % 2 == 0]
    return np.array(pts, np.float64).reshape(-1, 2)


def _rotate(grid: np.ndarray, theta: float) -> np.ndarray:
    """rotate_grid!: about (0.5, 0.5)"""
    c, s = math.cos(theta), math.sin(theta)
    x, y = grid[:, 0] - 0.5, grid[:, 1] - 0.5
    return np.stack([c * x + (-s) * y + 0.5, s * x + c * y + 0.5], axis=1)


def _trunc(grid, lo_x=0.0, hi_x=1.0, lo_y=0.0, hi_y=1.0):
    keep = (lo_x <= grid[:, 0]) & (grid[:, 0] <= hi_x) & (lo_y <= grid[:, 1]) & (grid[:, 1] <= hi_y)
    return grid[keep]


def _scale(grid, lo_x, hi_x, lo_y, hi_y):
    """scale_grid!: crop the unit square to the aspect ratio of the target box, then map onto it.
    Returns (grid, len) — the HexGrid method also multiplies node_size by len."""
    l1, l2 = hi_x - lo_x, hi_y - lo_y
    if l1 >= l2:
        ln = l1
        grid = _trunc(grid, 0.0, 1.0, 0.0, l2 / l1)
    else:
        ln = l2
        grid = _trunc(grid, 0.0, l1 / l2)
    return grid * ln + np.array([lo_x, lo_y]), ln


class HexGrid:
    """HexGrid.jl:5-27.  As in the reference, this constructor scales the POINTS only: node_size stays what
    the caller passed (grid_gen passes the size it wants in the unit square)."""

    def __init__(self, res, dilat=1.0, theta=0.0, bor_x=(0.0, 1.0), bor_y=(0.0, 1.0), node_size=1.0, scale=True):
        grid = ref_grid(res)
        if dilat != 1.0:
            grid = grid * dilat
        if theta != 0:
            grid = _rotate(grid, theta)
        if scale:
            grid, _ = _scale(grid, bor_x[0], bor_x[1], bor_y[0], bor_y[1])
        self.node_size = float(node_size)
        self.grid = grid


class HexModule:
    """HexGrid.jl:146-163: mod_size^2 copies of one grid, shifted against each other inside a cell."""

    def __init__(self, res, dilat=1.0, theta=0.0, bor_x=(0.0, 1.0), bor_y=(0.0, 1.0), mod_size=2):
        ref = HexGrid(res, scale=False)
        self.grids = []
        for placement in range(1, mod_size * mod_size + 1):
            g = HexGrid.__new__(HexGrid)
            grid = ref.grid.copy()
            # trans_grid! :165-189
            D = math.hypot(grid[0, 0] - grid[1, 0], grid[0, 1] - grid[1, 1])
            d_trans = D / math.sqrt(3) / mod_size
            ti = d_trans * np.array([math.cos(math.pi / 3), math.sin(math.pi / 3)])
            tj = d_trans * np.array([math.cos(-math.pi / 3), math.sin(-math.pi / 3)])
            for _ in range((placement - 1) // mod_size):
                grid = grid + ti
            for _ in range((placement - 1) % mod_size):
                grid = grid + tj
            node = d_trans / 2
            grid = grid * dilat                       # dilat_grid!(::HexGrid)
            node *= dilat
            grid = _rotate(grid, theta)
            grid, ln = _scale(grid, bor_x[0], bor_x[1], bor_y[0], bor_y[1])
            g.grid, g.node_size = grid, node * ln
            self.grids.append(g)
        self.n_grid = mod_size * mod_size


def grid_gen():
    """GridGenerator.jl:11-27: 4 + 8 + 16 + 32 grids at four scales, orientations from `Random.seed!(0); rand()`."""
    rng = DSFMT([0])
    dilats = [2 * math.sqrt(2), 2, math.sqrt(2), 1]
    grids = []
    for mod in range(1, 5):
        for _ in range(2 ** (mod + 1)):
            res = 0.03 * dilats[mod - 1]
            node_size = 0.25 * math.sqrt(3) * res * dilats[mod - 1]
            theta = rng.rand() * (2 * math.pi / 6) - math.pi / 6
            grids.append(HexGrid(res, theta=theta, node_size=node_size))
    return grids


def gauss(d, node_size):
    """StateSpikeGaussGrid.jl:9 (never reaches one: the latency code indexes with it)"""
    from . import py_oracle as po
    return 0.999999 * po.exp_fdlibm(-2 * (d / node_size))


def _first_close(grid, p, radius):
    """`filter(q -> dist(q, p) < radius, grid)[1]`: the FIRST grid point inside the radius (grid order), or None"""
    d = np.sqrt((grid[:, 0] - p[0]) ** 2 + (grid[:, 1] - p[1]) ** 2)
    hit = np.flatnonzero(d < radius)
    return None if hit.size == 0 else float(d[hit[0]])


def _split(obs):
    return (obs[0], math.fmod(obs[2], math.pi)), (obs[1], obs[3])   # Julia % keeps the sign of the dividend


def input_spikes(obs, cells, cells_vel=None, reach=1.5):
    """(ids, intensities).  Two forms, as in the reference:
    input_spikes(obs, grid_cells::Array{HexGrid})            StateSpikeGaussGrid.jl:83-109 — position and
        velocity both scaled from [-15, 15] to the unit square, neuron i / n_grid + i;
    input_spikes(obs, cells_pos, cells_vel::Array{HexModule}) :12-44 — grids already laid over the state box."""
    pos, vel = _split(obs)
    ids, inten = [], []
    if cells_vel is None:
        ps = ((pos[0] + 15.0) / 30.0, (pos[1] + 15.0) / 30.0)
        vs = ((vel[0] + 15.0) / 30.0, (vel[1] + 15.0) / 30.0)
        n = len(cells)
        for i, g in enumerate(cells):
            for p, off in ((ps, 0), (vs, n)):
                d = _first_close(g.grid, p, reach * g.node_size)
                if d is not None:
                    ids.append(off + i)
                    inten.append(gauss(d, g.node_size))
    else:
        n_tot = sum(m.n_grid for m in cells)
        passed = 0
        for mp, mv in zip(cells, cells_vel):
            d_pos, d_vel = mp.grids[0].node_size, mv.grids[0].node_size
            for j in range(mp.n_grid):
                d = _first_close(mp.grids[j].grid, pos, reach * d_pos)
                if d is not None:
                    ids.append(passed + j)
                    inten.append(gauss(d, d_pos))
                d = _first_close(mv.grids[j].grid, vel, reach * d_vel)
                if d is not None:
                    ids.append(n_tot + passed + j)
                    inten.append(gauss(d, d_vel))
            passed += mp.n_grid
    return np.array(ids, np.int32), np.array(inten, np.float64)


def latency_spikes(obs, cells, win_size, cells_vel=None):
    """StateSpikeGaussGrid.jl:46-81, 111-139: the closer the state is to a node, the earlier its neuron fires
    inside the win_size-ms window.  Returns a list of win_size id arrays (index 0 = the first millisecond)."""
    ids, inten = input_spikes(obs, cells, cells_vel, reach=1.25)
    out = [[] for _ in range(win_size)]
    for i, x in zip(ids, inten):
        out[win_size - int(math.floor(win_size * x)) - 1].append(int(i))
    return [np.array(o, np.int32) for o in out]


def binary_spikes(obs, cells_pos, cells_vel):
    """StateSpikeGrid.jl:6-30: ids of the grids with a node within one node_size of the state."""
    ids, _ = input_spikes(obs, cells_pos, cells_vel, reach=1.0)
    return ids
