"""Grid-cell state encoders on the device (include/snn_b200.h snn_gridcode_*; csrc/gridcode.cu): the distance tests
of Julia/Modules/Gridcells/StateSpikeGaussGrid.jl `input_spikes` / `latency_spikes` for a batch of observations.

The grids are the caller's (`HexGrid.grid`, `HexGrid.node_size` of HexGrid.jl / GridGenerator.jl): pass them as
[(nodes [k, 2], node_size), ...].  Nothing here computes on the CPU.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L


class GridCode:
    def __init__(self, grids, device=0):
        nodes = [np.ascontiguousarray(np.asarray(g, np.float64).reshape(-1, 2)) for g, _ in grids]
        self.n_grids = len(nodes)
        off = np.zeros(self.n_grids + 1, np.int64)
        np.cumsum([len(x) for x in nodes], out=off[1:])
        xy = np.ascontiguousarray(np.concatenate(nodes) if nodes else np.zeros((0, 2)))
        size = np.ascontiguousarray([float(s) for _, s in grids], np.float64)
        self._h = C.c_void_p()
        L.check(L.load().snn_gridcode_create(int(device), self.n_grids, off.ctypes.data_as(C.c_void_p),
                                             xy.ctypes.data_as(C.c_void_p), size.ctypes.data_as(C.c_void_p),
                                             C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            L.load().snn_gridcode_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def encode(self, obs, reach=1.5, win_size=0):
        """obs [n_obs, 4] -> (intensity [n_obs, 2*n_grids], slot [n_obs, 2*n_grids] or None)."""
        o = np.ascontiguousarray(np.asarray(obs, np.float64).reshape(-1, 4))
        inten = np.zeros((len(o), 2 * self.n_grids), np.float64)
        slot = np.zeros((len(o), 2 * self.n_grids), np.int32) if win_size > 0 else None
        L.check(L.load().snn_gridcode_encode(self._h, o.ctypes.data_as(C.c_void_p), len(o), float(reach), int(win_size),
                                             inten.ctypes.data_as(C.c_void_p),
                                             None if slot is None else slot.ctypes.data_as(C.c_void_p)))
        return inten, slot

    def input_spikes(self, obs, reach=1.5):
        """`input_spikes(obs, grid_cells)` StateSpikeGaussGrid.jl:83-109 per observation: [(ids, intensities), ...],
        0-based ids in the reference's push order (grid by grid: position neuron i, then velocity neuron n + i)."""
        inten, _ = self.encode(obs, reach)
        n = self.n_grids
        order = np.stack([np.arange(n), n + np.arange(n)], axis=1).reshape(-1)
        out = []
        for row in inten:
            ids = order[row[order] != 0.0]
            out.append((ids.astype(np.int32), row[ids]))
        return out

    def latency_spikes(self, obs, win_size):
        """`latency_spikes(obs, grid_cells, win_size)` :111-139 per observation: win_size id arrays (index 0 = the
        first millisecond of the window), each in the reference's push order."""
        _, slot = self.encode(obs, 1.25, win_size)
        n = self.n_grids
        order = np.stack([np.arange(n), n + np.arange(n)], axis=1).reshape(-1)
        out = []
        for row in slot:
            out.append([order[row[order] == ms].astype(np.int32) for ms in range(win_size)])
        return out
