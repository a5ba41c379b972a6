"""Offline study (not a test): can ONE tie-pick sequence — the unknown stream of Julia's
`rand(MersenneTwister(0), pool)` in Actors/mountcar_fitness.jl:48-107 — reproduce the ten fitness values
saved in Actors/sNESelites_mountcar.jld on the oracle's LIF step + a restated MountainCar-v0?
Every episode re-creates MersenneTwister(0), so the i-th tie of every episode sees the same i-th draw.
Joint depth-first search over the bit sequence; run in the build container (reads /root/reference)."""
import hashlib
import math
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import py_oracle as O  # noqa: E402
from rl_stdp_b200 import jld  # noqa: E402
from rl_stdp_b200.lif import borne, set_weight  # noqa: E402

ARCH = [8, 8, 8]


def start_position():
    h = int.from_bytes(hashlib.sha512(b"0").digest()[:8], "little")
    rs = np.random.RandomState()
    rs.seed([h % 2**32, h >> 32])
    return rs.uniform(-0.6, -0.4)


class NeedBit(Exception):
    pass


def episode(g, bits, p0):
    m = borne(g[:16], -1.0, 1.0).reshape((2, 8)).T
    w = set_weight(ARCH, g[32:])
    n = O.Layered(ARCH)
    for l in range(2):
        n.e(l)[:, :] = w[l]
    pos, vel, tot, used = p0, 0.0, 0.0, 0
    idx = np.arange(8, dtype=np.int32)
    for j in range(200):
        it = m @ np.array([(pos + 1.2) / 1.8, (vel + 0.07) / 0.14])
        cnt = np.zeros(8)
        n.v[:] = 0.0
        for ms in range(1, 41):
            spk = n.step_lif(idx, it, False)
            if ms >= 2:
                for x in spk:
                    if 16 <= x < 24:
                        cnt[x - 16] += 1
        sl, sm, sr = cnt[:2].sum(), cnt[2:4].sum(), cnt[4:].sum()
        a, fac = int(np.argmax([sl, sm, sr])), 1
        for d, pool in zip([sl - sr, sl - sm, sm - sr], [[1, 3], [1, 2], [2, 3]]):
            if d == 0.0:
                if used >= len(bits):
                    raise NeedBit()
                a, fac = pool[bits[used]] - 1, 2
                used += 1
        vel += (a - 1) * 0.001 + math.cos(3 * pos) * (-0.0025)
        vel = min(max(vel, -0.07), 0.07)
        pos = min(max(pos + vel, -1.2), 0.6)
        if pos == -1.2 and vel < 0:
            vel = 0.0
        tot -= fac
        if pos >= 0.5 and vel >= 0:
            break
    return tot, j + 1, used


def main():
    p0 = start_position()
    el = jld.load("/root/reference/Julia/Actors/sNESelites_mountcar.jld", "elites")
    genes = [np.asarray(e["genes"], float) for e in el]
    fit = [float(np.ravel(e["fitness"])[0]) for e in el]
    cache = {}

    def run(i, prefix):
        for k in range(len(prefix) + 1):  # an episode that finished on a shorter prefix is final
            if (i, prefix[:k]) in cache and cache[(i, prefix[:k])] is not None:
                return cache[(i, prefix[:k])]
        try:
            r = episode(genes[i], prefix, p0)
            r = (r[0], r[1], r[2])
            cache[(i, prefix[:r[2]])] = r
            return r
        except NeedBit:
            return None

    sols = []

    def dfs(prefix):
        pending = False
        for i in range(len(genes)):
            r = run(i, prefix)
            if r is None:
                pending = True
            elif r[0] != fit[i]:
                return
        if not pending:
            sols.append(prefix)
            print("consistent sequence:", prefix, [run(i, prefix) for i in range(len(genes))], flush=True)
            return
        if len(prefix) >= 40:
            return
        dfs(prefix + (0,))
        dfs(prefix + (1,))

    dfs(())
    print(len(sols), "consistent sequences")


if __name__ == "__main__":
    main()
