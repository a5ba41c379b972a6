#!/usr/bin/env python
"""Replay saved elites of the reference on the device — what Julia/Actors/cartpole_tryout.jl does one episode
at a time through PyCall, as two library calls.

  python examples/replay_elites.py [path/to/sNESelites_cartpole_N.jld] [--episodes 30]

1. every elite from gym's seed-0 reset state with the MersenneTwister(0) tie stream (`play_episode(...,
   rng_seed = true)`, the run that produced the saved `fitness`): must print the saved values;
2. `test_random` (cartpole_tryout.jl:53-64): `--episodes` unseeded episodes per elite, mean reward.
Needs a CUDA device (there is no CPU fallback)."""
import argparse
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rl_stdp_b200 import jld  # noqa: E402
from rl_stdp_b200.lif import ActorPopulation, borne, set_weight  # noqa: E402


def gym_reset_state(seed=0):
    """gym's classic-control seeding: SHA-512(str(seed))[:8] -> two 32-bit words -> numpy RandomState"""
    h = int.from_bytes(hashlib.sha512(str(seed).encode()).digest()[:8], "little")
    rs = np.random.RandomState()
    rs.seed([h % 2**32, h >> 32])
    return rs.uniform(-0.05, 0.05, size=(4,))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("path", nargs="?", default=os.path.join(ROOT, "tests", "golden", "sNESelites_cartpole_1.jld"))
    ap.add_argument("--episodes", type=int, default=30)
    a = ap.parse_args()
    arch = [8, 8, 8]                                   # Julia/Actors/cartpole_cfglif.yml
    elites = jld.load(a.path, "elites")
    genes = [np.asarray(e["genes"], np.float64) for e in elites]
    saved = [float(np.ravel(e["fitness"])[0]) for e in elites]
    matalpha = [borne(g[:arch[0] * 4], -1.0, 1.0).reshape((4, arch[0])).T for g in genes]   # reshape(genes_alpha, (8, 4))
    weights = [set_weight(arch, g[arch[0] * 4:]) for g in genes]

    def population(rep):
        pop = ActorPopulation(arch, len(genes) * rep)
        for l in range(len(arch) - 1):
            pop.set_weights(l, np.array([w[l] for w in weights for _ in range(rep)]))
        return pop, np.array([m for m in matalpha for _ in range(rep)])

    pop, ma = population(1)
    out = pop.play_episodes(ma, np.tile(gym_reset_state(0), (len(genes), 1)))          # ties: MersenneTwister(0)
    print("saved fitness   :", saved)
    print("replayed        :", out["rewards"].tolist(), f"({out['device_ms']:.2f} ms on the device)")
    rng = np.random.default_rng()
    pop, ma = population(a.episodes)
    n = len(genes) * a.episodes
    out = pop.play_episodes(ma, rng.uniform(-0.05, 0.05, (n, 4)), rng.integers(0, 2, (n, 200), dtype=np.int32))
    print(f"test_random, {a.episodes} episodes per elite:", np.round(out["rewards"].reshape(len(genes), -1).mean(axis=1), 1).tolist(),
          f"({out['device_ms']:.2f} ms on the device for {n} episodes)")


if __name__ == "__main__":
    main()
