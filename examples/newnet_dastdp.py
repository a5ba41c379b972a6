#!/usr/bin/env python
"""The reference's headline experiment, Julia/Izhikevic_dastdp/newnet_DASTDP.jl, on libsnn_b200:
1000 Izhikevich neurons (800 exc / 200 inh, p = 0.1), DA-STDP, learn! every 10 ms; the synapse
n1 -> n2 starts at 0 and every n1-then-n2 coincidence within 20 ms is rewarded with dopamine
(+0.5) 1-3 s later (:24-34, :52-56).  The ms loop runs on the GPU in 100-ms windows — the reference
samples `shist` every 100 ms (:57-60) and a reward is always >= 1001 ms in the future, so the reward
bookkeeping stays on the host between windows, exactly as in the Julia script.

  python examples/newnet_dastdp.py [--seconds 3600] [--mode fast|exact] [--out shist.csv]

Prints simulated seconds per wall second and the trajectory of the targeted synapse (the reference's
Images/*.png show it potentiating to the 4.0 clamp within the hour).
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rl_stdp_b200 import _lib as L, synth  # noqa: E402
from rl_stdp_b200.network import SparseNetwork  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=int, default=3600)
    ap.add_argument("--mode", default="fast", choices=["fast", "exact"])
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    N, Ne, W = 1000, 800, 100
    conn = synth.dense_connection(N, 0.1, a.seed)                         # Network(1000, 0.1, 800) :7
    rowptr, post, w = synth.dense_to_csr(synth.variant_a_weights(N, Ne, conn), conn)
    n1 = 1                                                                # const n1 = 2 (1-based) :15
    n2 = int(post[rowptr[n1]])                                            # post(n1, connection)[1] :16
    e_target = int(rowptr[n1])
    w[e_target] = 0.0                                                     # n.s[n1,n2] = 0.0 :19
    exact = a.mode == "exact"
    net = SparseNetwork(N, Ne, rowptr, post, w, dtype="f64" if exact else "f32", mode=a.mode,
                        noise="replay" if exact else "philox", noise_seed=a.seed + 1)
    rng = np.random.Generator(np.random.Philox(key=a.seed + 2))
    noise_rng = np.random.Generator(np.random.Philox(key=a.seed + 3))
    n1f, n2f, rew = [-100], [], {}
    train = np.array([(k + 1) % 10 == 0 for k in range(W)], np.uint8)     # msec % 10 == 0 :49
    shist = []
    interval = 20
    t0 = time.perf_counter()
    n_rewards = 0
    for win in range(a.seconds * 1000 // W):
        base = win * W                                                    # absolute ms of step 0 is base + 1
        da = [(t - base - 1, 0, 0.5) for t in sorted(rew) if base < t <= base + W]   # `if time in rew` :54: once per ms
        noise = 13.0 * (noise_rng.random((W, N)) - 0.5) if exact else None
        (ids, offs), _ = net.run(W, noise, train, da_events=da, raw=True)
        # reward!(...) :24-34 — only the steps in which n1 or n2 fired are looked at (the record is (ids, offsets))
        hits = np.flatnonzero((ids == n1) | (ids == n2))
        # (ms, 0 for n1 / 1 for n2): step k is ms base + k + 1; inside one ms n1 is looked at first, as in the script
        for t, is_n2 in sorted((base + int(np.searchsorted(offs, h, side="right")), int(ids[h] != n1)) for h in hits):
            if not is_n2:
                n1f.append(t)
            else:
                n2f.append(t)
                if t - n1f[-1] < interval and n2f[-1] > n1f[-1]:
                    tr = t + 1000 + int(rng.integers(1, 2001))
                    rew[tr] = rew.get(tr, 0) + 1
                    n_rewards += 1
        wt, sdt = net.get_synapses([e_target])                            # shist[...] :57-60: the target synapse only
        last = win + 1 == a.seconds * 1000 // W
        report = (win + 1) % (600 * 1000 // W) == 0 or last
        mean_s = float(net.get(L.FIELD_W).mean()) if report or (win + 1) % 100 == 0 else float("nan")  # every 10 s
        shist.append(((base + W) / 1000.0, wt[0], sdt[0], mean_s))
        if report:
            el = time.perf_counter() - t0
            print(f"t = {(base + W) / 1000:7.1f} s  s[n1,n2] = {wt[0]:.4f}  sd = {sdt[0]:+.4f}  mean s = {mean_s:.4f}  "
                  f"rewards = {n_rewards}  ({(base + W) / 1000 / el:.1f} simulated s per wall s)", flush=True)
    el = time.perf_counter() - t0
    print(f"{a.seconds} simulated s in {el:.1f} wall s ({a.seconds / el:.1f} x real time, mode {a.mode}); "
          f"target synapse {shist[0][1]:.3f} -> {shist[-1][1]:.3f}")
    if a.out:
        np.savetxt(a.out, np.array(shist), delimiter=",", header="t_s,s_n1_n2,sd_n1_n2,mean_s", comments="")


if __name__ == "__main__":
    main()
