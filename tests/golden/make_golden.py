"""Generates the committed fixtures with the INDEPENDENT numpy restatement (oracle/np_restate.py),
so the C oracle is checked against vectors it did not produce.  Julia is not installed here, so
the vectors come from a restatement of the cited reference lines, not from the reference binary.
Run from the repo root: python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import np_restate as nr  # noqa: E402
from rl_stdp_b200 import synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def variant_a():
    N, Ne, T = 150, 120, 300
    conn = synth.dense_connection(N, 0.1, 1)
    noise = synth.thalamic_noise(T, N, 2, amp=17.0)  # amplitude raised so the fixture has >300 spikes
    net = nr.NetA(N, Ne, conn)
    da_steps = [99, 210]
    ids, offs = [], [0]
    for t in range(T):
        s = net.step(noise[t], (t + 1) % 10 == 0)
        ids.extend(int(x) for x in s); offs.append(len(ids))
        if t in da_steps:
            net.da += 0.5
    np.savez_compressed(os.path.join(HERE, "variant_a_n150_t300.npz"), N=N, Ne=Ne, T=T, connection=conn.astype(np.int8),
                        noise=noise, da_steps=np.array(da_steps), spike_ids=np.array(ids, np.int32),
                        spike_offsets=np.array(offs, np.int64), s_final=net.s, sd_final=net.sd, v_final=net.v,
                        da_final=net.da)
    print("variant A:", len(ids), "spikes")


if __name__ == "__main__":
    variant_a()
