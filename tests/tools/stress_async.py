#!/usr/bin/env python
"""Repeats the frozen-weight production-mode case (async learn pass) to expose rare races.
  python tests/tools/stress_async.py [--reps 20] [--exp-flags 0]"""
import argparse, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import py_oracle as po
from rl_stdp_b200 import _lib as L, synth
from rl_stdp_b200.network import SparseNetwork
from tests.helpers import run_oracle_sparse

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--exp-flags", type=int, default=0)
ap.add_argument("--learn-blocks", type=int, default=0)
ap.add_argument("--learn-async", default="force")
a = ap.parse_args()
n, ne, fan, dmax, t = 3000, 2400, 80, 12, 600
rowptr, post, w, delay = synth.fixed_fanout(n, ne, fan, dmax, seed=71)
noise = synth.thalamic_noise(t, n, seed=12)
train = np.array([(k + 1) % 10 == 0 for k in range(t)], np.uint8)
ora = po.Sparse(n, ne, rowptr, post, w, delay, max_delay=dmax, learn_base=0.0)
ref = run_oracle_sparse(ora, noise, train)
plastic = np.repeat(np.arange(n), np.diff(rowptr)) < ne
sd_o = ora.get(5)[plastic]
bad = 0
for rep in range(a.reps):
    net = SparseNetwork(n, ne, rowptr, post, w, delay, dtype="f64", mode="fast", max_delay=dmax, learn_base=0.0,
                        learn_async=a.learn_async, exp_flags=a.exp_flags, learn_blocks=a.learn_blocks)
    got = []
    for lo in range(0, t, 200):
        sp, _ = net.run(200, noise[lo:lo + 200], train[lo:lo + 200])
        got += sp
    ok_sp = all(np.array_equal(got[k], ref[k]) for k in range(t))
    sd_g = net.get(L.FIELD_SD)[plastic]
    err = np.abs(sd_g - sd_o)
    nbad = int((err > 1e-12).sum())
    if not ok_sp or nbad:
        bad += 1
        ix = np.flatnonzero(err > 1e-12)[:8]
        print(f"rep {rep}: spikes_ok={ok_sp} bad_sd={nbad} max_err={err.max():.4g} at {ix} got={sd_g[ix]} want={sd_o[ix]}")
        pre = np.repeat(np.arange(n), np.diff(rowptr))[plastic][ix]; pe = np.flatnonzero(plastic)[ix]
        print("   pre", pre, "post", post[pe], "delay", delay[pe])
    net.close()
print(f"exp_flags={a.exp_flags}: {bad}/{a.reps} bad")
