import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import py_oracle as po
from rl_stdp_b200 import _lib as L, synth
from rl_stdp_b200.network import SparseNetwork
n, ne, fan, dmax, t = 3000, 2400, 80, 12, 600
rowptr, post, w, delay = synth.fixed_fanout(n, ne, fan, dmax, seed=71)
noise = synth.thalamic_noise(t, n, seed=12)
train = np.array([(k + 1) % 10 == 0 for k in range(t)], np.uint8)
ora = po.Sparse(n, ne, rowptr, post, w, delay, max_delay=dmax, learn_base=0.0)
W = int(sys.argv[1]) if len(sys.argv) > 1 else 200
snaps = {}
spk = []
for k in range(t):
    spk.append(ora.step(noise[k], bool(train[k]), None))
    if (k + 1) % W == 0:
        snaps[k + 1] = ora.get(5).copy()
plastic = np.repeat(np.arange(n), np.diff(rowptr)) < ne
pre_all = np.repeat(np.arange(n), np.diff(rowptr))
for rep in range(14):
    net = SparseNetwork(n, ne, rowptr, post, w, delay, dtype="f64", mode="fast", max_delay=dmax, learn_base=0.0,
                        learn_async="force")
    for lo in range(0, t, W):
        net.run(W, noise[lo:lo + W], train[lo:lo + W])
        sd = net.get(L.FIELD_SD)
        err = np.abs(sd - snaps[lo + W]) * plastic
        bad = np.flatnonzero(err > 1e-12)
        if len(bad):
            e = bad[0]
            i, j, d = pre_all[e], post[e], delay[e]
            print(f"rep {rep} window ending {lo + W}: {len(bad)} bad, e={e} pre={i} post={j} delay={d} got={sd[e]:.6f} want={snaps[lo + W][e]:.6f}")
            ti = [k for k in range(lo + W) if i in spk[k]]
            tj = [k for k in range(lo + W) if j in spk[k]]
            print("   pre spikes", ti, " post spikes", tj)
            break
    net.close()
print("done")
