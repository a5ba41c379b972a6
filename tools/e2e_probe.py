import sys, time, os
sys.path.insert(0, "/root/repo")
import numpy as np, bench
from rl_stdp_b200 import synth
from rl_stdp_b200.network import SparseNetwork
wl = bench.WORKLOADS[sys.argv[1]]
n, ne, fan, dmax = wl[:4]; B = wl[4] if len(wl) > 4 else 1
rowptr, post, w, delay = synth.fixed_fanout(n, ne, fan, dmax, seed=9, n_instances=B)
net = SparseNetwork(n, ne, rowptr, post, w, delay, n_instances=B, dtype="f32", mode="fast", noise="philox", max_delay=dmax, noise_seed=10, learn_async=sys.argv[2])
W = 100
for kw in range(12):
    tr, da = bench.window_inputs(kw, W)
    rec = kw >= 4
    t0 = time.perf_counter()
    sp, st = net.run(W, None, tr, da_events=da, record=rec, spikes_cap=1 << 24, profile=False if rec else 'learn')
    print(kw, "record" if rec else "norec", "wall_ms %.2f" % ((time.perf_counter() - t0) * 1e3), "dev_ms %.2f" % st.device_ms, "spikes", st.n_spikes)
