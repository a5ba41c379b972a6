"""Runs the OpenMP build of the CPU oracle on a synthetic network and saves its spike record.  A process of its
own: the oracle library variant (serial / OpenMP) is chosen when oracle.py_oracle is first imported, and the
1M / 100M network needs ~6 GB of host memory that should be gone when the GPU side runs.

  python tests/tools/full_size_oracle.py OUT.npz N NE FAN DMAX T TOPO_SEED NOISE_SEED [frozen]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
os.environ["ORACLE_OMP"] = "1"
import numpy as np  # noqa: E402
from oracle import py_oracle as po  # noqa: E402
from rl_stdp_b200 import synth  # noqa: E402


def main():
    out = sys.argv[1]
    n, ne, fan, dmax, t, seed, nseed = (int(x) for x in sys.argv[2:9])
    frozen = len(sys.argv) > 9 and sys.argv[9] == "frozen"
    threads = po.use_all_cores()
    rowptr, post, w, delay = synth.fixed_fanout(n, ne, fan, dmax, seed=seed)
    t0 = time.time()
    ora = po.Sparse(n, ne, rowptr, post, w, delay, max_delay=dmax)
    ids, offs = [], [0]
    for k in range(t):
        sp = ora.step(po.philox_noise(nseed, k, n), (not frozen) and (k + 1) % 10 == 0)
        ids.append(np.asarray(sp, np.int32).copy())
        offs.append(offs[-1] + len(sp))
    np.savez(out, ids=np.concatenate(ids) if ids else np.zeros(0, np.int32), offs=np.asarray(offs, np.int64),
             v=ora.get(0), threads=threads, seconds=time.time() - t0)


if __name__ == "__main__":
    main()
