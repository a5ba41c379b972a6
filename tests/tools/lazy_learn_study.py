"""Design study for DESIGN.md section 11.1 (not a test): how far does event-driven ("lazy") learning with
re-basing every M learn steps move the weight distribution of variant A away from the exact rule, given
that the clamp 0 <= s <= 4 is only applied at read time and at re-basing?

Exact rule (new_net.jl:135-142), every 10th ms:  s += (0.002 + da) * sd ; clamp ; sd *= 0.99
Lazy form: sd' = sd / r^k, H[k] = sum_{j<=k} (0.002 + da_j) r^(j-1), C = sum_events d' * H[k_e];
           s(k) = clamp(s0 + H[k] * sd' - C); dense pass (re-base) every M learn steps, and every learn step
           while da > da_dense.
Prints the two-sample KS distance of the excitatory weights against the exact run on the same inputs, next
to the KS distance between two exact runs that differ only in the noise seed (the chaos floor)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle.np_restate import NetA  # noqa: E402

R = 0.99


class NetALazy(NetA):
    def __init__(self, n, ne, conn, M, da_dense):
        super().__init__(n, ne, conn)
        self.M, self.da_dense = M, da_dense
        self.sdp = np.zeros((n, n))
        self.C = np.zeros((n, n))
        self.H, self.k, self.rpow = 0.0, 0, 1.0
        self.dense_learns = self.learns = 0
        self.emask = np.zeros(n, bool)
        self.emask[:ne] = True

    def weights_row(self, p):
        if not self.emask[p]:
            return self.s[p, :]
        x = self.s[p, :] + self.H * self.sdp[p, :] - self.C[p, :]
        return np.clip(x, 0.0, 4.0) * self.connection[p, :]

    def spike(self, noise):
        self.I = np.array(noise, dtype=np.float64)
        spiked = np.flatnonzero(self.v > 30)
        self.v[spiked] = -65.0
        self.u[spiked] += self.d[spiked]
        for p in spiked:
            self.I += self.weights_row(p)
        v, u, I = self.v, self.u, self.I
        self.v = v + 0.5 * ((0.04 * v + 5) * v + 140 - u + I)
        v = self.v
        self.v = v + 0.5 * ((0.04 * v + 5) * v + 140 - u + I)
        self.u = u + self.a * (0.2 * self.v - u)
        self.STDP[spiked] = 0.1
        for p in spiked:
            dl = 1.0 * self.STDP / self.rpow
            self.sdp[:, p] += dl
            self.C[:, p] += dl * self.H
            dl = 1.5 * self.STDP / self.rpow
            self.sdp[p, :] -= dl
            self.C[p, :] -= dl * self.H
        self.STDP *= 0.95
        self.da *= 0.995
        return spiked

    def rebase(self):
        e = self.exc
        x = self.s[e, :] + self.H * self.sdp[e, :] - self.C[e, :]
        self.s[e, :] = np.clip(x, 0.0, 4.0) * self.connection[e, :]
        self.sdp *= self.rpow
        self.C[:] = 0.0
        self.H, self.k, self.rpow = 0.0, 0, 1.0

    def learn(self):
        self.learns += 1
        g = 0.002 + self.da
        self.H += g * self.rpow
        self.k += 1
        self.rpow *= R
        if self.k >= self.M or self.da > self.da_dense:
            self.dense_learns += 1
            self.rebase()

    def final_weights(self):
        self.rebase()
        return self.s


def ks(a, b):
    a, b = np.sort(a), np.sort(b)
    allv = np.concatenate([a, b])
    return float(np.max(np.abs(np.searchsorted(a, allv, side="right") / len(a) -
                               np.searchsorted(b, allv, side="right") / len(b))))


def run(net, T, seed, da_period=1000):
    rng = np.random.Generator(np.random.Philox(key=seed))
    nsp = 0
    for t in range(T):
        noise = 13.0 * (rng.random(net.n_neurons) - 0.5)
        nsp += len(net.step(noise, train=((t + 1) % 10 == 0)))
        if (t + 1) % da_period == 0:
            net.da += 0.5
    return nsp


def main():
    n, ne, T = 600, 480, int(os.environ.get("T_MS", "20000"))
    conn = (np.random.Generator(np.random.Philox(key=1)).random((n, n)) < 0.1)
    mask = conn[:ne, :]
    t0 = time.time()
    ex = NetA(n, ne, conn)
    s_ex = run(ex, T, 2)
    w_ex = ex.s[:ne, :][mask]
    ex2 = NetA(n, ne, conn)
    run(ex2, T, 3)
    w_ex2 = ex2.s[:ne, :][mask]
    print(f"exact: {s_ex} spikes, mean w {w_ex.mean():.4f}, at 0: {(w_ex == 0).mean():.3f}, at 4: {(w_ex == 4).mean():.3f}"
          f"  | chaos floor KS(exact seed 2, exact seed 3) = {ks(w_ex, w_ex2):.4f}   [{time.time() - t0:.0f} s]", flush=True)
    for M, dd in ((1, 1e9), (4, 1e9), (8, 1e9), (16, 1e9), (8, 0.05), (8, 0.01), (16, 0.01), (32, 0.01)):
        lz = NetALazy(n, ne, conn, M, dd)
        s_lz = run(lz, T, 2)
        w = lz.final_weights()[:ne, :][mask]
        print(f"M={M:3d} da_dense={dd:g}: spikes {s_lz} ({100.0 * (s_lz - s_ex) / s_ex:+.2f} %), mean w {w.mean():.4f}, "
              f"KS vs exact {ks(w, w_ex):.4f}, max|dw| {np.abs(w - w_ex).max():.3f}, dense passes "
              f"{lz.dense_learns}/{lz.learns}   [{time.time() - t0:.0f} s]", flush=True)


if __name__ == "__main__":
    main()
