"""The algebra behind the experimental event-driven learning, on the CPU: the numpy
prototype NetALazy (tests/tools/lazy_learn_study.py) keeps (s0, sd' = sd / r^k, C) per synapse and forms
s = clamp(s0 + H*sd' - C) when a weight is read.  M = 1 must be the literal variant A bit for bit; any M must
agree with it to rounding as long as no weight reaches a bound."""
import importlib.util
import os

import numpy as np

from oracle.np_restate import NetA

_spec = importlib.util.spec_from_file_location(
    "lazy_learn_study", os.path.join(os.path.dirname(os.path.abspath(__file__)), "tools", "lazy_learn_study.py"))
study = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(study)


def _run(net, t, seed):
    rng = np.random.Generator(np.random.Philox(key=seed))
    out = []
    for k in range(t):
        out.append(net.step(13.0 * (rng.random(net.n_neurons) - 0.5), train=((k + 1) % 10 == 0)).copy())
        if k == 150:
            net.da += 0.5
    return out


def test_m1_is_the_reference_rule_bit_for_bit():
    n, ne, t = 300, 240, 600
    conn = np.random.Generator(np.random.Philox(key=1)).random((n, n)) < 0.1
    a, b = NetA(n, ne, conn), study.NetALazy(n, ne, conn, 1, 1e9)
    sa, sb = _run(a, t, 2), _run(b, t, 2)
    assert all(np.array_equal(x, y) for x, y in zip(sa, sb)) and sum(len(x) for x in sa) > 100
    assert np.array_equal(b.final_weights(), a.s)
    assert np.array_equal(b.sdp, a.sd)          # after the final re-basing sd' is sd again
    assert np.abs(a.s[:ne][conn[:ne]] - 1.0).max() > 1e-3


def test_any_m_agrees_to_rounding_away_from_the_bounds():
    n, ne, t = 300, 240, 300
    conn = np.random.Generator(np.random.Philox(key=1)).random((n, n)) < 0.1
    a = NetA(n, ne, conn)
    sa = _run(a, t, 2)
    w = a.s[:ne][conn[:ne]]
    assert w.min() > 0.0 and w.max() < 4.0      # 300 ms: nothing has reached a bound yet
    for m in (3, 8):
        b = study.NetALazy(n, ne, conn, m, 1e9)
        sb = _run(b, t, 2)
        assert all(np.array_equal(x, y) for x, y in zip(sa, sb))
        assert np.abs(b.final_weights() - a.s).max() < 1e-12
        assert np.abs(b.sdp - a.sd).max() < 1e-12
        assert b.dense_learns <= b.learns // m + 1


def _csr(conn):
    n = conn.shape[0]
    rowptr = np.zeros(n + 1, np.int64)
    np.cumsum(conn.sum(axis=1), out=rowptr[1:])
    post = np.concatenate([np.flatnonzero(conn[i]) for i in range(n)]).astype(np.int32)
    return rowptr, post


def test_sparse_oracle_event_driven_mode_matches_the_prototype_with_clamps():
    """oracle/snn_oracle.c with lazy = M (the checker for the production build's event-driven learning) against
    the numpy prototype, dopamine high enough that weights pile up at both bounds: same spikes, weights and sd
    to 1e-9 — including the overshoot the clamp leaves inside a re-basing window — and NOT the exact rule."""
    from oracle import py_oracle as po
    n, ne, t, m = 300, 240, 900, 4
    conn = np.random.Generator(np.random.Philox(key=1)).random((n, n)) < 0.1
    rowptr, post = _csr(conn)
    w0 = np.where(np.repeat(np.arange(n), np.diff(rowptr)) < ne, 1.0, -1.0)

    def drive(step, add_da, finish):
        rng = np.random.Generator(np.random.Philox(key=2))
        out = []
        for k in range(t):
            out.append(np.array(step(13.0 * (rng.random(n) - 0.5), (k + 1) % 10 == 0)))
            if k % 150 == 100:
                add_da(0.5)
        return out, finish()

    proto = study.NetALazy(n, ne, conn, m, 1e9)
    sp_p, w_p = drive(lambda z, tr: proto.step(z, train=tr), lambda x: setattr(proto, "da", proto.da + x),
                      lambda: proto.final_weights()[conn].copy())
    lazy = po.Sparse(n, ne, rowptr, post, w0, lazy=m)
    sp_l, w_l = drive(lambda z, tr: lazy.step(z, tr), lambda x: lazy.add_da(0, x), lambda: lazy.get(4))
    exact = po.Sparse(n, ne, rowptr, post, w0)
    sp_e, w_e = drive(lambda z, tr: exact.step(z, tr), lambda x: exact.add_da(0, x), lambda: exact.get(4))
    plastic = np.repeat(np.arange(n), np.diff(rowptr)) < ne
    print("at 0:", (w_e[plastic] == 0).mean(), "at 4:", (w_e[plastic] == 4).mean(), "max|lazy-exact|:", np.abs(w_l - w_e).max())
    assert (w_e[plastic] == 0).mean() > 0.005   # the lower bound is populated
    assert all(np.array_equal(a, b) for a, b in zip(sp_p, sp_l))
    assert np.abs(w_p - w_l).max() < 1e-9
    assert np.abs(proto.sdp[conn][plastic] - lazy.get(5)[plastic]).max() < 1e-9   # (the prototype also keeps sd on inhibitory rows)
    assert np.abs(w_l - w_e).max() > 1e-3            # an approximation of the exact rule, not the rule itself
    assert abs(w_l[plastic].mean() - w_e[plastic].mean()) < 0.01 * w_e[plastic].mean()
    # M = 1 in the oracle is the exact rule bit for bit
    one = po.Sparse(n, ne, rowptr, post, w0, lazy=1)
    sp_1, w_1 = drive(lambda z, tr: one.step(z, tr), lambda x: one.add_da(0, x), lambda: one.get(4))
    assert all(np.array_equal(a, b) for a, b in zip(sp_1, sp_e)) and np.array_equal(w_1, w_e)
    # with the dopamine threshold (re-base at every learn step while da > 0.3) prototype and oracle agree too
    proto = study.NetALazy(n, ne, conn, m, 0.3)
    sp_p, w_p = drive(lambda z, tr: proto.step(z, train=tr), lambda x: setattr(proto, "da", proto.da + x),
                      lambda: proto.final_weights()[conn].copy())
    thr = po.Sparse(n, ne, rowptr, post, w0, lazy=m, lazy_da=0.3)
    sp_t, w_t = drive(lambda z, tr: thr.step(z, tr), lambda x: thr.add_da(0, x), lambda: thr.get(4))
    assert all(np.array_equal(a, b) for a, b in zip(sp_p, sp_t)) and np.abs(w_p - w_t).max() < 1e-9
    assert np.abs(w_t - w_e).max() < np.abs(w_l - w_e).max()      # and the threshold brings it closer to the rule
