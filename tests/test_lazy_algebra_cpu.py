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
