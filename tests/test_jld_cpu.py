"""The importer for the reference's `.jld` result / elite files (rl_stdp_b200/jld.py).
Fixtures: two files written by the reference's own runs (data, not code) — the sNES elites that
Julia/Actors/cartpole_tryout.jl:15 loads, and a `save_param` results file (Modules/save_param.jl)."""
import glob
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def test_elites_file():
    from rl_stdp_b200 import jld
    elites = jld.load(os.path.join(GOLD, "sNESelites_cartpole_1.jld"), "elites")
    assert len(elites) == 10
    for e in elites:
        assert set(e) == {"genes", "fitness"}
        assert e["genes"].shape == (160,) and e["genes"].dtype == np.float64   # 8*4 input gains + 2*64 weights
        assert e["fitness"].shape == (1,) and e["fitness"][0] == 200.0          # every elite reached the episode cap
        assert np.isfinite(e["genes"]).all() and np.abs(e["genes"]).max() < 100


def test_results_dict_file():
    from rl_stdp_b200 import jld
    r = jld.load(os.path.join(GOLD, "results_LIFbase888.jld"))["results"]   # load_param, save_param.jl:34-47
    assert set(r) == {"weights", "fit"}
    assert r["fit"].shape == (20,) and r["fit"].max() == 200.0 and np.all(np.diff(r["fit"]) >= 0)
    assert len(r["weights"]) == 20 and all(w.shape == (160,) for w in r["weights"])


def test_genome_decoding_of_an_elite():
    """genes -> (matalpha, e[l]) exactly as fitness() does (cartpole_fitness.jl:158-176)."""
    from rl_stdp_b200 import jld
    from rl_stdp_b200.lif import borne, set_weight
    arch = [8, 8, 8]
    genes = jld.load(os.path.join(GOLD, "sNESelites_cartpole_1.jld"), "elites")[0]["genes"]
    matalpha = borne(genes[:arch[0] * 4], -1.0, 1.0).reshape((4, arch[0])).T   # reshape(genes_alpha, (arch[1], 4)), column-major
    assert matalpha.shape == (8, 4) and np.abs(matalpha).max() <= 1.0
    e = set_weight(arch, genes[arch[0] * 4:])
    assert [m.shape for m in e] == [(8, 8), (8, 8)]
    for l, m in enumerate(e):
        assert np.all(m[:2] <= 0.0) and np.all(m[:2] >= -2.0)       # the first div(8,4) neurons of a layer are inhibitory
        assert np.all(m[2:] >= 0.0) and np.all(m[2:] <= 1.1)


def test_every_jld_file_of_the_reference_parses():
    from rl_stdp_b200 import jld
    files = sorted(glob.glob("/root/reference/**/*.jld", recursive=True))
    if not files:
        pytest.skip("the reference tree is not mounted here")
    n_arrays = 0
    for f in files:
        top = jld.load(f)
        assert top, f

        def walk(x):
            nonlocal n_arrays
            if isinstance(x, dict):
                for v in x.values():
                    walk(v)
            elif isinstance(x, list):
                for v in x:
                    walk(v)
            elif isinstance(x, np.ndarray) and x.dtype != object:
                n_arrays += 1
                assert np.isfinite(x.astype(float)).all()
        walk(top)
    assert len(files) >= 40 and n_arrays > 1000


def test_bad_files_are_rejected(tmp_path):
    from rl_stdp_b200 import jld
    p = tmp_path / "x.jld"
    p.write_bytes(b"not an hdf5 file" * 10)
    with pytest.raises(jld.JldError):
        jld.load(str(p))
    with pytest.raises(KeyError):
        jld.load(os.path.join(GOLD, "sNESelites_cartpole_1.jld"), "no_such_variable")
