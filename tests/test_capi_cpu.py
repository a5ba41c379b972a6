"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/snn_b200.h
declares, validates configs, and — with no CUDA device — refuses to run instead of falling back."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lib():
    from rl_stdp_b200 import build, _lib as L
    build.build()
    return L


def test_header_symbols_all_exported():
    L = _lib()
    lib = L.load()
    hdr = open(os.path.join(ROOT, "include", "snn_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(snn_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 16
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/snn_b200.h but not exported"
    assert names == set(L.SYMBOLS), names ^ set(L.SYMBOLS)
    assert lib.snn_abi_version() == 3


def test_config_struct_matches_header():
    L = _lib()
    cfg = L.SnnConfig()
    L.check(L.load().snn_config_default(C.byref(cfg)))
    assert cfg.struct_size == C.sizeof(L.SnnConfig)
    # variant A constants (Old_net/new_net.jl:80-142)
    assert (cfg.vthresh, cfg.v_reset, cfg.a_exc, cfg.a_inh, cfg.d_exc, cfg.d_inh) == (30.0, -65.0, 0.02, 0.1, 8.0, 2.0)
    assert (cfg.trace_set, cfg.trace_decay, cfg.apost, cfg.apre) == (0.1, 0.95, 1.0, 1.5)
    assert (cfg.da_decay, cfg.learn_base, cfg.wmin, cfg.wmax, cfg.sd_decay) == (0.995, 0.002, 0.0, 4.0, 0.99)
    assert cfg.noise_amp == 13.0 and cfg.plastic_ei == 1


def test_invalid_configs_rejected():
    L = _lib()
    lib = L.load()
    for field, val, word in [("max_delay", 0, b"max_delay"), ("max_delay", 64, b"max_delay"), ("dtype", 7, b"dtype"),
                             ("n_per_instance", 333, b"multiple"), ("n_exc", 5000, b"n_exc"),
                             ("struct_size", 8, b"struct_size")]:
        cfg = L.SnnConfig()
        lib.snn_config_default(C.byref(cfg))
        setattr(cfg, field, val)
        h = C.c_void_p()
        assert lib.snn_create(C.byref(cfg), C.byref(h)) == -1
        assert word in lib.snn_last_error(), (field, lib.snn_last_error())
        assert not h.value
    assert lib.snn_step(None, 1, None, None, None, 0, None, 0, None, None, 0, None) == -1


def test_no_cpu_fallback():
    """Without a CUDA device the product path fails loudly (SNN_ERR_NODEVICE)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    L = _lib()
    lib = L.load()
    cfg = L.SnnConfig()
    lib.snn_config_default(C.byref(cfg))
    h = C.c_void_p()
    rc = lib.snn_create(C.byref(cfg), C.byref(h))
    assert rc == -3 and b"no CPU fallback" in lib.snn_last_error()


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under rl_stdp_b200/ may reference it."""
    pkg = os.path.join(ROOT, "rl_stdp_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "oracle/_" not in txt, f


def test_yaml_param_files_map_to_configs(tmp_path):
    """cfg*.yml -> the Dict the Julia code indexes -> snn_config / snn_actor_config fields."""
    from rl_stdp_b200 import _lib as L, params
    p = params.load_params(os.path.join(ROOT, "tests", "golden", "cfg_example.yml"))
    assert p["apost"] == 1.2 and p["wie"] == -1.5 and "b" in p
    d = params.izhikevich_config(p, layered=False)     # net_res.jl
    assert d == dict(v_reset=-65.0, a_exc=0.02, a_inh=0.1, d_exc=8.0, d_inh=2.0, theta=1.5, ttheta=0.998, apost=1.2,
                     apre=1.7, da_decay=0.99, wmax=3.0, vthresh=30.0)
    e = params.izhikevich_config(p, layered=True)      # net_arch.jl (Izhikevich)
    assert e["rate_mode"] == L.SNN_RATE_DA and e["sd_decay_mode"] == L.SNN_SD_DECAY_NEVER and e["trace_decay"] == 0.9
    cfg = L.SnnConfig()
    L.check(L.load().snn_config_default(C.byref(cfg)))
    for k, v in e.items():
        assert hasattr(cfg, k), k
        setattr(cfg, k, v)
    a = params.actor_config(p)
    acfg = L.SnnActorConfig()
    for k, v in a.items():
        assert hasattr(acfg, k), k
    assert a["taulif"] == 5.0 and a["imin"] == 10.0
    bad = tmp_path / "bad.yml"
    bad.write_text("apost: 1.0\n")
    with pytest.raises(KeyError):
        params.izhikevich_config(params.load_params(str(bad)), layered=False)


def test_ctypes_signatures_match_the_prototypes():
    """rl_stdp_b200/_lib.py SYMBOLS against include/snn_b200.h: every declared function is bound, with the
    prototype's number of arguments, pointers where the prototype has pointers and the right scalar widths."""
    import ctypes as C
    from rl_stdp_b200 import _lib as L
    hdr = " ".join(re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "snn_b200.h")).read(), flags=re.S).split())
    protos = {m.group(1): [a.strip() for a in m.group(2).split(",") if a.strip() != "void"]
              for m in re.finditer(r"\b(snn_[a-z0-9_]+)\s*\(([^()]*)\)\s*;", hdr)}
    assert set(protos) == set(L.SYMBOLS), set(protos) ^ set(L.SYMBOLS)
    scalar = {"int32_t": C.c_int32, "int64_t": C.c_int64, "double": C.c_double, "uint64_t": C.c_uint64}
    for name, (ret, args) in L.SYMBOLS.items():
        cargs = protos[name]
        assert len(args) == len(cargs), (name, len(args), cargs)
        for at, ca in zip(args, cargs):
            if "*" in ca or ca.startswith("snn_exchange_fn"):
                assert at is L._VP or hasattr(at, "contents") or issubclass(at, (C._Pointer, C._CFuncPtr, C.c_char_p)), (name, ca, at)
            else:
                assert at is scalar[ca.split()[0]], (name, ca, at)
