"""Parameter files of the reference: `YAML.load(open("…/cfg.yml"))` of
Julia/Modules/Networks/net_res.jl:4, net_arch.jl:4 and Julia/Actors/cartpole_fitness.jl (cfg*.yml).

`load_params` reads such a file into the same flat Dict the Julia code indexes (`param["apost"]` …);
`izhikevich_config` / `actor_config` turn it into the keyword arguments of SparseNetwork
(snn_config) and ActorPopulation (snn_actor_config).  Keys the hot path never reads (`b`, `m`,
`std`, plotting switches) are carried along untouched; a key the hot path needs and the file lacks
raises KeyError, as the Julia Dict lookup would.
"""
from __future__ import annotations

import yaml

# keys read by the Izhikevich variants D / E (net_res.jl, net_arch.jl) and by the LIF actor (F)
IZH_KEYS = ("c", "ae", "ai", "de", "di", "theta", "ttheta", "apost", "apre", "tda", "wmax")
LIF_KEYS = ("vthresh", "c", "wei", "wie", "wmax", "theta", "ttheta", "imin", "apost", "apre", "tde", "ttrace", "taulif")


def load_params(path: str) -> dict:
    with open(path) as f:
        d = yaml.safe_load(f)
    if not isinstance(d, dict):
        raise ValueError(f"{path}: expected a flat mapping of parameter names to numbers")
    return {str(k): v for k, v in d.items()}


def izhikevich_config(param: dict, layered: bool) -> dict:
    """snn_config fields for `Network(arch, param)`.

    layered=False: net_res.jl — threshold 30 hard-coded (:167), learn! = (0.002+da), sd *= 0.99 (:211-223).
    layered=True : net_arch.jl Izhikevich path — threshold param["vthresh"] (:331), learn! e += da*de with
                   de never decayed (:213-220)."""
    for k in IZH_KEYS:
        param[k]  # KeyError for a missing key, like the Julia Dict
    cfg = dict(v_reset=float(param["c"]), a_exc=float(param["ae"]), a_inh=float(param["ai"]), d_exc=float(param["de"]),
               d_inh=float(param["di"]), theta=float(param["theta"]), ttheta=float(param["ttheta"]),
               apost=float(param["apost"]), apre=float(param["apre"]), da_decay=float(param["tda"]),
               wmax=float(param["wmax"]))
    if layered:
        from . import _lib as L
        cfg.update(vthresh=float(param["vthresh"]), rate_mode=L.SNN_RATE_DA, sd_decay_mode=L.SNN_SD_DECAY_NEVER,
                   trace_decay=float(param.get("ttrace", 0.95)))
    else:
        cfg.update(vthresh=30.0)
    return cfg


def actor_config(param: dict) -> dict:
    """snn_actor_config fields for the layered LIF actor (net_arch.jl:142-272, cartpole_cfglif.yml)."""
    for k in LIF_KEYS:
        param[k]
    out = {k: float(param[k]) for k in LIF_KEYS}
    for k in ("eta", "winh", "da_inc"):
        if k in param:
            out[k] = float(param[k])
    return out
