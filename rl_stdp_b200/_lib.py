"""ctypes binding of libsnn_b200.so — exactly the symbols include/snn_b200.h declares.

This is the Python twin of the Julia `ccall` layer in julia/RLSTDPB200.jl (Julia is the
reference's host language but is not installed in the build image).  There is no fallback: if
the shared library is missing, loading raises.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SNN_B200_LIB") or os.path.join(HERE, "libsnn_b200.so")  # override: A/B builds only

SNN_F32, SNN_F64 = 0, 1
SNN_MODE_FAST, SNN_MODE_EXACT = 0, 1
SNN_NOISE_NONE, SNN_NOISE_REPLAY, SNN_NOISE_PHILOX = 0, 1, 2
SNN_SD_DECAY_ON_LEARN, SNN_SD_DECAY_EVERY_STEP, SNN_SD_DECAY_NEVER = 0, 1, 2
SNN_RATE_BASE_PLUS_DA, SNN_RATE_DA, SNN_RATE_ETA_DA = 0, 1, 2
(FIELD_V, FIELD_U, FIELD_TRACE, FIELD_HO, FIELD_W, FIELD_SD, FIELD_DA, FIELD_I) = range(8)
SNN_ERR_CAPACITY = -5
SNN_TL_CLASSES, SNN_TL_STEPS = 16, 258
SNN_PEER_BLOB_BYTES = 192


class SnnConfig(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32), ("dtype", C.c_int32), ("mode", C.c_int32), ("device", C.c_int32),
        ("n_neurons", C.c_int64), ("n_per_instance", C.c_int64), ("n_exc", C.c_int64),
        ("max_delay", C.c_int32), ("threshold_ge", C.c_int32),
        ("vthresh", C.c_double), ("v_reset", C.c_double),
        ("a_exc", C.c_double), ("a_inh", C.c_double), ("d_exc", C.c_double), ("d_inh", C.c_double),
        ("trace_set", C.c_double), ("trace_decay", C.c_double), ("apost", C.c_double), ("apre", C.c_double),
        ("da_decay", C.c_double), ("learn_base", C.c_double), ("eta", C.c_double),
        ("wmin", C.c_double), ("wmax", C.c_double), ("sd_decay", C.c_double),
        ("sd_decay_mode", C.c_int32), ("rate_mode", C.c_int32), ("noise_mode", C.c_int32),
        ("plastic_ei", C.c_int32), ("noise_amp", C.c_double), ("noise_seed", C.c_uint64),
        ("theta", C.c_double), ("ttheta", C.c_double), ("ho_from", C.c_int64),
        ("ho_inh", C.c_int32), ("variant_g", C.c_int32),
        ("reserved", C.c_int32 * 8),
    ]


class SnnActorConfig(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32), ("device", C.c_int32), ("n_agents", C.c_int32), ("n_layers", C.c_int32),
        ("arch", C.c_int32 * 8),
        ("vthresh", C.c_double), ("c", C.c_double), ("wei", C.c_double), ("wie", C.c_double), ("wmax", C.c_double),
        ("theta", C.c_double), ("ttheta", C.c_double), ("imin", C.c_double), ("apost", C.c_double),
        ("apre", C.c_double), ("tde", C.c_double), ("ttrace", C.c_double), ("taulif", C.c_double),
        ("eta", C.c_double), ("winh", C.c_double), ("da_inc", C.c_double), ("reserved", C.c_int32 * 8),
    ]


class SnnDaEvent(C.Structure):
    _fields_ = [("step", C.c_int32), ("instance", C.c_int32), ("amount", C.c_double)]


class SnnForceEvent(C.Structure):
    _fields_ = [("step", C.c_int32), ("neuron", C.c_int32)]


class SnnCurrentEvent(C.Structure):
    _fields_ = [("step", C.c_int32), ("neuron", C.c_int32), ("current", C.c_double)]


class SnnStepStats(C.Structure):
    _fields_ = [
        ("n_spikes", C.c_int64), ("n_syn_events", C.c_int64), ("n_ltp_events", C.c_int64),
        ("device_ms", C.c_double), ("learn_ms", C.c_double), ("neuron_ms", C.c_double),
        ("synapse_ms", C.c_double), ("learn_launches", C.c_int32), ("neuron_launches", C.c_int32),
        ("synapse_launches", C.c_int32), ("reserved", C.c_int32),
    ]


EXCHANGE_FN = C.CFUNCTYPE(C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p)

class SnnStepIO(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32), ("n_steps", C.c_int32), ("noise", C.c_void_p), ("train_flags", C.c_void_p),
        ("force", C.c_void_p), ("da", C.c_void_p), ("currents", C.c_void_p),
        ("n_force", C.c_int32), ("n_da", C.c_int32), ("n_currents", C.c_int32), ("reserved", C.c_int32),
        ("spikes_out", C.c_void_p), ("spike_offsets", C.c_void_p), ("spikes_cap", C.c_int64),
        ("stats", C.POINTER(SnnStepStats)),
        ("probe_syn", C.c_void_p), ("probe_out", C.c_void_p), ("n_probes", C.c_int32), ("reserved2", C.c_int32),
    ]


# every symbol of include/snn_b200.h: name -> (restype, argtypes)
_VP = C.c_void_p
SYMBOLS = {
    "snn_last_error": (C.c_char_p, []),
    "snn_abi_version": (C.c_int32, []),
    "snn_config_default": (C.c_int32, [C.POINTER(SnnConfig)]),
    "snn_create": (C.c_int32, [C.POINTER(SnnConfig), C.POINTER(_VP)]),
    "snn_destroy": (C.c_int32, [_VP]),
    "snn_set_synapses_csr": (C.c_int32, [_VP, _VP, _VP, _VP, _VP]),
    "snn_set_synapses_dense": (C.c_int32, [_VP, _VP, _VP]),
    "snn_set_synapses_dense_col": (C.c_int32, [_VP, _VP, _VP]),
    "snn_set_param": (C.c_int32, [_VP, C.c_char_p, C.c_double]),
    "snn_get_param": (C.c_int32, [_VP, C.c_char_p, C.POINTER(C.c_double)]),
    "snn_checkpoint_bytes": (C.c_int32, [_VP, C.POINTER(C.c_int64)]),
    "snn_checkpoint": (C.c_int32, [_VP, _VP, C.c_int64]),
    "snn_restore": (C.c_int32, [_VP, _VP, C.c_int64]),
    "snn_step": (C.c_int32, [_VP, C.c_int32, _VP, _VP, _VP, C.c_int32, _VP, C.c_int32, _VP, _VP, C.c_int64,
                             C.POINTER(SnnStepStats)]),
    "snn_step_ex": (C.c_int32, [_VP, C.POINTER(SnnStepIO)]),
    "snn_add_dopamine": (C.c_int32, [_VP, C.c_int32, C.c_double]),
    "snn_reset_v": (C.c_int32, [_VP]),
    "snn_get_array": (C.c_int32, [_VP, C.c_int32, _VP, C.c_int64]),
    "snn_get_synapses": (C.c_int32, [_VP, _VP, C.c_int32, _VP, _VP]),
    "snn_set_array": (C.c_int32, [_VP, C.c_int32, _VP, C.c_int64]),
    "snn_n_synapses": (C.c_int64, [_VP]),
    "snn_step_index": (C.c_int64, [_VP]),
    "snn_set_profiling": (C.c_int32, [_VP, C.c_int32]),
    "snn_set_stream": (C.c_int32, [_VP, _VP]),
    "snn_debug_timeline": (C.c_int32, [_VP, _VP, C.c_int64]),
    "snn_set_partition": (C.c_int32, [_VP, C.c_int64, C.c_int64, _VP, _VP]),
    "snn_partition_export": (C.c_int32, [_VP, _VP]),
    "snn_partition_connect": (C.c_int32, [_VP, C.c_int32, C.c_int32, _VP]),
    # batched LIF actors
    "snn_actor_config_default": (C.c_int32, [C.POINTER(SnnActorConfig)]),
    "snn_actor_create": (C.c_int32, [C.POINTER(SnnActorConfig), C.POINTER(_VP)]),
    "snn_actor_destroy": (C.c_int32, [_VP]),
    "snn_actor_get_array": (C.c_int32, [_VP, C.c_int32, C.c_int32, _VP, C.c_int64]),
    "snn_actor_set_array": (C.c_int32, [_VP, C.c_int32, C.c_int32, _VP, C.c_int64]),
    "snn_actor_window": (C.c_int32, [_VP, C.c_int32, _VP, C.c_int32, C.c_int32, C.c_int32, _VP, _VP, _VP]),
    "snn_actor_learn": (C.c_int32, [_VP, C.c_int32]),
    "snn_actor_episodes": (C.c_int32, [_VP, _VP, _VP, _VP, _VP, C.c_int32, C.c_int32, C.c_int32, _VP, _VP, _VP, _VP,
                                       _VP]),
    "snn_actor_episodes_env": (C.c_int32, [_VP, C.c_int32, _VP, _VP, _VP, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                           _VP, _VP, _VP, _VP, _VP]),
    # grid-cell encoders
    "snn_gridcode_create": (C.c_int32, [C.c_int32, C.c_int32, _VP, _VP, _VP, C.POINTER(_VP)]),
    "snn_gridcode_destroy": (C.c_int32, [_VP]),
    "snn_gridcode_encode": (C.c_int32, [_VP, _VP, C.c_int32, C.c_double, C.c_int32, _VP, _VP]),
}

SNN_ENV_CARTPOLE, SNN_ENV_MOUNTAINCAR = 0, 1

_lib = None


class SnnError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libsnn_b200 error {code}: {msg}")
        self.code = code


def load() -> C.CDLL:
    """Load libsnn_b200.so (building is the job of rl_stdp_b200.build / __graft_entry__.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise OSError(f"{LIB_PATH} is missing: run `python -m rl_stdp_b200.build` (there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)  # raises AttributeError if the export is missing
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int):
    if rc != 0:
        raise SnnError(rc, load().snn_last_error().decode("utf-8", "replace"))
