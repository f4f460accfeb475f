"""ctypes binding of libfcwdm.so (C-ABI declared in include/fcwdm.h).

This is the thin stub a maintainer of the reference would add: the reference is pure
Python/PyTorch and has no FFI of its own, so the binding is new.  There is NO CPU or PyTorch fallback: if the
shared library is missing or the device is not a B200, importing / calling raises.
"""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FCWDM_LIB_PATH") or os.path.join(_HERE, "libfcwdm.so")   # override: A/B runs of two builds

_c_i64 = ctypes.c_int64
_c_p = ctypes.c_void_p
_c_f = ctypes.c_float
_c_int = ctypes.c_int

class ChainLayer(ctypes.Structure):
    """fcwdm_chain_layer of include/fcwdm.h: one layer of a fcwdm_conv3d_chain launch."""
    _fields_ = [("x", _c_p), ("x_ld", _c_i64), ("wp", _c_p), ("bias", _c_p), ("chan_bias", _c_p), ("cb_ld", _c_i64),
                ("residual", _c_p), ("res_ld", _c_i64), ("y", _c_p), ("y_ld", _c_i64), ("gn_stats", _c_p),
                ("gn_groups", _c_i64), ("gn_in_stats", _c_p), ("gn_in_gamma", _c_p), ("gn_in_beta", _c_p),
                ("gn_in_groups", _c_i64), ("gn_in_eps", _c_f), ("N", _c_i64), ("D", _c_i64), ("H", _c_i64), ("W", _c_i64),
                ("Cin", _c_i64), ("Cout", _c_i64), ("kind", _c_i64), ("aux", _c_p), ("aux_ld", _c_i64), ("aux_sb", _c_i64),
                ("lll_scale", _c_f), ("hi_scale", _c_f)]


# name -> (restype, argtypes); mirrors include/fcwdm.h one to one
PROTOTYPES = {
    "fcwdm_version": (_c_int, []),
    "fcwdm_last_error": (ctypes.c_char_p, []),
    "fcwdm_init": (_c_int, [_c_int]),
    "fcwdm_dwt3d_fwd": (_c_int, [_c_p, _c_p, _c_int] + [_c_i64] * 10 + [_c_f, _c_p]),
    "fcwdm_idwt3d_fwd": (_c_int, [_c_p, _c_p, _c_int] + [_c_i64] * 10 + [_c_f, _c_p]),
    "fcwdm_dwt3d_cl": (_c_int, [_c_p, _c_i64, _c_p, _c_i64, _c_p, _c_i64, _c_i64, _c_p] + [_c_i64] * 6 + [_c_f, _c_f, _c_p]),
    "fcwdm_idwt3d_cl": (_c_int, [_c_p, _c_i64, _c_p, _c_i64, _c_i64, _c_p, _c_i64, _c_p] + [_c_i64] * 6 + [_c_f, _c_p]),
    "fcwdm_p_sample_step": (_c_int, [_c_p, _c_i64, _c_p, _c_p, _c_p, _c_p, _c_p, _c_i64, _c_p, _c_p] + [_c_i64] * 5
                            + [_c_int, _c_int, _c_p]),
    "fcwdm_q_sample": (_c_int, [_c_p, _c_p, _c_p, _c_p, _c_p, _c_i64, _c_i64, _c_i64, _c_p]),
    "fcwdm_sample_to_image": (_c_int, [_c_p, _c_p, _c_p, _c_i64, _c_i64, _c_i64, _c_i64, _c_p]),
    "fcwdm_planar_to_cl": (_c_int, [_c_p, _c_p, _c_i64, _c_i64, _c_i64, _c_i64, _c_p]),
    "fcwdm_cl_to_planar": (_c_int, [_c_p, _c_i64, _c_p, _c_i64, _c_i64, _c_i64, _c_p]),
    "fcwdm_groupnorm_stats": (_c_int, [_c_p, _c_i64, _c_p, _c_i64, _c_i64, _c_i64, _c_i64, _c_p]),
    "fcwdm_groupnorm_apply": (_c_int, [_c_p, _c_i64, _c_p, _c_i64, _c_p, _c_p, _c_p, _c_i64, _c_i64, _c_i64, _c_i64,
                                       _c_f, _c_int, _c_p]),
    "fcwdm_timestep_embedding": (_c_int, [_c_p, _c_p, _c_i64, _c_i64, _c_f, _c_p]),
    "fcwdm_timestep_embedding_f32": (_c_int, [_c_p, _c_p, _c_i64, _c_i64, _c_f, _c_p]),
    "fcwdm_linear": (_c_int, [_c_p, _c_p, _c_p, _c_p, _c_i64, _c_i64, _c_i64, _c_int, _c_int, _c_p]),
    "fcwdm_conv3d_packed_elems": (_c_i64, [_c_i64, _c_i64, _c_int]),
    "fcwdm_conv3d_pack_weights": (_c_int, [_c_p, _c_p, _c_i64, _c_i64, _c_int, _c_p]),
    "fcwdm_conv3d_fwd": (_c_int, [_c_p, _c_i64, _c_p, _c_p, _c_p, _c_i64, _c_p, _c_i64, _c_p, _c_i64, _c_p, _c_i64]
                         + [_c_i64] * 6
                         + [_c_int, _c_p]),
    "fcwdm_conv3d_pair_supported": (_c_int, [_c_i64, _c_i64, _c_int]),
    "fcwdm_conv3d_pair_packed_elems": (_c_i64, [_c_i64, _c_i64]),
    "fcwdm_conv3d_pair_pack_weights": (_c_int, [_c_p, _c_p, _c_i64, _c_i64, _c_p]),
    "fcwdm_conv3d_pair_fwd": (_c_int, [_c_p, _c_i64, _c_p, _c_p, _c_p, _c_i64, _c_p, _c_i64, _c_p, _c_i64, _c_p, _c_i64,
                                       _c_p, _c_p, _c_p, _c_i64, _c_f] + [_c_i64] * 6 + [_c_p]),
    "fcwdm_conv3d_wgrad_workspace_bytes": (_c_i64, [_c_i64] * 6 + [_c_int]),
    "fcwdm_conv3d_wgrad": (_c_int, [_c_p, _c_i64, _c_p, _c_i64, _c_p, _c_p, _c_i64, _c_int] + [_c_i64] * 6 + [_c_int, _c_p]),
    "fcwdm_conv3d_transpose_flip_weights": (_c_int, [_c_p, _c_p, _c_i64, _c_i64, _c_int, _c_p]),
    "fcwdm_groupnorm_bwd": (_c_int, [_c_p, _c_i64, _c_p, _c_i64, _c_p, _c_p, _c_p, _c_p, _c_p, _c_i64, _c_p, _c_i64, _c_p,
                                     _c_p] + [_c_i64] * 4 + [_c_f, _c_int, _c_p]),
    "fcwdm_groupnorm_bwd_colsum": (_c_int, [_c_p, _c_i64, _c_p, _c_i64, _c_p, _c_p, _c_p, _c_p, _c_p, _c_i64, _c_p, _c_i64,
                                            _c_p, _c_p, _c_p, _c_i64] + [_c_i64] * 4 + [_c_f, _c_int, _c_p]),
    "fcwdm_colsum_scatter": (_c_int, [_c_p, _c_i64, _c_p, _c_i64, _c_p, _c_i64, _c_i64, _c_p]),
    "fcwdm_colsum_cl": (_c_int, [_c_p, _c_i64, _c_p, _c_i64, _c_p, _c_i64, _c_i64, _c_i64, _c_p]),
    "fcwdm_dwt3d_cl_bwd": (_c_int, [_c_p, _c_i64, _c_p, _c_i64, _c_i64, _c_p, _c_i64, _c_p, _c_i64] + [_c_i64] * 5
                           + [_c_f, _c_f, _c_p]),
    "fcwdm_idwt3d_cl_bwd": (_c_int, [_c_p, _c_i64, _c_p, _c_i64, _c_p, _c_i64, _c_p, _c_i64, _c_i64, _c_int] + [_c_i64] * 5
                            + [_c_f, _c_p]),
    "fcwdm_add_cl": (_c_int, [_c_p, _c_i64, _c_p, _c_i64, _c_p, _c_i64, _c_i64, _c_i64, _c_p]),
    "fcwdm_linear_bwd": (_c_int, [_c_p, _c_p, _c_p, _c_i64, _c_p, _c_p, _c_p, _c_i64, _c_i64, _c_i64, _c_int, _c_int, _c_p]),
    "fcwdm_adamw": (_c_int, [_c_p, _c_p, _c_p, _c_p, _c_i64, _c_f, _c_f, _c_f, _c_f, _c_f, _c_i64, _c_f, _c_p]),
    "fcwdm_avgpool2_cl": (_c_int, [_c_p, _c_i64, _c_p, _c_i64] + [_c_i64] * 5 + [_c_int, _c_p]),
    "fcwdm_upsample2_cl": (_c_int, [_c_p, _c_i64, _c_p, _c_i64] + [_c_i64] * 5 + [_c_int, _c_p]),
    "fcwdm_conv3d_pack_all": (_c_int, [_c_p, _c_i64, _c_i64, _c_p]),
    "fcwdm_conv3d_gn_fwd": (_c_int, [_c_p, _c_i64, _c_p, _c_p, _c_p, _c_i64, _c_p, _c_i64, _c_p, _c_i64, _c_p, _c_i64,
                                     _c_p, _c_p, _c_p, _c_i64, _c_f] + [_c_i64] * 6 + [_c_p]),
    "fcwdm_avgpool2_cl_bwd": (_c_int, [_c_p, _c_i64, _c_p, _c_i64, _c_p, _c_i64] + [_c_i64] * 5 + [_c_int, _c_p]),
    "fcwdm_upsample2_cl_bwd": (_c_int, [_c_p, _c_i64, _c_p, _c_i64, _c_p, _c_i64] + [_c_i64] * 5 + [_c_int, _c_p]),
    "fcwdm_clip_normalize_workspace_bytes": (_c_i64, [_c_i64]),
    "fcwdm_clip_normalize": (_c_int, [_c_p, _c_p, _c_p, _c_p, _c_i64] + [_c_i64] * 7 + [ctypes.c_double, ctypes.c_double, _c_p]),
    "fcwdm_debug_set_conv_trace": (_c_int, [_c_p]),
    "fcwdm_conv3d_chain_supported": (_c_int, [_c_i64, _c_i64, _c_int]),
    "fcwdm_conv3d_chain_max_layers": (_c_int, []),
    "fcwdm_conv3d_chain": (_c_int, [ctypes.POINTER(ChainLayer), _c_i64, _c_p, _c_p]),
    "fcwdm_debug_set_chain_trace": (_c_int, [_c_p]),
}

FCWDM_F32, FCWDM_BF16 = 0, 1

_lib = None
_lock = threading.Lock()
_inited_devices = set()
launch_count = 0   # number of C-ABI kernel-launching calls made (bench.py reports it as gpu_launches)


class FcwdmError(RuntimeError):
    pass


def load():
    """Load the shared library (no device needed) and attach prototypes."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise FcwdmError(
                    f"{LIB_PATH} is missing: build it with `python fast-cwdm_b200/fcwdm/build.py` "
                    "(there is no CPU / PyTorch fallback for the fcwdm hot path)")
            lib = ctypes.CDLL(LIB_PATH)
            for name, (res, args) in PROTOTYPES.items():
                if os.environ.get("FCWDM_LIB_PATH") and not hasattr(lib, name):
                    continue                      # A/B runs against an older build: entry points added since are absent
                fn = getattr(lib, name)
                fn.restype = res
                fn.argtypes = args
            _lib = lib
    return _lib


def init(device_index):
    lib = load()
    if device_index not in _inited_devices:
        rc = lib.fcwdm_init(int(device_index))
        if rc != 0:
            raise FcwdmError(f"fcwdm_init({device_index}) failed [{rc}]: {lib.fcwdm_last_error().decode()}")
        _inited_devices.add(device_index)
    return lib


def call(name, *args):
    """Invoke a status-returning entry point; raise FcwdmError with the library's message on failure."""
    global launch_count
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise FcwdmError(f"{name} failed [{rc}]: {lib.fcwdm_last_error().decode()}")
    launch_count += 1
    return rc
