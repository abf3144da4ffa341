"""ctypes binding of libmpc_b200.so (the C ABI declared in include/mpc_b200.h).

There is no CPU fallback: if the library is missing, or no CUDA device is present when an op is called,
the call raises.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libmpc_b200.so")

_i64, _f32, _ptr, _int = ctypes.c_int64, ctypes.c_float, ctypes.c_void_p, ctypes.c_int

# name -> argument ctypes (every entry point of include/mpc_b200.h; tests check the two lists agree)
SIGNATURES = {
    "mpc_version": [],
    "mpc_compiled_arch": [],
    "mpc_fps_f32": [_ptr, _ptr, _ptr, _i64, _i64, _i64, _i64, _ptr],
    "mpc_knn_f32": [_ptr, _ptr, _ptr, _ptr, _i64, _i64, _i64, _i64, _i64, _ptr],
    "mpc_ball_query_f32": [_ptr, _ptr, _ptr, _f32, _i64, _i64, _i64, _i64, _i64, _ptr],
    "mpc_gather_f32": [_ptr, _ptr, _ptr, _i64, _i64, _i64, _i64, _ptr],
    "mpc_gather_bwd_f32": [_ptr, _ptr, _ptr, _i64, _i64, _i64, _i64, _ptr],
    "mpc_gather_i64": [_ptr, _ptr, _ptr, _i64, _i64, _i64, _ptr],
    "mpc_transition_fwd_f32": [_ptr, _ptr, _ptr, _ptr, _i64, _i64, _i64, _i64, _i64, _ptr],
    "mpc_transition_bwd_f32": [_ptr, _ptr, _ptr, _ptr, _i64, _i64, _i64, _i64, _i64, _ptr],
    "mpc_three_interpolate_fwd_f32": [_ptr, _ptr, _ptr, _ptr, _ptr, _i64, _i64, _i64, _i64, _ptr],
    "mpc_three_interpolate_bwd_f32": [_ptr, _ptr, _ptr, _ptr, _i64, _i64, _i64, _i64, _ptr],
    "mpc_attn_feat_fwd_f32": [_ptr, _i64, _ptr, _ptr, _i64, _ptr, _ptr, _i64, _i64, _i64, _i64, _i64, _ptr],
    "mpc_attn_feat_bwd_f32": [_ptr, _ptr, _i64, _ptr, _ptr, _i64, _ptr, _ptr, _i64, _ptr, _ptr, _i64,
                              _i64, _i64, _i64, _i64, _i64, _ptr],
    "mpc_attn_xyz_fwd_f32": [_ptr] * 10 + [_i64] * 6 + [_ptr],
    "mpc_attn_xyz_bwd_f32": [_ptr] * 17 + [_i64] * 6 + [_ptr],
    "mpc_bn_stats_f32": [_ptr, _ptr, _ptr, _ptr, _ptr, _f32, _ptr, _i64, _i64, _ptr],
    "mpc_bn_act_fwd_f32": [_ptr, _ptr, _ptr, _ptr, _ptr, _f32, _f32, _ptr, _i64, _i64, _ptr],
    "mpc_bn_act_bwd_f32": [_ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _f32, _f32, _int, _ptr, _ptr, _ptr, _ptr,
                           _i64, _i64, _ptr],
}

_lib = None
launch_count = 0  # number of C-ABI calls that enqueued GPU work (bench.py reports it)


class MpcError(RuntimeError):
    pass


def load():
    """Load the shared library (no GPU needed to load it or to resolve its symbols)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MpcError("libmpc_b200.so is not built (%s). Run `python __graft_entry__.py build`; "
                           "there is no CPU fallback." % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, args in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.argtypes = args
            fn.restype = ctypes.c_int
        _lib = lib
    return _lib


def stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def call(name, *args):
    """Call an entry point on the current torch CUDA stream and raise on a non-zero status."""
    global launch_count
    lib = load()
    rc = getattr(lib, name)(*args, stream_ptr())
    if rc != 0:
        kind = {-1: "invalid arguments", -2: "unsupported shape"}.get(rc, "CUDA error %d" % rc)
        raise MpcError("%s failed: %s" % (name, kind))
    launch_count += 1


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise MpcError("mpc_b200 ops run on a CUDA device only (got a %s tensor); there is no CPU fallback"
                           % t.device.type)
