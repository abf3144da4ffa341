"""ctypes binding of libmpc_b200.so (the C ABI declared in include/mpc_b200.h).

There is no CPU fallback: if the library is missing, or no CUDA device is present when an op is called,
the call raises.
"""
import ctypes
import functools
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
    "mpc_knn_tc_workspace_bytes": [_i64, _i64, _i64, _i64, _i64, _ptr],
    "mpc_knn_tc_f32": [_ptr, _ptr, _ptr, _ptr, _ptr, _i64, _i64, _i64, _i64, _i64, _i64, _ptr],
    "mpc_knn3_grid_workspace_bytes": [_i64, _i64, _i64, _i64, _ptr],
    "mpc_knn3_grid_f32": [_ptr, _ptr, _ptr, _ptr, _ptr, _i64, _i64, _i64, _i64, _i64, _ptr],
    "mpc_ball_query_f32": [_ptr, _ptr, _ptr, _f32, _i64, _i64, _i64, _i64, _i64, _ptr],
    "mpc_gather_f32": [_ptr, _ptr, _ptr, _i64, _i64, _i64, _i64, _ptr],
    "mpc_gather_bwd_f32": [_ptr, _ptr, _ptr, _i64, _i64, _i64, _i64, _ptr],
    "mpc_gather_i64": [_ptr, _ptr, _ptr, _i64, _i64, _i64, _ptr],
    "mpc_gather_bf16": [_ptr, _ptr, _ptr, _i64, _i64, _i64, _i64, _ptr],
    "mpc_gather_bwd_bf16": [_ptr, _ptr, _ptr, _i64, _i64, _i64, _i64, _ptr],
    "mpc_reduction_scratch_bytes": [_i64],
    "mpc_transition_fwd_f32": [_ptr, _ptr, _ptr, _ptr, _i64, _i64, _i64, _i64, _i64, _ptr],
    "mpc_transition_bwd_f32": [_ptr, _ptr, _ptr, _ptr, _i64, _i64, _i64, _i64, _i64, _ptr],
    "mpc_transition_fwd_csr_f32": [_ptr, _ptr, _ptr, _ptr, _ptr, _i64, _i64, _i64, _i64, _i64, _ptr],
    "mpc_transition_csr_build": [_ptr, _ptr, _i64, _i64, _i64, _i64, _ptr],
    "mpc_transition_csr_apply_f32": [_ptr, _ptr, _ptr, _ptr, _i64, _i64, _i64, _i64, _i64, _ptr],
    "mpc_three_interpolate_fwd_f32": [_ptr, _ptr, _ptr, _ptr, _ptr, _i64, _i64, _i64, _i64, _ptr],
    "mpc_three_interpolate_bwd_f32": [_ptr, _ptr, _ptr, _ptr, _i64, _i64, _i64, _i64, _ptr],
    "mpc_attn_feat_fwd_f32": [_ptr, _i64, _ptr, _ptr, _i64, _ptr, _ptr, _i64, _i64, _i64, _i64, _i64, _ptr],
    "mpc_attn_feat_fwd_bf16": [_ptr, _i64, _ptr, _ptr, _i64, _ptr, _ptr, _i64, _i64, _i64, _i64, _i64, _ptr],
    "mpc_linear_bf16": [_ptr, _i64, _ptr, _i64, _ptr, _ptr, _f32, _ptr, _i64, _ptr, _i64, _i64, _i64, _i64, _i64, _ptr],
    "mpc_f32_to_bf16": [_ptr, _i64, _ptr, _i64, _i64, _i64, _ptr],
    "mpc_attn_feat_bwd_f32": [_ptr, _ptr, _i64, _ptr, _ptr, _i64, _ptr, _ptr, _i64, _ptr, _ptr, _i64, _ptr,
                              _i64, _i64, _i64, _i64, _i64, _ptr],
    "mpc_attn_xyz_fwd_f32": [_ptr] * 13 + [_i64] * 6 + [_ptr],
    "mpc_attn_xyz_bwd_f32": [_ptr] * 21 + [_i64] * 6 + [_ptr],
    "mpc_bn_stats_f32": [_ptr, _ptr, _ptr, _ptr, _ptr, _f32, _ptr, _i64, _i64, _ptr],
    "mpc_col_sum_f32": [_ptr, _ptr, _ptr, _i64, _i64, _ptr],
    "mpc_bn_act_fwd_sums_f32": [_ptr, _ptr, _ptr, _ptr, _f32, _f32, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _f32, _i64,
                                _i64, _ptr],
    "mpc_bn_act_fwd_f32": [_ptr, _ptr, _ptr, _ptr, _ptr, _f32, _f32, _ptr, _ptr, _i64, _i64, _ptr],
    "mpc_bn_act_bwd_f32": [_ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _f32, _f32, _int, _ptr, _ptr, _ptr, _ptr, _ptr, _i64,
                           _i64, _i64, _i64, _ptr],
    "mpc_linear_fwd_f32": [_ptr, _i64, _ptr, _i64, _ptr, _ptr, _i64, _ptr, _ptr, _i64, _i64, _i64, _i64, _ptr],
    "mpc_linear_affine_act_f32": [_ptr, _i64, _ptr, _i64, _ptr, _ptr, _f32, _ptr, _i64, _ptr, _i64, _i64, _i64, _i64, _ptr],
    "mpc_bn_finalize_f32": [_ptr, _ptr, _ptr, _ptr, _ptr, _f32, _i64, _i64, _ptr],
    "mpc_linear_wgrad_f32": [_ptr, _i64, _ptr, _i64, _ptr, _i64, _i64, _i64, _i64, _i64, _ptr],
    "mpc_linear_dgrad_f32": [_ptr, _i64, _ptr, _i64, _ptr, _i64, _i64, _i64, _i64, _ptr, _i64, _ptr],
    "mpc_smooth_ce_fwd_f32": [_ptr, _i64, _ptr, _f32, _ptr, _ptr, _ptr, _i64, _i64, _ptr],
    "mpc_smooth_ce_bwd_f32": [_ptr, _i64, _ptr, _f32, _ptr, _ptr, _ptr, _i64, _i64, _i64, _ptr],
    "mpc_umbrella_features_f32": [_ptr, _ptr, _i64, _ptr, _ptr, _i64, _i64, _i64, _i64, _ptr],
    "mpc_debug_trace_buffer": [_ptr],
    "mpc_debug_set_knob": [_int, _i64],
}

# kernels each entry point enqueues (cudaMemsetAsync calls not counted)
KERNELS_PER_CALL = {
    "mpc_fps_f32": 1, "mpc_knn_f32": 1, "mpc_knn_tc_workspace_bytes": 0, "mpc_knn_tc_f32": 4, "mpc_knn3_grid_workspace_bytes": 0, "mpc_knn3_grid_f32": 5, "mpc_ball_query_f32": 1, "mpc_gather_f32": 1, "mpc_gather_bwd_f32": 1,
    "mpc_gather_i64": 1, "mpc_gather_bf16": 1, "mpc_gather_bwd_bf16": 1, "mpc_reduction_scratch_bytes": 0, "mpc_transition_fwd_f32": 3, "mpc_transition_bwd_f32": 1, "mpc_transition_fwd_csr_f32": 4, "mpc_transition_csr_build": 3, "mpc_transition_csr_apply_f32": 1,
    "mpc_three_interpolate_fwd_f32": 2, "mpc_three_interpolate_bwd_f32": 1, "mpc_attn_feat_fwd_f32": 1, "mpc_attn_feat_fwd_bf16": 1, "mpc_linear_bf16": 1, "mpc_f32_to_bf16": 1,
    "mpc_attn_feat_bwd_f32": 1, "mpc_attn_xyz_fwd_f32": 1, "mpc_attn_xyz_bwd_f32": 1, "mpc_bn_stats_f32": 1, "mpc_col_sum_f32": 1, "mpc_bn_finalize_f32": 1, "mpc_bn_act_fwd_sums_f32": 1,
    "mpc_bn_act_fwd_f32": 1, "mpc_bn_act_bwd_f32": 2, "mpc_linear_fwd_f32": 1, "mpc_linear_affine_act_f32": 1,
    "mpc_linear_wgrad_f32": 1,
    "mpc_linear_dgrad_f32": 1,
    "mpc_smooth_ce_fwd_f32": 1, "mpc_smooth_ce_bwd_f32": 1,
    "mpc_umbrella_features_f32": 1,
    "mpc_debug_trace_buffer": 0,
    "mpc_debug_set_knob": 0,
}

_lib = None
launch_count = 0   # C-ABI calls that enqueued GPU work
kernel_count = 0   # kernels of ours those calls launched (bench.py reports it as gpu_launches)
# Optional per-op device timing for bench.py: {"names": set or None (= all), "records": {name: [(ev0, ev1, bytes)]}}
# With "calls": [] instead of "records", the matching calls are only logged as (name, args, bytes) so that bench.py
# can re-issue exactly the same launches back to back inside a CUDA graph (pure device time, no host gaps).
profiler = None
# Called (if set) when an entry point fails, before MpcError is raised: ops.py re-zeroes the pooled reduction scratch,
# because a failure between a producer and its consumer would otherwise leave partial sums behind (scratch contract).
on_error = None


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
        # debug / tuning: MPC_KNOBS="id=value,..." applies mpc_debug_set_knob at load (A/B runs of bench.py)
        for kv in filter(None, os.environ.get("MPC_KNOBS", "").split(",")):
            k, v = kv.split("=")
            lib.mpc_debug_set_knob(int(k), int(v))
    return _lib


def stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def call(name, *args, algo_bytes=0):
    """Call an entry point on the current torch CUDA stream and raise on a non-zero status.  algo_bytes is the
    op's algorithmic HBM traffic (SURVEY.md 8d formulas), recorded only when bench.py's profiler is on."""
    global launch_count, kernel_count
    lib = load()
    prof = profiler
    timed = prof is not None and (prof["names"] is None or name in prof["names"])
    if timed and "calls" in prof:
        prof["calls"].append((name, args, algo_bytes))
        timed = False
    if timed:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        if prof.get("keep_args"):
            prof.setdefault("args", {}).setdefault(name, []).append(args)
    rc = getattr(lib, name)(*args, stream_ptr())
    if rc != 0:
        kind = {-1: "invalid arguments", -2: "unsupported shape"}.get(rc, "CUDA error %d" % rc)
        if on_error is not None and rc < 0:  # (after a CUDA error the context may be unusable: do not touch it)
            try:
                on_error()
            except Exception:  # noqa: BLE001 -- the original failure is the one to report
                pass
        raise MpcError("%s failed: %s" % (name, kind))
    if timed:
        ev1.record()
        prof["records"].setdefault(name, []).append((ev0, ev1, algo_bytes))
    launch_count += 1
    kernel_count += KERNELS_PER_CALL[name]


def replay(calls):
    """Re-issue logged calls on the current stream (used under CUDA-graph capture by bench.py)."""
    lib = load()
    for name, args, _ in calls:
        rc = getattr(lib, name)(*args, stream_ptr())
        if rc != 0:
            raise MpcError("%s failed on replay: %d" % (name, rc))


def require_cuda(*tensors):
    """Every operand on a CUDA device, and all of them on the same one."""
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise MpcError("mpc_b200 ops run on a CUDA device only (got a %s tensor); there is no CPU fallback"
                           % t.device.type)
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise MpcError("mpc_b200 ops need all operands on one device (got %s and %s)" % (dev, t.device))


def _first_cuda_tensor(args, kwargs):
    for a in args:
        if isinstance(a, torch.Tensor) and a.is_cuda:
            return a
    for a in kwargs.values():
        if isinstance(a, torch.Tensor) and a.is_cuda:
            return a
    return None


def on_tensor_device(fn):
    """Run `fn` with the CUDA device of its first CUDA tensor argument current.  Kernels are launched on torch's
    current stream, which belongs to torch's current device: a model on cuda:1 called while cuda:0 is current would
    otherwise be launched on device 0 with device-1 pointers.  A no-op (one comparison) when the devices agree."""
    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        t = _first_cuda_tensor(args, kwargs)
        if t is None or t.device.index == torch.cuda.current_device():
            return fn(*args, **kwargs)
        with torch.cuda.device(t.device):
            return fn(*args, **kwargs)
    return wrapper
