"""CPU-side checks of the drop-in boundary: the C-ABI library loads without a GPU and exports exactly the
symbols include/mpc_b200.h declares; the Python binding table agrees; the product fails loudly (no CPU
fallback) and never imports the oracle."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
HEADER = os.path.join(ROOT, "include", "mpc_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\bint\s+(mpc_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_expected_entry_points():
    syms = declared_symbols()
    for must in ("mpc_fps_f32", "mpc_knn_f32", "mpc_ball_query_f32", "mpc_gather_f32", "mpc_gather_bwd_f32",
                 "mpc_transition_fwd_f32", "mpc_transition_bwd_f32", "mpc_three_interpolate_fwd_f32",
                 "mpc_attn_feat_fwd_f32", "mpc_attn_xyz_fwd_f32", "mpc_bn_act_fwd_f32"):
        assert must in syms


def test_library_loads_and_exports_every_declared_symbol(mpc):
    mpc.build.build()
    lib = ctypes.CDLL(mpc._lib.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), name
    assert lib.mpc_version() >= 100
    assert lib.mpc_compiled_arch() == 1000
    out = subprocess.run(["nm", "-D", "--defined-only", mpc._lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r"\bT (mpc_[a-z0-9_]+)", out)))
    assert exported == declared_symbols()  # nothing undeclared leaks out, nothing declared is missing


def test_binding_table_matches_header(mpc):
    assert sorted(mpc._lib.SIGNATURES) == declared_symbols()
    text = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    for name, args in mpc._lib.SIGNATURES.items():
        proto = re.search(r"\bint\s+%s\s*\((.*?)\)\s*;" % name, text, flags=re.S).group(1)
        n_params = 0 if proto.strip() == "void" else proto.count(",") + 1
        n_bound = len(args)  # includes the trailing stream argument where the entry point takes one
        assert n_bound == n_params, (name, n_bound, n_params)


def test_sass_is_sm100a(mpc):
    out = subprocess.run(["cuobjdump", "-lelf", mpc._lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_cpu_fallback(mpc):
    xyz = torch.rand(1, 16, 3)
    with pytest.raises(mpc._lib.MpcError):
        mpc.ops.farthest_point_sample(xyz, 4)
    with pytest.raises(mpc._lib.MpcError):
        mpc.ops.knn_point(4, xyz, xyz)
    with pytest.raises(mpc._lib.MpcError):
        mpc.ops.index_points(xyz, torch.zeros(1, 2, dtype=torch.long))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "markov-process-analysis-on-point-cloud_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "liboracle" not in src, f
    code = ("import importlib,sys; sys.path.insert(0,%r); "
            "importlib.import_module('markov-process-analysis-on-point-cloud_b200'); "
            "assert not any(m.startswith('oracle') for m in sys.modules)" % ROOT)
    subprocess.check_call([sys.executable, "-c", code])


def test_state_dict_layout_matches_reference(mpc, golden_specs):
    import argparse

    m = mpc.task_models.Model(argparse.Namespace(num_point=1024, return_dist=True, cuda_ops=False, num_class=40))
    assert [(k, list(v.shape)) for k, v in m.state_dict().items()] == [(k, s) for k, s, _ in golden_specs["cls"]]
    s = mpc.task_models.get_model(50)
    assert [(k, list(v.shape)) for k, v in s.state_dict().items()] == [(k, s_) for k, s_, _ in golden_specs["seg"]]
