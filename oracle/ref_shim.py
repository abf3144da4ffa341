"""Loader for the REAL reference (test infrastructure only, this container only).

/root/reference does not import as shipped: modules/pointnet2_utils.py:8-9
imports `models.polar_utils` / `models.recons_utils` (they live in modules/), modules/repsurface_utils.py:5
imports a `query_knn_point` that is defined nowhere, matplotlib is imported and unused, and
`upsample` hard-codes torch.cuda.FloatTensor (modules/pointnet2_utils.py:36).  This file applies the
minimal import aliases so the unmodified reference source executes on CPU.  Nothing is copied from
the reference; it is imported from where it lies.  Only `tests/golden/make_golden.py` and the
optional `tests/test_oracle_vs_reference.py` (skipped when /root/reference is absent, e.g. on the
GPU box) use this loader.  The product package never imports it.
"""
import contextlib
import os
import sys
import types

REF_ROOT = os.environ.get(
    "MPC_REFERENCE_ROOT", "/root/reference/Markov_Process_Analysis_on_Point_Cloud")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "modules"))


_loaded = {}


def load():
    """Return a namespace with the reference's modules: .pn2 (modules.pointnet2_utils),
    .rep (modules.repsurface_utils), .cls (models.repsurf.repsurf_ssg_umb),
    .seg (models.repsurf.pointnet2_part_seg_msg)."""
    if _loaded:
        return types.SimpleNamespace(**_loaded)
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    import importlib
    import torch

    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    # (2) matplotlib is imported but never used on the hot path.
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        mpl.pyplot = types.ModuleType("matplotlib.pyplot")
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = mpl.pyplot
    # (3) break the modules.pointnet2_utils <-> modules.recons_utils import cycle with a placeholder.
    ph = types.ModuleType("modules.pointnet2_utils")
    ph.query_knn_point = None
    ph.index_points = None
    sys.modules["modules.pointnet2_utils"] = ph
    polar = importlib.import_module("modules.polar_utils")
    recons = importlib.import_module("modules.recons_utils")
    del sys.modules["modules.pointnet2_utils"]
    # (4) the reference imports these under the wrong package name.
    sys.modules["models.polar_utils"] = polar
    sys.modules["models.recons_utils"] = recons
    pn2 = importlib.import_module("modules.pointnet2_utils")
    sys.modules["models.pointnet2_utils"] = pn2

    # (5) `query_knn_point` is referenced but defined nowhere; call sites use it as "indices of knn_point".
    def query_knn_point(k, xyz, new_xyz, cuda=False):
        return pn2.knn_point(k, xyz, new_xyz)[1]

    pn2.query_knn_point = query_knn_point
    rep = importlib.import_module("modules.repsurface_utils")
    cls = importlib.import_module("models.repsurf.repsurf_ssg_umb")
    seg = importlib.import_module("models.repsurf.pointnet2_part_seg_msg")
    _loaded.update(pn2=pn2, rep=rep, cls=cls, seg=seg, torch=torch)
    return types.SimpleNamespace(**_loaded)


@contextlib.contextmanager
def cpu_float_tensor_redirect():
    """(6) `upsample` allocates with torch.cuda.FloatTensor (modules/pointnet2_utils.py:36); on a
    CPU-only host redirect that constructor to the CPU one for the duration of a reference call."""
    import torch

    saved = torch.cuda.FloatTensor
    torch.cuda.FloatTensor = lambda *shape: torch.FloatTensor(*shape)
    try:
        yield
    finally:
        torch.cuda.FloatTensor = saved
