"""markov-process-analysis-on-point-cloud_b200: the B200 (sm_100a) point-set hot path of
ssr0512/Markov-Process-Analysis-on-Point-Cloud behind the reference's own function / nn.Module interface.

The directory name is not a Python identifier; import it with
    mpc = importlib.import_module("markov-process-analysis-on-point-cloud_b200")
Sub-modules: ops (free functions + autograd Functions), pointnet2_utils / repsurface_utils (drop-in mirrors of
the reference's two module files), task_models (classifier / part-seg callers), dist (batch sharding + gradient all-reduce), harness (vote-evaluation loops, checkpoints), _lib (C-ABI binding),
build (nvcc recipe).
"""
from . import _lib, build, dist, harness, ops, pointnet2_utils, repsurface_utils, task_models  # noqa: F401

__version__ = "0.1.0"
