"""CPU checks of the evaluation harness and of INTEGRATION.md Option A: metric bookkeeping of the
vote-evaluation loops, the reference checkpoint format, and -- when the reference tree is present (build container
only; skipped on the GPU box) -- that the reference's UNMODIFIED model files construct on the drop-in modules and that
a checkpoint written from the reference's own model class loads into the drop-in."""
import argparse
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
REF = "/root/reference/Markov_Process_Analysis_on_Point_Cloud"
has_ref = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "modules")), reason="reference tree not present")


def test_classification_accuracy_bookkeeping(mpc):
    pred = torch.tensor([[0.1, 0.9, 0.0], [0.8, 0.1, 0.1], [0.2, 0.3, 0.5], [0.6, 0.3, 0.1]])
    target = torch.tensor([1, 0, 1, 2])
    acc, cls = mpc.harness.classification_accuracy(pred, target, 3)
    assert acc == 0.5
    np.testing.assert_allclose(cls, [[1.0, 1], [0.5, 1], [0.0, 1]])


def test_segmentation_metrics_follow_the_reference(mpc):
    h = mpc.harness
    B, N = 2, 6
    target = torch.tensor([[12, 12, 13, 14, 15, 12], [28, 29, 29, 28, 28, 29]])  # Chair, Laptop
    logits = torch.full((B, N, 50), -5.0)
    for i in range(B):
        for n in range(N):
            logits[i, n, int(target[i, n])] = 5.0
    logits[0, 0, 13] = 9.0  # one wrong point in the chair
    pred, correct, seen, ious = h.segmentation_metrics(logits, target)
    assert seen == 12
    # the reference takes the arg-max INSIDE the category's part range and never adds the range offset back
    assert pred[1].tolist() == [0, 1, 1, 0, 0, 1] and pred[0].tolist() == [1, 0, 1, 2, 3, 0]
    assert len(ious["Chair"]) == 1 and len(ious["Laptop"]) == 1 and len(ious["Table"]) == 0
    assert h.SEG_LABEL_TO_CAT[0] == "Airplane" and h.SEG_LABEL_TO_CAT[49] == "Table" and len(h.SEG_LABEL_TO_CAT) == 50


def test_pointcloud_scale_draw_order(mpc):
    np.random.seed(3)
    pc = torch.ones(3, 5, 6)
    out = mpc.harness.PointcloudScale(0.95, 1.05)(pc.clone())
    np.random.seed(3)
    expect = np.stack([np.random.uniform(0.95, 1.05, size=[3]) for _ in range(3)]).astype(np.float32)
    np.testing.assert_allclose(out[:, 0, :3].numpy(), expect, rtol=1e-6)
    assert torch.equal(out[:, :, 3:], pc[:, :, 3:])  # only the coordinates are scaled


def test_checkpoint_roundtrip_reference_format(mpc, tmp_path):
    m = mpc.task_models.get_model(50)
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)
    path = str(tmp_path / "best_model.pth")
    mpc.harness.save_checkpoint(path, m, opt, epoch=3, inctance_avg_iou=0.5)
    ck = torch.load(path, weights_only=False)
    assert set(ck) >= {"epoch", "inctance_avg_iou", "model_state_dict", "optimizer_state_dict"}
    m2 = mpc.task_models.get_model(50)
    mpc.harness.load_checkpoint(path, m2)
    for (k, a), (_, b) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert torch.equal(a, b), k


OPTION_A = r"""
import importlib, sys, argparse, torch
sys.path.insert(0, %(root)r); sys.path.insert(0, %(ref)r)
import types
mpl = types.ModuleType("matplotlib"); mpl.pyplot = types.ModuleType("matplotlib.pyplot")
sys.modules.setdefault("matplotlib", mpl); sys.modules.setdefault("matplotlib.pyplot", mpl.pyplot)
mpc = importlib.import_module("markov-process-analysis-on-point-cloud_b200")
sys.modules["modules.repsurface_utils"] = mpc.repsurface_utils
sys.modules["models.pointnet2_utils"] = mpc.pointnet2_utils
sys.modules["modules.pointnet2_utils"] = mpc.pointnet2_utils
from models.repsurf.repsurf_ssg_umb import Model                      # reference files, unchanged
from models.repsurf.pointnet2_part_seg_msg import get_model, get_loss
import json
specs = json.load(open(%(specs)r))
m = Model(argparse.Namespace(num_point=1024, return_dist=True, cuda_ops=False, num_class=40))
assert type(m.keepHigh).__module__.startswith("markov-process"), type(m.keepHigh).__module__
assert [(k, list(v.shape)) for k, v in m.state_dict().items()] == [(k, s) for k, s, _ in specs["cls"]]
s = get_model(50)
assert type(s.keepHigh).__module__.startswith("markov-process")
assert [(k, list(v.shape)) for k, v in s.state_dict().items()] == [(k, s_) for k, s_, _ in specs["seg"]]
print("option-a ok", len(m.state_dict()), len(s.state_dict()))
"""


@has_ref
def test_integration_option_a_reference_models_on_dropin_modules():
    """INTEGRATION.md Option A: alias the two module names, import the reference's model files unchanged."""
    code = OPTION_A % {"root": ROOT, "ref": REF, "specs": os.path.join(ROOT, "tests", "golden", "specs.json")}
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "option-a ok 743 2189" in out.stdout


@has_ref
def test_reference_checkpoint_loads_into_dropin(mpc, tmp_path):
    """A checkpoint written from the REFERENCE's own model classes (its training scripts' dictionary format,
    R/tool/train_partseg.py:297-306) loads into the drop-in models with strict key / shape matching."""
    sys.path.insert(0, ROOT)
    from oracle import ref_shim

    r = ref_shim.load()
    torch.manual_seed(0)
    ref_seg = r.seg.get_model(50)
    ref_cls = r.cls.Model(argparse.Namespace(num_point=1024, return_dist=True, cuda_ops=False, num_class=40))
    for name, ref_model, ours in (("seg", ref_seg, mpc.task_models.get_model(50)),
                                  ("cls", ref_cls, mpc.task_models.Model(argparse.Namespace(
                                      num_point=1024, return_dist=True, cuda_ops=False, num_class=40)))):
        path = str(tmp_path / ("%s_best_model.pth" % name))
        torch.save({"epoch": 1, "model_state_dict": ref_model.state_dict()}, path)
        mpc.harness.load_checkpoint(path, ours, strict=True)
        for (k, a), (k2, b) in zip(ref_model.state_dict().items(), ours.state_dict().items()):
            assert k == k2 and torch.equal(a, b), k
