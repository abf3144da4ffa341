"""RepSurf set abstraction (SurfaceAbstraction / SurfaceAbstractionCD = FPS -> ball query -> grouping
-> shared MLP -> max over the group).  CPU: oracle restatement vs the fixture generated from the reference
(tests/golden/make_golden_sa.py).  GPU: the drop-in modules (sm_100a FPS / ball-query / gather / GEMM / BatchNorm
kernels) vs the same fixture.  Sampled coordinates are gathers of the input: exact.  Features: rtol 1e-4, atol 1e-5."""
import os

import numpy as np
import pytest
import torch

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "sa.npz")
CFG = {
    "sa": dict(npoint=64, radius=0.4, nsample=16, in_channel=3 + 3 + 10 + 32, mlp=[32, 64], group_all=False,
               return_polar=True, return_normal=True),
    "sacd": dict(npoint=64, radius=0.4, nsample=16, feat_channel=10 + 32, pos_channel=6, mlp=[32, 64], group_all=False,
                 return_polar=True, return_normal=True),
}


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLDEN)


def _params(orc, gold, name):
    spec = []
    for k, shp in zip(gold[name + "_keys"], gold[name + "_shapes"]):
        k, shp = str(k), str(shp)
        spec.append((k, [int(v) for v in shp.split(",")] if shp else [],
                     "int64" if k.endswith("num_batches_tracked") else "float32"))
    return orc.synthetic_state_dict(spec, seed=4)


@pytest.mark.parametrize("name", ["sa", "sacd"])
def test_oracle_surface_abstraction_matches_reference(orc, gold, name):
    P = {k: v.clone() for k, v in _params(orc, gold, name).items()}
    c, n, f = (torch.from_numpy(gold[k]) for k in ("center", "normal", "feature"))
    for mode in ("train", "eval"):  # the fixture's eval pass follows its train pass (running statistics carry over)
        ctx = orc.Ctx(train=mode == "train", fps_starts=[torch.from_numpy(gold["%s_%s_start" % (name, mode)])])
        nc, nn_, nf = orc.surface_abstraction(P, "", c, n, f, ctx, 64, 0.4, 16, n_layers=2 if name == "sa" else 1,
                                              pos_channel=6 if name == "sacd" else None)
        assert np.array_equal(nc.numpy(), gold["%s_%s_center" % (name, mode)])
        assert np.array_equal(nn_.numpy(), gold["%s_%s_normal" % (name, mode)])
        np.testing.assert_allclose(nf.detach().numpy(), gold["%s_%s_feature" % (name, mode)], rtol=1e-5, atol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["sa", "sacd"])
def test_gpu_surface_abstraction_golden(mpc, orc, gold, name):
    rep = mpc.repsurface_utils
    m = (rep.SurfaceAbstraction if name == "sa" else rep.SurfaceAbstractionCD)(**CFG[name])
    assert list(m.state_dict().keys()) == [str(k) for k in gold[name + "_keys"]]  # reference checkpoint layout
    m.load_state_dict(_params(orc, gold, name))
    m = m.cuda()
    c, n, f = (torch.from_numpy(gold[k]).cuda() for k in ("center", "normal", "feature"))
    for mode in ("train", "eval"):
        m.train(mode == "train")
        with mpc.ops.index_tape(fps_starts=[torch.from_numpy(gold["%s_%s_start" % (name, mode)])]):
            nc, nn_, nf = m(c, n, f)
        assert np.array_equal(nc.cpu().numpy(), gold["%s_%s_center" % (name, mode)])
        assert np.array_equal(nn_.cpu().numpy(), gold["%s_%s_normal" % (name, mode)])
        np.testing.assert_allclose(nf.detach().cpu().numpy(), gold["%s_%s_feature" % (name, mode)], rtol=1e-4, atol=1e-5)


@pytest.mark.gpu
def test_gpu_surface_abstraction_backward_matches_oracle(mpc, orc, gold):
    """Gradients of the shared-MLP weights and of the input feature through gather / GEMM / BatchNorm / max."""
    rep = mpc.repsurface_utils
    P = {k: v.clone().requires_grad_(v.dtype.is_floating_point and "running" not in k)
         for k, v in _params(orc, gold, "sa").items()}
    c, n = torch.from_numpy(gold["center"]), torch.from_numpy(gold["normal"])
    f = torch.from_numpy(gold["feature"]).clone().requires_grad_(True)
    start = torch.from_numpy(gold["sa_train_start"])
    _, _, ref = orc.surface_abstraction(P, "", c, n, f, orc.Ctx(train=True, fps_starts=[start]), 64, 0.4, 16, 2)
    w = torch.randn(ref.shape, generator=torch.Generator().manual_seed(0))
    (ref * w).sum().backward()
    m = rep.SurfaceAbstraction(**CFG["sa"])
    m.load_state_dict(_params(orc, gold, "sa"))
    m = m.cuda().train()
    fg = torch.from_numpy(gold["feature"]).cuda().requires_grad_(True)
    with mpc.ops.index_tape(fps_starts=[start]):
        _, _, out = m(c.cuda(), n.cuda(), fg)
    (out * w.cuda()).sum().backward()
    np.testing.assert_allclose(fg.grad.cpu().numpy(), f.grad.numpy(), rtol=1e-3, atol=1e-5)
    for k, p in m.named_parameters():
        if P[k].grad is not None:
            # a conv bias in front of BatchNorm has an exactly-zero gradient: both sides only hold rounding noise there
            # (oracle ~1e-5, CUDA path ~1e-7), hence the absolute floor
            np.testing.assert_allclose(p.grad.cpu().numpy(), P[k].grad.numpy(), rtol=1e-3, atol=5e-5, err_msg=k)
