"""Generate the golden fixtures by executing the UNMODIFIED reference (this container only).

    python tests/golden/make_golden.py

Imports the reference from /root/reference through oracle/ref_shim.py (import aliases only -- no
reference source is edited or copied), runs each hot-path function / module / model on seeded synthetic
inputs on CPU, and writes small .npz fixtures next to this script.  The fixtures travel with the repo
(the reference tree does not exist on the GPU box); tests/test_oracle_golden.py checks the CPU oracle
against them and the `-m gpu` tests check the CUDA path against the same files.

Weights are a deterministic function of (state_dict key, shape) -- oracle.markov_oracle.
synthetic_state_dict -- so no checkpoint needs to be stored: the fixture keeps the key/shape list
(which also pins state_dict compatibility: 743 keys for the classifier, 2189 for part-seg).
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
from oracle import markov_oracle as orc  # noqa: E402
from oracle import ref_shim  # noqa: E402


def spec_of(module):
    return [(k, list(v.shape), str(v.dtype).replace("torch.", "")) for k, v in module.state_dict().items()]


def load_synth(module, seed=0):
    spec = spec_of(module)
    module.load_state_dict(orc.synthetic_state_dict(spec, seed=seed))
    return spec


def np_(t):
    return t.detach().cpu().numpy()


def cloud(B, N, C=3):
    """The reference's own smoke input: rand*2-1 (R/modules/recons_utils.py:232)."""
    return torch.rand(B, N, C) * 2 - 1


class Recorder:
    """Wraps the reference's farthest_point_sample / knn_point inside a reference module namespace to
    record every index tensor in call order (the `index tape`)."""

    def __init__(self, ns):
        self.ns = ns
        self.tape = []
        self._fps = ns.farthest_point_sample
        self._knn = ns.knn_point

    def __enter__(self):
        def fps(xyz, npoint):
            out = self._fps(xyz, npoint)
            self.tape.append(("fps", out.clone()))
            return out

        def knn(k, xyz, new_xyz):
            d, i = self._knn(k, xyz, new_xyz)
            self.tape.append(("knn", i.clone()))
            return d, i

        self.ns.farthest_point_sample = fps
        self.ns.knn_point = knn
        return self

    def __exit__(self, *a):
        self.ns.farthest_point_sample = self._fps
        self.ns.knn_point = self._knn

    def arrays(self, prefix="tape"):
        out = {}
        for n, (kind, t) in enumerate(self.tape):
            out["%s_%03d_%s" % (prefix, n, kind)] = np_(t).astype(np.int16)  # every index < 32768
        return out


def ops_fixture(r):
    pn2 = r.pn2
    out = {}
    torch.manual_seed(1)
    # --- FPS: tie-free random cloud, a duplicated-points cloud (all-duplicate tail -> index 0), C=5
    xyz = cloud(3, 257)
    torch.manual_seed(2)
    out["fps_xyz"] = np_(xyz)
    out["fps_idx"] = np_(pn2.farthest_point_sample(xyz, 64))
    dup = cloud(2, 40)
    dup[:, 20:] = dup[:, :20]
    torch.manual_seed(3)
    out["fpsdup_xyz"] = np_(dup)
    out["fpsdup_idx"] = np_(pn2.farthest_point_sample(dup, 32))
    feat = cloud(2, 96, 5)
    torch.manual_seed(4)
    out["fpsc5_xyz"] = np_(feat)
    out["fpsc5_idx"] = np_(pn2.farthest_point_sample(feat, 24))
    # --- kNN / square_distance (C = 3 and a feature-space C = 16)
    torch.manual_seed(5)
    ref, qry = cloud(2, 300), cloud(2, 77)
    d, i = pn2.knn_point(8, ref, qry)
    out.update(knn3_ref=np_(ref), knn3_qry=np_(qry), knn3_dist=np_(d), knn3_idx=np_(i),
               sqdist3=np_(pn2.square_distance(qry, ref)))
    ref, qry = cloud(2, 128, 16), cloud(2, 40, 16)
    d, i = pn2.knn_point(8, ref, qry)
    out.update(knn16_ref=np_(ref), knn16_qry=np_(qry), knn16_dist=np_(d), knn16_idx=np_(i))
    # self-kNN (query set == reference set): self-distance is +-4.8e-7, not 0, in expanded form
    ref = cloud(1, 200)
    d, i = pn2.knn_point(8, ref, ref)
    out.update(knnself_ref=np_(ref), knnself_dist=np_(d), knnself_idx=np_(i))
    # --- ball query (incl. a query with no hit -> N)
    ref, qry = cloud(2, 256), cloud(2, 33)
    qry[0, 0] = 5.0
    out.update(ball_ref=np_(ref), ball_qry=np_(qry), ball_idx=np_(pn2.query_ball_point(0.35, 16, ref, qry)))
    # --- index_points, rank 2 and rank 3
    pts = torch.randn(2, 50, 7)
    i2 = torch.randint(0, 50, (2, 13))
    i3 = torch.randint(0, 50, (2, 13, 4))
    out.update(gather_pts=np_(pts), gather_i2=np_(i2), gather_o2=np_(pn2.index_points(pts, i2)),
               gather_i3=np_(i3), gather_o3=np_(pn2.index_points(pts, i3)))
    # --- transition (`upsample`), dense reference; unreached rows, a zero in channel 0, ratio 4
    with ref_shim.cpu_float_tensor_redirect():
        for name, (B, S, C, K, ratio) in dict(a=(2, 16, 8, 4, 2), b=(2, 12, 5, 3, 4), c=(1, 64, 64, 8, 2)).items():
            p = torch.randn(B, S, C)
            p[0, 1, 0] = 0.0  # counted out of the denominator by count_nonzero (:44)
            small = cloud(B, S)
            big = cloud(B, S * ratio)
            _, kidx = pn2.knn_point(K, big, small)
            p.requires_grad_(True)
            o = pn2.upsample(p, kidx, scale_ratio=ratio)
            w = torch.randn_like(o)
            (o * w).sum().backward()
            out.update({"tr%s_points" % name: np_(p), "tr%s_idx" % name: np_(kidx), "tr%s_out" % name: np_(o),
                        "tr%s_w" % name: np_(w), "tr%s_grad" % name: np_(p.grad),
                        "tr%s_ratio" % name: np.int64(ratio)})
    return out


def blocks_fixture(r):
    pn2, rep = r.pn2, r.rep
    out, specs = {}, {}
    torch.manual_seed(11)
    # --- Linear (bn=False => BatchNorm1d) train + eval
    lin = pn2.Linear(12, 20, bn=False)
    specs["linear"] = load_synth(lin)
    x = torch.randn(3, 17, 12)
    lin.train()
    y = lin(x)
    out.update(linear_x=np_(x), linear_train=np_(y), linear_rm=np_(lin.norm2.running_mean), linear_rv=np_(lin.norm2.running_var))
    lin.eval()
    out["linear_eval"] = np_(lin(x))
    # --- LocalTrans: xyz=True with FPS, features with FPS, features without FPS (decoder), each fwd+bwd (train)
    B, N, S, K = 2, 96, 48, 8
    xyz = cloud(B, N)
    fpsi = torch.stack([torch.randperm(N)[:S] for _ in range(B)])
    sub = pn2.index_points(xyz, fpsi)
    _, idx = pn2.knn_point(K, xyz, sub)
    _, idx_self = pn2.knn_point(K, xyz, xyz)
    feat = torch.randn(B, N, 16)
    out.update(lt_xyz=np_(xyz), lt_fps=np_(fpsi), lt_idx=np_(idx), lt_idx_self=np_(idx_self), lt_feat=np_(feat))
    for name, (in_c, out_c, res, use_xyz, use_fps) in dict(
            xyz=(3, 32, True, True, True), feat=(16, 32, True, False, True),
            dec=(16, 16, False, False, False), xyz0=(3, 64, True, True, False)).items():
        m = pn2.LocalTrans(in_c, out_c, K, residual=res)
        specs["lt_" + name] = load_synth(m)
        m.train()
        f = (xyz if use_xyz else feat).clone().requires_grad_(True)
        y = m(features=f, idx=idx if use_fps else idx_self, pos=xyz, FPS_idx=fpsi if use_fps else None, xyz=use_xyz)
        w = torch.randn_like(y)
        (y * w).sum().backward()
        out.update({"lt_%s_out" % name: np_(y), "lt_%s_w" % name: np_(w), "lt_%s_gin" % name: np_(f.grad)})
        for k, p in m.named_parameters():
            if p.grad is not None:
                out["lt_%s_g.%s" % (name, k)] = np_(p.grad)
    # --- LocalMerge (seg 3-branch and cls 2-branch), eval, with the index tape
    for variant, ns in (("seg", pn2), ("cls", rep)):
        m = ns.LocalMerge(16, 32, K, residual=True)
        specs["lm_" + variant] = load_synth(m)
        m.eval()
        with Recorder(ns) as rec:
            y, _, i, d = m(xyz=sub, base_xyz=xyz, normal=xyz, feature=feat, FPS_idx=fpsi)
        out.update({"lm_%s_out" % variant: np_(y), "lm_%s_idx" % variant: np_(i), "lm_%s_dist" % variant: np_(d)})
        out.update(rec.arrays("lm_%s_tape" % variant))
    # --- PointNetFeaturePropagation (three_nn + three_interpolate + Linear), train fwd+bwd
    fp = pn2.PointNetFeaturePropagation(16, [24], act=True)
    specs["fp"] = load_synth(fp)
    fp.train()
    p2 = torch.randn(B, S, 16, requires_grad=True)
    y = fp(xyz, sub, None, p2)
    w = torch.randn_like(y)
    (y * w).sum().backward()
    out.update(fp_out=np_(y), fp_w=np_(w), fp_p2=np_(p2), fp_gp2=np_(p2.grad))
    return out, specs


def grads_summary(module, prefix):
    out = {}
    for k, p in module.named_parameters():
        if p.grad is None:
            continue
        g = p.grad.detach().flatten()
        out["%s.%s" % (prefix, k)] = np.concatenate([[float(g.norm())], np_(g[:15])]).astype(np.float32)
    return out


def models_fixture(r):
    out, specs = {}, {}
    # ---------------- classifier (config 1 shape, reduced batch): eval fwd, train fwd+bwd
    args = argparse.Namespace(num_point=1024, return_dist=True, cuda_ops=False, num_class=40)
    torch.manual_seed(0)
    m = r.cls.Model(args)
    specs["cls"] = load_synth(m)
    m.drop1.p = m.drop2.p = 0.0
    torch.manual_seed(1)
    pts = torch.rand(6, 3, 1024) * 2 - 1  # eval uses the first 2 clouds, train all 6 (BatchNorm over
    out["cls_points"] = np_(pts)          # 2 samples in the head would make the gradients ill-conditioned)
    m.eval()
    torch.manual_seed(2)
    with Recorder(r.rep) as rec:
        y = m(pts[:2])
    out["cls_eval_out"] = np_(y)
    out.update(rec.arrays("cls_eval_tape"))
    m.train()
    m.load_state_dict(orc.synthetic_state_dict(specs["cls"]))
    tgt = torch.tensor([3, 17, 0, 39, 21, 8])
    torch.manual_seed(2)
    with Recorder(r.rep) as rec:
        y = m(pts)
    loss = orc.smooth_cls_loss(y, tgt)
    loss.backward()
    out.update(cls_train_out=np_(y), cls_train_loss=np_(loss), cls_target=np_(tgt))
    out.update(rec.arrays("cls_train_tape"))
    out.update(grads_summary(m, "cls_grad"))
    out["cls_train_rm.keepHigh.la1.fc2.norm2"] = np_(m.keepHigh.la1.fc2.norm2.running_mean)
    # ---------------- part-seg (config 2 shape, reduced batch): eval fwd B=1, train fwd+bwd B=2
    torch.manual_seed(0)
    s = r.seg.get_model(50)
    specs["seg"] = load_synth(s)
    s.drop1.p = s.drop2.p = 0.0
    torch.manual_seed(1)
    xyz = torch.rand(2, 3, 2048) * 2 - 1
    lab = torch.eye(16)[torch.randint(0, 16, (2,))].unsqueeze(1)
    out.update(seg_xyz=np_(xyz), seg_label=np_(lab))
    with ref_shim.cpu_float_tensor_redirect():
        s.eval()
        torch.manual_seed(2)
        with Recorder(r.pn2) as rec:
            y, _ = s(xyz[:1], lab[:1])
        out["seg_eval_out"] = np_(y)
        out.update(rec.arrays("seg_eval_tape"))
        s.train()
        s.load_state_dict(orc.synthetic_state_dict(specs["seg"]))
        tgt = torch.randint(0, 50, (2 * 2048,))
        torch.manual_seed(2)
        with Recorder(r.pn2) as rec:
            y, _ = s(xyz, lab)
        loss = r.seg.get_loss()(y.reshape(-1, 50), tgt, None)
        loss.backward()
        out.update(seg_train_out=np_(y[:, ::8]), seg_train_loss=np_(loss), seg_target=np_(tgt))
        out.update(rec.arrays("seg_train_tape"))
        out.update(grads_summary(s, "seg_grad"))
    return out, specs


def main():
    r = ref_shim.load()
    torch.set_num_threads(os.cpu_count())
    ops = ops_fixture(r)
    np.savez_compressed(os.path.join(HERE, "ops.npz"), **ops)
    blocks, specs_b = blocks_fixture(r)
    np.savez_compressed(os.path.join(HERE, "blocks.npz"), **blocks)
    models, specs_m = models_fixture(r)
    np.savez_compressed(os.path.join(HERE, "models.npz"), **models)
    specs_b.update(specs_m)
    with open(os.path.join(HERE, "specs.json"), "w") as f:
        json.dump(specs_b, f)
    for n in ("ops.npz", "blocks.npz", "models.npz", "specs.json"):
        print(n, os.path.getsize(os.path.join(HERE, n)) // 1024, "KiB")


if __name__ == "__main__":
    main()
