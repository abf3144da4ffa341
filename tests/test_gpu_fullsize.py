"""Parity at the FULL sizes of BASELINE.json's configs, against the CPU oracle run in the test on the same seeded
inputs (VERDICT r1 "what's weak" 1-3):

  configs[0]  classifier forward, 16 x 1024, eval()
  configs[1]  part-seg, 32 x 2048, train, forward + loss + backward (every parameter gradient, element-wise)
  configs[2]  24 000-point blocks (2 blocks), train, forward + loss + backward

Method ("teacher forcing"): the oracle's FPS / kNN indices are injected so the float comparison is not
at the mercy of an arbitrary choice among tied neighbours, while every neighbour search STILL RUNS on this path and
is audited on the very operands it saw:
  (1) exactness -- the C oracle, fed the same fp32 operands, must return bit-identical indices AND distances;
  (2) tie audit -- wherever this path's neighbours differ from the injected ones (the oracle ranked ITS features,
      which differ from ours in the last bits), the two neighbour sets must be at the same distances within a few ulp
      of the operands' squared norms: a genuine ranking error would show as a distance gap, a tie does not.
Measured errors are appended to gpurun_out/parity_r2.txt next to the tolerance they were checked against.
"""
import argparse
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
REPORT = os.path.join(ROOT, "gpurun_out", "parity_r2.txt")


def report(line):
    os.makedirs(os.path.dirname(REPORT), exist_ok=True)
    with open(REPORT, "a") as f:
        f.write(line + "\n")


def synth(orc, specs, name):
    return orc.synthetic_state_dict([tuple(e) for e in specs[name]])


def _cls(mpc, orc, specs):
    m = mpc.task_models.Model(argparse.Namespace(num_point=1024, return_dist=True, cuda_ops=True, num_class=40))
    m.load_state_dict(synth(orc, specs, "cls"))
    m.drop1.p = m.drop2.p = 0.0
    return m.cuda()


def _seg(mpc, orc, specs):
    m = mpc.task_models.get_model(50)
    m.load_state_dict(synth(orc, specs, "seg"))
    m.drop1.p = m.drop2.p = 0.0
    return m.cuda()


def audit_searches(orc, audit, tag, exact_budget=4.0e11):
    """(1) exactness vs the C oracle on the same operands (searches are skipped, largest first, once the pair-flop
    budget of the CPU check is used up -- reported), (2) tie audit of rows that differ from the injected indices."""
    n_rows = n_diff = 0
    worst_gap = 0.0
    spent, checked, skipped = 0.0, 0, 0
    for kind, idx, dist, taped, ref, qry in audit:
        B, N, C = ref.shape
        S, K = idx.shape[1], idx.shape[2]
        cost = float(B) * S * N * (2 * C + 3)
        if spent + cost <= exact_budget:
            spent += cost
            checked += 1
            d0, i0 = orc.knn_point(K, ref.detach().cpu(), qry.detach().cpu())
            assert torch.equal(idx.cpu(), i0), "%s: %s search %s in %s differs from the oracle on the same operands" % (
                tag, kind, tuple(qry.shape), tuple(ref.shape))
            assert torch.equal(dist.cpu(), d0), "%s: %s distances differ from the oracle" % (tag, kind)
        else:
            skipped += 1
        if taped is None:
            continue
        rows = (idx != taped).any(-1)
        n_rows += rows.numel()
        nd = int(rows.sum())
        if nd == 0:
            continue
        n_diff += nd
        # distances of the injected neighbours on OUR operands, fp64 expanded form, vs this path's own distances
        b, s = rows.nonzero(as_tuple=True)
        q = qry[b, s].double()                                     # [R,C]
        nb = ref[b.unsqueeze(1), taped[b, s]].double()             # [R,K,C]
        d_inj = (q * q).sum(-1, keepdim=True) + (nb * nb).sum(-1) - 2 * (nb * q.unsqueeze(1)).sum(-1)
        d_own = dist[b, s].double()
        scale = (q * q).sum(-1, keepdim=True) + (nb * nb).sum(-1).amax(-1, keepdim=True)
        gap = ((d_inj.sort(-1)[0] - d_own.sort(-1)[0]).abs() / scale.clamp_min(1e-30)).max().item()
        worst_gap = max(worst_gap, gap)
    # 1e-4 of the squared norms: the oracle ranked features that differ from ours by ~1e-6 relative per element (two
    # GEMM implementations), which moves an expanded-form distance by up to ~C * 1e-6 of the norms
    tol = 1e-4
    report("%s: %d searches (%d checked bit-exact vs the C oracle on the same operands, %d over the CPU budget); rows "
           "differing from the injected indices %d / %d; worst distance gap of a differing row %.3g of the squared "
           "norms (tolerance %.1g)" % (tag, len(audit), checked, skipped, n_diff, n_rows, worst_gap, tol))
    assert worst_gap <= tol, "%s: a differing neighbour row is not a tie (gap %.3g of the squared norms)" % (tag, worst_gap)
    return n_diff, n_rows


def compare_grads(model, P, tag, rel=5e-3, rel_cancel=2e-2, cos_min=0.9995):
    """Every parameter gradient, element-wise: |ours - oracle| <= rel * max|oracle| (+ noise floor for true zeros),
    and the two gradients point the same way (cosine).  The weights of the coordinate branch (`xyz_Trans.{q,k,v}`,
    `conv_res`: gradients that are sums over ALL B*S*K signed coordinate differences, i.e. up to 2M fp32 terms that
    cancel to ~1e-3 of their magnitude) get the wider bound rel_cancel: both sides accumulate in fp32, in different
    orders.  All errors are collected first; the five worst go to the report."""
    params = dict(model.named_parameters())
    rows = []
    for key, p in P.items():
        if not p.requires_grad:
            continue
        ours = params[key].grad
        if p.grad is None:
            assert ours is None, "%s: %s has a gradient here but none in the oracle" % (tag, key)
            continue
        assert ours is not None, "%s: %s has no gradient" % (tag, key)
        a, b = ours.detach().cpu().double().flatten(), p.grad.double().flatten()
        scale = float(b.abs().max())
        if scale < 1e-3:  # true-zero gradients (bias in front of a train-mode BatchNorm): rounding noise both sides
            assert float(a.abs().max()) < 2e-3, (tag, key)
            continue
        err = float((a - b).abs().max()) / scale
        cos = float((a * b).sum() / (a.norm() * b.norm()))
        rows.append((err, cos, key))
    rows.sort(reverse=True)
    worst_cos = min(r[1] for r in rows)
    plain = [r for r in rows if "xyz_Trans" not in r[2]]
    report("%s: %d parameter gradients compared element-wise; worst max-abs error %.3g of the gradient's max over the "
           "coordinate-branch weights (tolerance %.1g) and %.3g over all others (tolerance %.1g); worst cosine %.7f "
           "(tolerance %.4f); five worst: %s" % (tag, len(rows), rows[0][0], rel_cancel, plain[0][0], rel, worst_cos,
                                                 cos_min, "; ".join("%s %.2g" % (k, e) for e, _, k in rows[:5])))
    for err, cos, key in rows:
        lim = rel_cancel if "xyz_Trans" in key else rel
        assert err <= lim, "%s: %s gradient off by %.3g of its max (tolerance %.1g)" % (tag, key, err, lim)
        assert cos >= cos_min, "%s: %s gradient cosine %.6f" % (tag, key, cos)
    return len(rows)


def test_cls_16x1024_eval_vs_oracle(mpc, orc, golden_specs):
    """BASELINE configs[0]."""
    P = synth(orc, golden_specs, "cls")
    m = _cls(mpc, orc, golden_specs).eval()
    gen = torch.Generator().manual_seed(100)
    pts = torch.rand(16, 3, 1024, generator=gen) * 2 - 1
    ctx = orc.Ctx(train=False)
    torch.manual_seed(5)
    with torch.no_grad():
        ref = orc.cls_model(P, pts, ctx)
    starts = [t[:, 0].clone() for k, t in ctx.tape if k == "fps"]
    audit = []
    with torch.no_grad(), mpc.ops.index_tape(inject=[t for _, t in ctx.tape], fps_starts=starts, audit=audit):
        y = m(pts.cuda())
    err = float((y.cpu() - ref).abs().max())
    report("cls 16x1024 eval: log-probabilities max abs error %.3g (tolerance rtol 1e-3 atol 1e-3)" % err)
    torch.testing.assert_close(y.cpu(), ref, rtol=1e-3, atol=1e-3)
    assert (y.argmax(1).cpu() == ref.argmax(1)).all()
    audit_searches(orc, audit, "cls 16x1024 eval")
    # FPS free-running from the same start indices: bit-exact against the oracle's tape
    rec = []
    with torch.no_grad(), mpc.ops.index_tape(record=rec, fps_starts=starts):
        m(pts.cuda())
    ours = [i for k, i in rec if k == "fps"]
    theirs = [t for k, t in ctx.tape if k == "fps"]
    assert len(ours) == len(theirs) == 5
    assert torch.equal(ours[0].cpu(), theirs[0])  # (later levels sample from features-independent coordinates too)
    for a, b in zip(ours, theirs):
        assert torch.equal(a.cpu(), b)


def _seg_train_vs_oracle(mpc, orc, specs, B, N, tag, classes=50, seed=7, exact_budget=4.0e11):
    P = {k: v.clone().requires_grad_(v.dtype.is_floating_point and "running" not in k)
         for k, v in synth(orc, specs, "seg").items()}
    m = _seg(mpc, orc, specs).train()
    gen = torch.Generator().manual_seed(seed)
    xyz = torch.rand(B, 3, N, generator=gen) * 2 - 1
    lab = torch.eye(16)[torch.randint(0, 16, (B,), generator=gen)].unsqueeze(1)
    tgt = torch.randint(0, classes, (B * N,), generator=gen)
    ctx = orc.Ctx(train=True)
    torch.manual_seed(11)
    ref = orc.partseg_model(P, xyz, lab, ctx)
    ref_loss = orc.partseg_loss(ref.reshape(-1, 50), tgt)
    ref_loss.backward()
    starts = [t[:, 0].clone() for k, t in ctx.tape if k == "fps"]
    audit = []
    with mpc.ops.index_tape(inject=[t for _, t in ctx.tape], fps_starts=starts, audit=audit):
        y, _ = m(xyz.cuda(), lab.cuda())
    loss = mpc.task_models.get_loss()(y.reshape(-1, 50), tgt.cuda(), None)
    loss.backward()
    torch.cuda.synchronize()
    err = float((y.detach().cpu() - ref.detach()).abs().max())
    report("%s: logits max abs error %.3g (tolerance rtol 2e-3 atol 2e-3); loss %.7f vs oracle %.7f (tolerance rtol "
           "1e-4)" % (tag, err, loss.item(), ref_loss.item()))
    torch.testing.assert_close(y.detach().cpu(), ref.detach(), rtol=2e-3, atol=2e-3)
    np.testing.assert_allclose(loss.item(), ref_loss.item(), rtol=1e-4)
    assert compare_grads(m, P, tag) >= 400
    audit_searches(orc, audit, tag, exact_budget=exact_budget)
    # running statistics advanced exactly once and match
    key = "keepHigh.la1.fc2.norm2.running_mean"
    torch.testing.assert_close(dict(m.named_buffers())[key].cpu(), P[key], rtol=1e-4, atol=1e-5)
    return m


def test_seg_32x2048_train_vs_oracle(mpc, orc, golden_specs):
    """BASELINE configs[1]: logits, loss, every gradient element-wise, every neighbour search audited."""
    _seg_train_vs_oracle(mpc, orc, golden_specs, 32, 2048, "part-seg 32x2048 train")


def test_sem_24k_blocks_train_vs_oracle(mpc, orc, golden_specs):
    """BASELINE configs[2] shape: two 24 000-point blocks through the size-generalised module (states 24000 / 12000 /
    6000 / 3000 / 1500; cluster FPS, large feature-space searches), fwd + loss + bwd against the oracle."""
    _seg_train_vs_oracle(mpc, orc, golden_specs, 2, 24000, "24k-point blocks x2 train", seed=24,
                         exact_budget=1.0e12)


def test_seg_fps_free_running_24k(mpc, orc, golden_specs):
    """The four sampling steps of a 24 000-point block, computed by this path from the oracle's start indices, are
    bit-identical to the oracle's (24000 -> 12000 -> 6000 -> 3000 -> 1500)."""
    gen = torch.Generator().manual_seed(3)
    pts = (torch.rand(2, 24000, 3, generator=gen) * 2 - 1)
    cur_o, cur_g = pts, pts.cuda()
    for npoint in (12000, 6000, 3000, 1500):
        start = torch.randint(0, cur_o.shape[1], (2,), generator=gen)
        io = orc.farthest_point_sample(cur_o, npoint, start)
        ig = mpc.ops.farthest_point_sample(cur_g, npoint, start=start.cuda())
        assert torch.equal(ig.cpu(), io)
        cur_o, cur_g = orc.index_points(cur_o, io), mpc.ops.index_points(cur_g, ig)


@pytest.mark.parametrize("mode", ["defer_wgrad", "graph_replay"])
def test_seg_train_grads_bench_paths(mpc, orc, golden_models, golden_specs, mode):
    """The two execution modes bench.py times -- weight gradients deferred to their own stream, and the whole step
    replayed from a CUDA graph -- produce the fixture's loss and gradients (reference run, 2 x 2048)."""
    from conftest import tape_of

    g = golden_models
    T = lambda a: torch.from_numpy(np.asarray(a))
    m = _seg(mpc, orc, golden_specs).train()
    theirs = tape_of(g, "seg_train_tape")
    starts = [T(t[:, 0].copy()).cuda() for kind, t in theirs if kind == "fps"]
    inject = [T(t).cuda() for _, t in theirs]
    xyz, lab, tgt = T(g["seg_xyz"]).cuda(), T(g["seg_label"]).cuda(), T(g["seg_target"]).cuda()
    loss_fn = mpc.task_models.get_loss()
    params = list(m.parameters())

    def step():
        for p in params:
            p.grad = None
        with mpc.ops.index_tape(inject=inject, fps_starts=starts):
            y, _ = m(xyz, lab)
        loss = loss_fn(y.reshape(-1, 50), tgt, None)
        loss.backward()
        return loss

    mpc.ops.set_defer_wgrad(True)
    try:
        if mode == "graph_replay":
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
                    step()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                loss = step()
            for _ in range(2):
                graph.replay()
        else:
            loss = step()
        torch.cuda.synchronize()
    finally:
        mpc.ops.set_defer_wgrad(False)
    # (running statistics moved during the warm-up steps; the batch-statistics forward does not depend on them)
    np.testing.assert_allclose(loss.item(), g["seg_train_loss"], rtol=1e-4)
    named = dict(m.named_parameters())
    n = 0
    for k in g.files:
        if k.startswith("seg_grad."):
            ours = named[k[len("seg_grad."):]].grad.detach().flatten().cpu().numpy()
            norm, first = float(g[k][0]), g[k][1:16]
            assert abs(float(np.linalg.norm(ours)) - norm) <= 1e-2 * norm + 5e-4, k
            scale = max(float(np.abs(first).max()), norm / np.sqrt(ours.size))
            np.testing.assert_allclose(ours[:first.size], first, rtol=5e-3, atol=5e-3 * scale + 1e-5, err_msg=k)
            n += 1
    assert n >= 400


def test_accumulated_microbatches_under_deferred_wgrad(mpc, orc, golden_specs):
    """Gradient accumulation (p.grad already set in the second backward) with deferred weight gradients switched on:
    the second micro-batch must fall back to in-order launches, so the accumulated gradient equals the sum of the two
    micro-batch gradients computed separately (deferred wgrad raced with AccumulateGrad)."""
    m = mpc.pointnet2_utils.Linear(64, 128, bn=False).cuda().train()
    gen = torch.Generator().manual_seed(9)
    xs = [torch.randn(4, 4096, 64, generator=gen).cuda().requires_grad_(True) for _ in range(2)]

    def grads(defer, accumulate):
        mpc.ops.set_defer_wgrad(defer)
        try:
            out = []
            m.zero_grad(set_to_none=True)
            for x in xs:
                if not accumulate:
                    m.zero_grad(set_to_none=True)
                m(x).square().mean().backward()
                if not accumulate:
                    torch.cuda.synchronize()
                    out.append(m.linear.weight.grad.clone())
            torch.cuda.synchronize()
            return out if not accumulate else m.linear.weight.grad.clone()
        finally:
            mpc.ops.set_defer_wgrad(False)

    a, b = grads(False, False)
    acc = grads(True, True)
    torch.testing.assert_close(acc, a + b, rtol=1e-4, atol=1e-6)
    # a weight used twice in one backward
    mpc.ops.set_defer_wgrad(True)
    try:
        m.zero_grad(set_to_none=True)
        (m(xs[0]).square().mean() + m(xs[1]).square().mean()).backward()
        torch.cuda.synchronize()
        twice = m.linear.weight.grad.clone()
    finally:
        mpc.ops.set_defer_wgrad(False)
    torch.testing.assert_close(twice, a + b, rtol=1e-4, atol=1e-6)


def test_device_guard_and_knn_small_reference_sets(mpc, orc):
    """k <= N < compiled list length (k = 5, N = 6) works and matches the oracle; FPS rejects a start vector
    of the wrong shape; operands on different devices are rejected."""
    gen = torch.Generator().manual_seed(1)
    ref, qry = torch.rand(2, 6, 3, generator=gen), torch.rand(2, 4, 3, generator=gen)
    d0, i0 = orc.knn_point(5, ref, qry)
    d1, i1 = mpc.ops.knn_point(5, ref.cuda(), qry.cuda())
    assert torch.equal(i1.cpu(), i0) and torch.equal(d1.cpu(), d0)
    with pytest.raises(ValueError):
        mpc.ops.farthest_point_sample(ref.cuda(), 3, start=torch.zeros(5, dtype=torch.long).cuda())
    if torch.cuda.device_count() > 1:
        with pytest.raises(mpc._lib.MpcError):
            mpc.ops.knn_point(3, ref.cuda(0), qry.cuda(1))
        # a model on cuda:1 called while cuda:0 is current runs on cuda:1
        with torch.cuda.device(0):
            d2, i2 = mpc.ops.knn_point(5, ref.cuda(1), qry.cuda(1))
        assert torch.equal(i2.cpu(), i0)
