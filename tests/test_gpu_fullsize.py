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


def _grad_rows(grads, P64, tag, who):
    """[(max-abs error / max|exact|, cosine, key)] of `grads` (key -> tensor or None) against the float64 gradients."""
    rows = []
    for key, p in P64.items():
        if not (p.dtype.is_floating_point and p.requires_grad):
            continue
        g = grads.get(key)
        if p.grad is None:
            assert g is None, "%s: %s has a gradient in %s but none in the float64 oracle" % (tag, key, who)
            continue
        assert g is not None, "%s: %s has no gradient in %s" % (tag, key, who)
        a, b = g.detach().cpu().double().flatten(), p.grad.flatten()
        scale = float(b.abs().max())
        if scale < 1e-3:  # true-zero gradients (bias in front of a train-mode BatchNorm): rounding noise on every side
            assert float(a.abs().max()) < 2e-3, (tag, key, who)
            continue
        rows.append((float((a - b).abs().max()) / scale, float((a * b).sum() / (a.norm() * b.norm())), key))
    rows.sort(reverse=True)
    return rows


def compare_grads(model, P32, P64, tag, cos_min=0.999):
    """Every parameter gradient, element-wise, against the FLOAT64 evaluation of the oracle on the same neighbourhoods.

    Why not "ours == fp32 oracle within 5e-3": the network's gradient is discontinuous.  `max_K(attention * v)`
    (R/modules/pointnet2_utils.py:543,568) routes each channel's gradient to ONE neighbour; wherever two products are
    within rounding of each other, the route flips with the last bits of the forward.  Measured (scratch/grad_f64.py,
    tests/test_oracle_golden.py::test_gradient_is_discontinuous): a 1e-7 relative perturbation of the weights, in
    float64 arithmetic, moves gradient elements by up to 7 % of the gradient's max at 2 x 512 points; at 32 x 2048 the
    fp32 CPU oracle ITSELF is up to 1.4 % (median 0.24 %) away from its own float64 evaluation.  So two correct fp32
    implementations with different rounding (ATen/MKL vs 3xTF32 tensor-core GEMMs) cannot agree element-wise any
    better than each agrees with float64.  The yardstick is therefore the fp32 oracle's own distance to float64:
      * median and 90th-percentile error of this path <= 2 x the fp32 oracle's (+1e-4),
      * worst error of this path <= 5 x the fp32 oracle's worst (and never above 10 % of the gradient's max),
      * every gradient points the same way as the float64 one (cosine >= 0.999).
    Sign errors, transposed or permuted slices and missing terms show as cosine << 1 / errors of order 1.
    Flip-free gradient parity (1e-4-level) is pinned at op level (tests/test_gpu_modules.py, test_gpu_ops.py)."""
    ours = _grad_rows({k: p.grad for k, p in model.named_parameters()}, P64, tag, "this path")
    orc32 = _grad_rows({k: p.grad for k, p in P32.items() if p.dtype.is_floating_point}, P64, tag, "the fp32 oracle")
    q = lambda rows, f: sorted(r[0] for r in rows)[min(len(rows) - 1, int(f * len(rows)))]
    med_o, p90_o, worst_o = q(ours, 0.5), q(ours, 0.9), ours[0][0]
    med_r, p90_r, worst_r = q(orc32, 0.5), q(orc32, 0.9), orc32[0][0]
    worst_cos = min(r[1] for r in ours)
    report("%s: %d parameter gradients compared element-wise with the float64 oracle; error as a fraction of the "
           "gradient's max -- this path: median %.3g, p90 %.3g, worst %.3g (%s); fp32 CPU oracle: median %.3g, p90 %.3g, "
           "worst %.3g (%s); bounds: median/p90 <= 2x, worst <= 5x the fp32 oracle's; worst cosine of this path %.7f "
           "(>= %.3f)" % (tag, len(ours), med_o, p90_o, worst_o, ours[0][2], med_r, p90_r, worst_r, orc32[0][2],
                          worst_cos, cos_min))
    assert med_o <= 2 * med_r + 1e-4, "%s: median gradient error %.3g vs the fp32 oracle's %.3g" % (tag, med_o, med_r)
    assert p90_o <= 2 * p90_r + 1e-4, "%s: p90 gradient error %.3g vs the fp32 oracle's %.3g" % (tag, p90_o, p90_r)
    assert worst_o <= min(5 * worst_r + 1e-3, 0.1), "%s: %s gradient off by %.3g of its max (fp32 oracle's worst %.3g)" % (
        tag, ours[0][2], worst_o, worst_r)
    for err, cos, key in ours:
        assert cos >= cos_min, "%s: %s gradient cosine %.6f" % (tag, key, cos)
    return len(ours)


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
    sd = synth(orc, specs, "seg")
    as_params = lambda dt: {k: (v.to(dt) if v.dtype.is_floating_point else v.clone()).clone().requires_grad_(
        v.dtype.is_floating_point and "running" not in k) for k, v in sd.items()}
    P = as_params(torch.float32)
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
    tape = [t for _, t in ctx.tape]
    starts = [t[:, 0].clone() for k, t in ctx.tape if k == "fps"]
    # the same network on the same neighbourhoods in float64: the yardstick for the gradients (see compare_grads)
    P64 = as_params(torch.float64)
    ref64 = orc.partseg_model(P64, xyz.double(), lab.double(), orc.Ctx(train=True, inject=tape))
    orc.partseg_loss(ref64.reshape(-1, 50), tgt).backward()
    audit = []
    with mpc.ops.index_tape(inject=tape, fps_starts=starts, audit=audit):
        y, _ = m(xyz.cuda(), lab.cuda())
    loss = mpc.task_models.get_loss()(y.reshape(-1, 50), tgt.cuda(), None)
    loss.backward()
    torch.cuda.synchronize()
    err = float((y.detach().cpu() - ref.detach()).abs().max())
    err64 = float((y.detach().cpu().double() - ref64.detach()).abs().max())
    report("%s: logits max abs error %.3g vs the fp32 oracle, %.3g vs its float64 evaluation (fp32 oracle vs float64: "
           "%.3g; tolerance rtol 2e-3 atol 2e-3); loss %.7f vs oracle %.7f (tolerance rtol 1e-4)"
           % (tag, err, err64, float((ref.detach().double() - ref64.detach()).abs().max()), loss.item(), ref_loss.item()))
    torch.testing.assert_close(y.detach().cpu(), ref.detach(), rtol=2e-3, atol=2e-3)
    np.testing.assert_allclose(loss.item(), ref_loss.item(), rtol=1e-4)
    assert compare_grads(m, P, P64, tag) >= 120  # (parameters whose gradient is not identically zero)
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

    # the path's own in-order eager gradients: what the two execution modes must reproduce
    step()
    torch.cuda.synchronize()
    eager = {k: p.grad.detach().clone() for k, p in m.named_parameters() if p.grad is not None}
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
    # (1) same gradients as the in-order eager step: the forward is deterministic, the backward differs only in the
    #     order of float atomics (attention / gather scatter-adds) -> 1e-4 of each gradient's max
    for k, ge in eager.items():
        gm = named[k].grad
        assert gm is not None, k
        scale = float(ge.abs().max())
        assert float((gm - ge).abs().max()) <= 1e-4 * scale + 1e-6, "%s: %s differs from the eager step" % (mode, k)
    # (2) the reference run's gradients (fixture: norm + first 15 elements).  Element-wise agreement between two fp32
    #     implementations is limited by the arg-max routing of the attention (see compare_grads): 5 % of the
    #     gradient's scale per element, 1 % on the norm
    n = 0
    for k in g.files:
        if k.startswith("seg_grad."):
            ours = named[k[len("seg_grad."):]].grad.detach().flatten().cpu().numpy()
            norm, first = float(g[k][0]), g[k][1:16]
            assert abs(float(np.linalg.norm(ours)) - norm) <= 1e-2 * norm + 5e-4, k
            scale = max(float(np.abs(first).max()), norm / np.sqrt(ours.size))
            np.testing.assert_allclose(ours[:first.size], first, rtol=0, atol=5e-2 * scale + 1e-5, err_msg=k)
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
