"""Edge cases of the point-set ops on the GPU: empty and ragged shapes, non-contiguous inputs, wrong dtypes,
large single clouds, degenerate clouds (all points identical), and the classifier at BASELINE config-1 size."""
import argparse

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def cloud(B, N, C=3, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(B, N, C, generator=g) * 2 - 1


def test_empty_batches_and_query_sets(mpc):
    ops = mpc.ops
    xyz = cloud(2, 64).cuda()
    empty_q = torch.empty(2, 0, 3, device="cuda")
    d, i = ops.knn_point(8, xyz, empty_q)
    assert d.shape == (2, 0, 8) and i.shape == (2, 0, 8)
    assert ops.farthest_point_sample(xyz, 0).shape == (2, 0)
    assert ops.query_ball_point(0.3, 4, xyz, empty_q).shape == (2, 0, 4)
    feats = torch.randn(2, 64, 16, device="cuda", requires_grad=True)
    out = ops.index_points(feats, torch.empty(2, 0, dtype=torch.long, device="cuda"))
    assert out.shape == (2, 0, 16)
    out.sum().backward()
    assert torch.count_nonzero(feats.grad) == 0
    b0 = torch.empty(0, 64, 3, device="cuda")
    assert ops.farthest_point_sample(b0, 8).shape == (0, 8)
    assert ops.knn_point(8, b0, b0)[1].shape == (0, 64, 8)


def test_non_contiguous_and_wrong_dtype(mpc, orc):
    ops = mpc.ops
    base = cloud(2, 128, 3, seed=4)
    view = base.cuda().permute(0, 2, 1).permute(0, 2, 1)[:, ::2]  # strided view, 64 points
    ref = base[:, ::2].contiguous()
    d0, i0 = orc.knn_point(8, ref, ref)
    d1, i1 = ops.knn_point(8, view, view)
    assert torch.equal(i1.cpu(), i0) and torch.equal(d1.cpu(), d0)
    with pytest.raises(TypeError):
        ops.knn_point(8, view.double(), view.double())
    with pytest.raises(TypeError):
        ops.index_points(torch.zeros(1, 4, 2, dtype=torch.float16, device="cuda"),
                         torch.zeros(1, 2, dtype=torch.long, device="cuda"))


def test_degenerate_cloud_all_points_identical(mpc, orc):
    """Every distance is an exact tie: FPS falls back to index 0 after the start, kNN returns indices 0..K-1."""
    xyz = torch.full((2, 50, 3), 0.25)
    start = torch.tensor([7, 0])
    ref = orc.farthest_point_sample(xyz, 10, start)
    out = mpc.ops.farthest_point_sample(xyz.cuda(), 10, start=start.cuda())
    assert torch.equal(out.cpu(), ref) and (ref[:, 1:] == 0).all()
    d, i = mpc.ops.knn_point(8, xyz.cuda(), xyz.cuda())
    assert torch.equal(i.cpu(), torch.arange(8).expand(2, 50, 8))
    d0, i0 = orc.knn_point(8, xyz, xyz)
    assert torch.equal(d.cpu(), d0) and torch.equal(i.cpu(), i0)


def test_knn_large_reference_set(mpc, orc):
    ref, qry = cloud(1, 100000, seed=1), cloud(1, 300, seed=2)
    d0, i0 = orc.knn_point(16, ref, qry)
    d1, i1 = mpc.ops.knn_point(16, ref.cuda(), qry.cuda())
    assert torch.equal(i1.cpu(), i0) and torch.equal(d1.cpu(), d0)


def test_ragged_sizes_every_op(mpc, orc):
    """Sizes that are not multiples of any tile: N=1031 points, S=257 queries, C=20 channels, K=5."""
    ops = mpc.ops
    g = torch.Generator().manual_seed(9)
    xyz, sub = cloud(3, 1031, seed=5), cloud(3, 257, seed=6)
    d0, i0 = orc.knn_point(5, xyz, sub)
    d1, i1 = ops.knn_point(5, xyz.cuda(), sub.cuda())
    assert torch.equal(i1.cpu(), i0) and torch.equal(d1.cpu(), d0)
    feats = torch.randn(3, 257, 20, generator=g)
    up0 = orc.upsample(feats, i0, n_out=1031)
    up1 = ops.upsample(feats.cuda(), i1, n_out=1031)
    torch.testing.assert_close(up1.cpu(), up0, rtol=1e-5, atol=1e-6)
    x = torch.randn(3 * 257, 20, generator=g).cuda()
    torch.manual_seed(9)
    lin = mpc.pointnet2_utils.Linear(20, 28, bn=False).cuda().train()  # K % 32 != 0 -> library GEMM path
    y = lin(x.view(3, 257, 20))
    z = torch.nn.functional.linear(x, lin.linear.weight, lin.linear.bias)
    z = torch.nn.functional.leaky_relu(torch.nn.functional.batch_norm(z, None, None, lin.norm2.weight, lin.norm2.bias,
                                                                       training=True), 0.2)
    torch.testing.assert_close(y.reshape(-1, 28), z, rtol=1e-4, atol=1e-5)


def test_tensor_core_linear_block_matches_library_path(mpc):
    """Linear(64, 128): the tcgen05 path (GEMM + epilogue statistics + fused BatchNorm) against the library ops."""
    g = torch.Generator().manual_seed(3)
    torch.manual_seed(3)  # the layer's random init comes from the global generator
    x = torch.randn(4, 1000, 64, generator=g).cuda().requires_grad_(True)
    lin = mpc.pointnet2_utils.Linear(64, 128, bn=False).cuda().train()
    w = torch.randn(4, 1000, 128, generator=g).cuda()
    y = lin(x)
    (y * w).sum().backward()
    xr = x.detach().clone().requires_grad_(True)
    z = torch.nn.functional.linear(xr.double(), lin.linear.weight.double(), lin.linear.bias.double())
    z = torch.nn.functional.batch_norm(z.reshape(-1, 128), None, None, lin.norm2.weight.double(),
                                       lin.norm2.bias.double(), training=True).reshape(4, 1000, 128)
    z = torch.nn.functional.leaky_relu(z, 0.2)
    (z * w.double()).sum().backward()
    torch.testing.assert_close(y, z.float(), rtol=1e-4, atol=1e-5)
    gscale = float(xr.grad.abs().max())
    torch.testing.assert_close(x.grad, xr.grad.float(), rtol=1e-3, atol=1e-4 * gscale)
    rm = lin.norm2.running_mean
    ref_mean = torch.nn.functional.linear(x.detach(), lin.linear.weight, lin.linear.bias).reshape(-1, 128).mean(0)
    torch.testing.assert_close(rm, 0.1 * ref_mean, rtol=1e-4, atol=1e-6)
    assert int(lin.norm2.num_batches_tracked) == 1


def test_classifier_config1_shape(mpc):
    """BASELINE config 1: classifier forward, batch 16 x 1024 points, eval: finite log-probabilities that sum to one;
    streams on/off agree bit for bit (same kernels, same order of arithmetic)."""
    torch.manual_seed(0)
    m = mpc.task_models.Model(argparse.Namespace(num_point=1024, return_dist=True, cuda_ops=True, num_class=40))
    m = m.cuda().eval()
    pts = (torch.rand(16, 3, 1024, generator=torch.Generator().manual_seed(1)) * 2 - 1).cuda()
    starts = [torch.randint(0, n, (16,), generator=torch.Generator().manual_seed(n)) for n in (1024, 512, 256, 128, 64)]
    with torch.no_grad(), mpc.ops.index_tape(fps_starts=starts):
        a = m(pts)
    assert a.shape == (16, 40) and torch.isfinite(a).all()
    torch.testing.assert_close(a.exp().sum(1), torch.ones(16, device="cuda"), rtol=1e-4, atol=1e-4)
    old = mpc.ops._STREAMS_ENABLED
    try:
        mpc.ops._STREAMS_ENABLED = False
        with torch.no_grad(), mpc.ops.index_tape(fps_starts=starts):
            b = m(pts)
    finally:
        mpc.ops._STREAMS_ENABLED = old
    assert torch.equal(a, b)


def test_reduction_scratch_is_left_zero(mpc):
    """include/mpc_b200.h scratch contract: zero on entry, zero again once the consumer has run -- forward (GEMM
    epilogue sums -> normalise kernel), backward (column reduction -> apply kernel) and the bias column sum."""
    torch.manual_seed(0)
    lin = mpc.pointnet2_utils.Linear(64, 128, bn=False).cuda().train()
    x = torch.randn(4, 500, 64, device="cuda", requires_grad=True)
    for _ in range(2):
        lin(x).sum().backward()
    q = torch.randn(2000, 64, device="cuda", requires_grad=True)
    w = torch.randn(32, 64, device="cuda", requires_grad=True)
    b = torch.zeros(32, device="cuda", requires_grad=True)
    mpc.ops.linear(q, w, b).sum().backward()
    torch.cuda.synchronize()
    assert len(mpc.ops._scratch_pool) >= 3
    for key, buf in mpc.ops._scratch_pool.items():
        assert float(buf.abs().sum()) == 0.0, key


def _sum_close(a, b, rtol, atol, msg=None):
    """Comparison for gradients that are long fp32 sums (weight gradients, gradients summed over rows): two correct
    evaluations differ by ~1e-6 of the tensor's largest element (accumulation order -- the split-K reduction of the
    weight-gradient GEMM is ordered by arrival), so the absolute tolerance carries a term relative to that scale."""
    torch.testing.assert_close(a, b, rtol=rtol, atol=atol + 1e-5 * float(b.abs().max()), msg=msg)


@pytest.mark.parametrize("M,C,padded", [(1, 2, False), (1000, 50, False), (65536, 50, True), (333, 13, True), (64, 300, False)])
def test_smooth_cross_entropy_matches_reference_chain(mpc, M, C, padded):
    """Fused label-smoothed CE vs the reference's op chain (R/models/repsurf/pointnet2_part_seg_msg.py:159-180) on
    the same device; loss rtol 1e-5, gradient rtol 1e-4 / atol 1e-9 (values are O(1/M))."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(M + C)
    base = torch.randn(M, (C + 3) & ~3 if padded else C, generator=g).cuda() * 3
    target = torch.randint(0, C, (M,), generator=g).cuda()
    a = base.clone().requires_grad_(True)
    pa = a[:, :C]
    loss = mpc.ops.smooth_cross_entropy(pa, target, 0.1)
    (loss * 1.7).backward()
    b = base.clone().requires_grad_(True)
    pb = b[:, :C]
    one_hot = torch.zeros_like(pb).scatter(1, target.view(-1, 1), 1)
    one_hot = one_hot * (1 - 0.1) + (1 - one_hot) * 0.1 / (C - 1)
    ref = -(one_hot * F.log_softmax(pb, dim=1)).sum(dim=1).mean()
    (ref * 1.7).backward()
    torch.testing.assert_close(loss, ref, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(a.grad, b.grad, rtol=1e-4, atol=1e-9)
    assert float(mpc.ops._ce_scratch(a.device).abs().sum()) == 0.0  # scratch contract


def test_deferred_weight_gradients_match(mpc):
    """ops.set_defer_wgrad(True) moves the wgrad GEMMs to their own stream (joined by an end-of-backward callback):
    same gradients (the split-K reduce order differs run to run: measured 4e-5 abs on O(1) sums; rtol 1e-3, atol 1e-4)."""
    torch.manual_seed(0)
    lin1 = mpc.pointnet2_utils.Linear(64, 128, bn=False).cuda().train()
    lin2 = mpc.pointnet2_utils.Linear(128, 64, bn=False).cuda().train()
    x = torch.randn(8, 700, 64, device="cuda")
    grads = []
    for on in (False, True, True):
        mpc.ops.set_defer_wgrad(on)
        for p in list(lin1.parameters()) + list(lin2.parameters()):
            p.grad = None
        xi = x.clone().requires_grad_(True)
        lin2(lin1(xi)).square().sum().backward()
        torch.cuda.current_stream().synchronize()
        grads.append([p.grad.clone() for p in list(lin1.parameters()) + list(lin2.parameters()) if p.grad is not None]
                     + [xi.grad.clone()])
    mpc.ops.set_defer_wgrad(False)
    for a, b in zip(grads[0], grads[1]):
        _sum_close(a, b, rtol=1e-3, atol=1e-4)
    for a, b in zip(grads[1], grads[2]):
        _sum_close(a, b, rtol=1e-3, atol=1e-4)


@pytest.mark.parametrize("Np", [256, 375, 1000])  # clouds that are / are not a whole number of 128-row GEMM tiles
@pytest.mark.parametrize("train", [True, False])
def test_split_projection_equals_concatenated_projection(mpc, train, Np):
    """Linear.forward_split(x_a, g) == Linear(cat(x_a, broadcast g)): outputs rtol 1e-4, gradients rtol 1e-3."""
    torch.manual_seed(3)
    B, Ka, Kb, N = 3, 64, 96, 128
    lin = mpc.pointnet2_utils.Linear(Ka + Kb, N, bn=False).cuda().train(train)
    ref = mpc.pointnet2_utils.Linear(Ka + Kb, N, bn=False).cuda().train(train)
    ref.load_state_dict(lin.state_dict())
    xa = torch.randn(B, Np, Ka, device="cuda")
    g = torch.randn(B, Kb, device="cuda")
    w = torch.randn(B, Np, N, device="cuda")
    a1, g1 = xa.clone().requires_grad_(True), g.clone().requires_grad_(True)
    y1 = lin.forward_split(a1, g1)
    (y1 * w).sum().backward()
    a2, g2 = xa.clone().requires_grad_(True), g.clone().requires_grad_(True)
    y2 = ref(torch.cat((a2, g2[:, None, :].expand(-1, Np, -1)), 2))
    (y2 * w).sum().backward()
    torch.cuda.synchronize()
    torch.testing.assert_close(y1, y2, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(a1.grad, a2.grad, rtol=1e-3, atol=1e-5)
    _sum_close(g1.grad, g2.grad, rtol=1e-3, atol=1e-4)
    for (k, p), (_, q) in zip(lin.named_parameters(), ref.named_parameters()):
        if q.grad is not None:
            _sum_close(p.grad, q.grad, rtol=1e-3, atol=1e-4, msg=k)
    if train:
        torch.testing.assert_close(lin.norm2.running_var, ref.norm2.running_var, rtol=1e-5, atol=1e-6)


def test_bn_backward_reads_strided_grad_rows(mpc):
    """Three Linear+BN blocks feeding a concatenation: cat's backward hands each block a column slice of one wide
    gradient, which the BatchNorm-backward kernels read in place (row stride 3C); gradients must equal the run in which
    every slice is copied to a contiguous tensor first."""
    torch.manual_seed(1)
    lins = [mpc.pointnet2_utils.Linear(64, 64, bn=False).cuda().train() for _ in range(3)]
    head = mpc.pointnet2_utils.Linear(192, 64, bn=False).cuda().train()
    x = torch.randn(4, 512, 64, device="cuda")

    def run(force_copy):
        old = mpc.ops._grad_rows
        if force_copy:
            mpc.ops._grad_rows = lambda g, C: (g.contiguous(), C)
        try:
            for m in lins + [head]:
                for p in m.parameters():
                    p.grad = None
            xi = x.clone().requires_grad_(True)
            head(torch.cat([m(xi) for m in lins], 2)).square().sum().backward()
            torch.cuda.synchronize()
            return [xi.grad.clone()] + [p.grad.clone() for m in lins for p in m.parameters() if p.grad is not None]
        finally:
            mpc.ops._grad_rows = old

    for a, b in zip(run(False), run(True)):
        _sum_close(a, b, rtol=1e-3, atol=1e-4)


@pytest.mark.parametrize("M,K,N,res", [(4096, 64, 64, True), (1000, 128, 256, False), (16384, 512, 1024, True),
                                       (130, 32, 64, True), (2048, 896, 512, False)])
def test_eval_block_fused_into_gemm_epilogue(mpc, M, K, N, res):
    """Inference form of the shared-MLP block (mpc_linear_affine_act_f32: BatchNorm affine + LeakyReLU + residual in the
    tcgen05 GEMM's epilogue) against the unfused kernels and against nn.Linear -> nn.BatchNorm1d.eval() -> LeakyReLU."""
    torch.manual_seed(M + N)
    m = mpc.pointnet2_utils.Linear(K, N, bn=False).cuda().eval()
    with torch.no_grad():
        m.norm2.running_mean.normal_()
        m.norm2.running_var.uniform_(0.5, 2.0)
        m.norm2.weight.normal_()
        m.norm2.bias.normal_()
    x = torch.randn(2, M // 2, K, device="cuda")
    r = torch.randn(2, M // 2, N, device="cuda") if res else None
    ref = torch.nn.functional.leaky_relu(m.norm2(m.linear(x.double().float()).reshape(-1, N)), 0.2).view(2, M // 2, N)
    if res:
        ref = ref + r
    with torch.no_grad():
        n0 = mpc.ops.launches()
        fused = m(x, residual=r)
        assert mpc.ops.launches() - n0 == 1
        mpc.ops.set_fuse_eval(False)
        try:
            unfused = m(x, residual=r)
        finally:
            mpc.ops.set_fuse_eval(True)
    torch.testing.assert_close(fused, unfused, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(fused, ref.detach(), rtol=2e-4, atol=2e-4)  # (nn.Linear on cuBLAS may use TF32-free fp32)
