"""CUDA kernels (through the C ABI) vs the CPU oracle and the golden fixtures.  Index ops: bit-exact.
Floats: tolerance written per test."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def T(a):
    return torch.from_numpy(np.asarray(a))


def cloud(B, N, C=3, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(B, N, C, generator=g) * 2 - 1


# ---------------------------------------------------------------------------------------------- FPS
def test_fps_golden(mpc, golden_ops):
    g = golden_ops
    for name, npoint in (("fps", 64), ("fpsdup", 32), ("fpsc5", 24)):
        ref = g[name + "_idx"]
        out = mpc.ops.farthest_point_sample(T(g[name + "_xyz"]).cuda(), npoint, start=T(ref[:, 0]).cuda())
        assert np.array_equal(out.cpu().numpy(), ref), name


@pytest.mark.parametrize("B,N,npoint", [(1, 1, 1), (2, 7, 7), (3, 128, 128), (2, 129, 40), (4, 1000, 300),
                                        (16, 1024, 512), (8, 2048, 1024), (2, 4096, 512), (2, 8192, 300),
                                        (3, 12000, 200), (2, 24000, 256), (1, 50000, 64), (1, 100000, 32),
                                        (1, 200000, 16)])
def test_fps_vs_oracle(mpc, orc, B, N, npoint):
    xyz = cloud(B, N, seed=N)
    start = torch.randint(0, N, (B,), generator=torch.Generator().manual_seed(1))
    ref = orc.farthest_point_sample(xyz, npoint, start)
    out = mpc.ops.farthest_point_sample(xyz.cuda(), npoint, start=start.cuda())
    assert torch.equal(out.cpu(), ref)


def test_fps_ties_and_duplicates(mpc, orc):
    # a regular grid is full of exact distance ties; duplicated points make the tail all-zero distances
    ax = torch.linspace(-1, 1, 8)
    grid = torch.stack(torch.meshgrid(ax, ax, ax, indexing="ij"), -1).reshape(1, -1, 3)
    xyz = torch.cat([grid, grid], 1).repeat(2, 1, 1)
    start = torch.tensor([5, 700])
    ref = orc.farthest_point_sample(xyz, 600, start)
    out = mpc.ops.farthest_point_sample(xyz.cuda(), 600, start=start.cuda())
    assert torch.equal(out.cpu(), ref)


def test_fps_feature_space(mpc, orc):
    feat = cloud(3, 500, 64, seed=3)
    start = torch.tensor([0, 10, 499])
    ref = orc.farthest_point_sample(feat, 100, start)
    out = mpc.ops.farthest_point_sample(feat.cuda(), 100, start=start.cuda())
    assert torch.equal(out.cpu(), ref)


def test_fps_consumes_cpu_generator_like_reference(mpc):
    xyz = cloud(4, 256).cuda()
    torch.manual_seed(7)
    a = mpc.ops.farthest_point_sample(xyz, 8)
    torch.manual_seed(7)
    expect = torch.randint(0, 256, (4,), dtype=torch.long)
    assert torch.equal(a[:, 0].cpu(), expect)


# ---------------------------------------------------------------------------------------------- kNN
def test_knn_golden(mpc, golden_ops):
    g = golden_ops
    for name in ("knn3", "knnself"):
        ref = T(g[name + "_ref"]).cuda()
        qry = T(g[name + "_qry"]).cuda() if name == "knn3" else ref
        d, i = mpc.ops.knn_point(8, ref, qry)
        assert np.array_equal(i.cpu().numpy(), g[name + "_idx"]), name
        assert np.array_equal(d.cpu().numpy(), g[name + "_dist"]), name  # bit-exact distances


@pytest.mark.parametrize("B,N,S,C,K", [(2, 300, 77, 3, 8), (1, 8, 8, 3, 8), (2, 2048, 2048, 3, 8), (3, 1000, 129, 3, 16),
                                       (2, 5000, 300, 3, 32), (2, 512, 700, 3, 3), (2, 300, 50, 3, 5),
                                       (2, 256, 100, 64, 8), (1, 1024, 1024, 64, 8), (2, 200, 130, 128, 8),
                                       (1, 300, 64, 256, 8), (2, 100, 40, 16, 16), (1, 70, 33, 7, 9)])
def test_knn_vs_oracle_bit_exact(mpc, orc, B, N, S, C, K):
    ref, qry = cloud(B, N, C, seed=1), cloud(B, S, C, seed=2)
    d0, i0 = orc.knn_point(K, ref, qry)
    d1, i1 = mpc.ops.knn_point(K, ref.cuda(), qry.cuda())
    assert torch.equal(i1.cpu(), i0)
    assert torch.equal(d1.cpu(), d0)


def test_knn_exact_ties_resolve_to_lower_index(mpc, orc):
    pts = cloud(1, 64, seed=5)
    ref = torch.cat([pts, pts, pts], 1)  # every distance appears three times
    d0, i0 = orc.knn_point(8, ref, pts)
    d1, i1 = mpc.ops.knn_point(8, ref.cuda(), pts.cuda())
    assert torch.equal(i1.cpu(), i0) and torch.equal(d1.cpu(), d0)


def test_knn_k_larger_than_n_raises(mpc):
    x = cloud(1, 4).cuda()
    with pytest.raises(RuntimeError):
        mpc.ops.knn_point(8, x, x)


def test_three_nn(mpc, orc):
    a, b = cloud(2, 500, seed=8), cloud(2, 100, seed=9)
    d0, i0 = orc.three_nn(a, b)
    d1, i1 = mpc.ops.three_nn(a.cuda(), b.cuda())
    assert torch.equal(i1.cpu(), i0) and torch.equal(d1.cpu(), d0)


# ---------------------------------------------------------------------------------------------- ball query
def test_ball_query(mpc, orc, golden_ops):
    g = golden_ops
    out = mpc.ops.query_ball_point(0.35, 16, T(g["ball_ref"]).cuda(), T(g["ball_qry"]).cuda())
    assert np.array_equal(out.cpu().numpy(), g["ball_idx"])
    for (B, N, S, ns, r) in [(2, 3000, 257, 32, 0.2), (1, 100, 10, 64, 0.05), (2, 1500, 300, 24, 1.5)]:
        ref, qry = cloud(B, N, seed=N), cloud(B, S, seed=S)
        assert torch.equal(mpc.ops.query_ball_point(r, ns, ref.cuda(), qry.cuda()).cpu(),
                           orc.query_ball_point(r, ns, ref, qry))


# ---------------------------------------------------------------------------------------------- gather
@pytest.mark.parametrize("C", [1, 3, 7, 64, 128])
def test_index_points_fwd_bwd(mpc, orc, C):
    g = torch.Generator().manual_seed(C)
    pts = torch.randn(3, 50, C, generator=g)
    for shape in ((3, 13), (3, 13, 8)):
        idx = torch.randint(0, 50, shape, generator=g)
        p0 = pts.clone().requires_grad_(True)
        o0 = orc.index_points(p0, idx)
        w = torch.randn(o0.shape, generator=g)
        (o0 * w).sum().backward()
        p1 = pts.cuda().requires_grad_(True)
        o1 = mpc.ops.index_points(p1, idx.cuda())
        (o1 * w.cuda()).sum().backward()
        assert torch.equal(o1.detach().cpu(), o0.detach())  # pure data movement: exact
        torch.testing.assert_close(p1.grad.cpu(), p0.grad, rtol=1e-5, atol=1e-5)  # atomic summation order


def test_index_points_int64_payload(mpc, orc):
    fps0 = torch.randint(0, 1000, (2, 200))
    comp = torch.randint(0, 200, (2, 50))
    ref = orc.index_points(fps0.unsqueeze(-1), comp).squeeze(-1)
    out = mpc.ops.index_points(fps0.cuda().unsqueeze(-1), comp.cuda()).squeeze(-1)
    assert torch.equal(out.cpu(), ref)
    wide = torch.randint(0, 1 << 40, (2, 30, 3))
    assert torch.equal(mpc.ops.index_points(wide.cuda(), comp[:, :10].clamp(max=29).cuda()).cpu(),
                       orc.index_points(wide, comp[:, :10].clamp(max=29)))


# ---------------------------------------------------------------------------------------------- transition
@pytest.mark.parametrize("name", ["a", "b", "c"])
def test_transition_golden(mpc, golden_ops, name):
    g = golden_ops
    p = T(g["tr%s_points" % name]).cuda().requires_grad_(True)
    o = mpc.ops.upsample(p, T(g["tr%s_idx" % name]).cuda(), scale_ratio=int(g["tr%s_ratio" % name]))
    np.testing.assert_allclose(o.detach().cpu().numpy(), g["tr%s_out" % name], rtol=1e-5, atol=1e-6)
    (o * T(g["tr%s_w" % name]).cuda()).sum().backward()
    np.testing.assert_allclose(p.grad.cpu().numpy(), g["tr%s_grad" % name], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("B,S,ratio,C,K", [(2, 128, 2, 256, 8), (2, 512, 4, 64, 8), (1, 100, 3, 5, 4), (2, 64, 16, 128, 8)])
def test_transition_vs_oracle(mpc, orc, B, S, ratio, C, K):
    g = torch.Generator().manual_seed(S)
    N = S * ratio
    pts = torch.randn(B, S, C, generator=g)
    pts[0, 3, 0] = 0.0
    _, idx = orc.knn_point(K, cloud(B, N, seed=1), cloud(B, S, seed=2))
    idx[0, 5, 1] = idx[0, 5, 0]  # a repeated neighbour inside one row counts once
    p0 = pts.clone().requires_grad_(True)
    o0 = orc.upsample(p0, idx, n_out=N)
    w = torch.randn(o0.shape, generator=g)
    (o0 * w).sum().backward()
    p1 = pts.cuda().requires_grad_(True)
    o1 = mpc.ops.upsample(p1, idx.cuda(), scale_ratio=ratio)
    (o1 * w.cuda()).sum().backward()
    torch.testing.assert_close(o1.detach().cpu(), o0.detach(), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(p1.grad.cpu(), p0.grad, rtol=1e-5, atol=1e-6)
    assert (o0.detach().abs().sum(-1) == 0).any()  # the case includes unreached rows (-> 0)


@pytest.mark.parametrize("C", [4, 8, 12, 32, 64, 128, 256, 512])
def test_transition_group_kernel_shapes(mpc, orc, C):
    """Every lane grouping of the cooperative gather kernel (C/4 = 1 .. 128 lanes per row; C = 12 takes the per-thread
    kernel), lists longer than a lane group (many sources pointing at one target), empty lists, bit-identical to the
    oracle's ascending-source summation; and the reverse-neighbour lists shared between two feature tensors."""
    g = torch.Generator().manual_seed(C)
    B, S, K, N = 2, 300, 8, 640
    pts = torch.randn(B, S, C, generator=g)
    pts[1, 7, 0] = 0.0
    idx = torch.randint(0, N, (B, S, K), generator=g)
    idx[0, :150, 0] = 17          # a list of 150 sources: longer than any lane group
    idx[1, :40, 3] = 5            # and one of 40
    idx[:, :, 1] = idx[:, :, 0]   # duplicates inside a row count once
    o0 = orc.upsample(pts, idx, n_out=N)
    o1 = mpc.ops.upsample(pts.cuda(), idx.cuda(), n_out=N)
    assert torch.equal(o1.cpu(), o0)
    # one neighbour table, two feature tensors: the lists are built once inside a geometry scope
    pts2 = torch.randn(B, S, C, generator=g)
    idx_g = idx.cuda()
    n0 = mpc.ops.launches()
    with mpc.ops.geometry_scope():
        a = mpc.ops.upsample(pts.cuda(), idx_g, n_out=N)
        b = mpc.ops.upsample(pts2.cuda(), idx_g, n_out=N)
    assert mpc.ops.launches() - n0 == 3  # build + apply + apply
    assert torch.equal(a.cpu(), o0) and torch.equal(b.cpu(), orc.upsample(pts2, idx, n_out=N))


# ---------------------------------------------------------------------------------------------- three_interpolate
@pytest.mark.parametrize("C", [5, 64])
def test_three_interpolate(mpc, orc, C):
    g = torch.Generator().manual_seed(C)
    a, b = cloud(2, 300, seed=1), cloud(2, 60, seed=2)
    d, i = orc.three_nn(a, b)
    p2 = torch.randn(2, 60, C, generator=g)
    p0 = p2.clone().requires_grad_(True)
    o0 = orc.three_interpolate(p0, d, i)
    w = torch.randn(o0.shape, generator=g)
    (o0 * w).sum().backward()
    p1 = p2.cuda().requires_grad_(True)
    o1 = mpc.ops.three_interpolate(p1, d.cuda(), i.cuda())
    (o1 * w.cuda()).sum().backward()
    torch.testing.assert_close(o1.detach().cpu(), o0.detach(), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(p1.grad.cpu(), p0.grad, rtol=1e-4, atol=1e-5)


# ---------------------------------------------------------------------------------------------- BN + LeakyReLU
@pytest.mark.parametrize("M,C,train,slope", [(51, 20, True, 0.2), (4096, 64, True, 0.2), (1000, 256, False, 0.2),
                                            (333, 7, True, 1.0), (65536, 64, True, 0.2)])
def test_bn_act(mpc, M, C, train, slope):
    g = torch.Generator().manual_seed(M)
    y = torch.randn(M, C, generator=g) * 2 + 0.5
    gamma, beta = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.1
    rm, rv = torch.randn(C, generator=g) * 0.1, torch.rand(C, generator=g) + 0.5
    w = torch.randn(M, C, generator=g)
    y0, g0, b0 = y.clone().requires_grad_(True), gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    rm0, rv0 = rm.clone(), rv.clone()
    o0 = torch.nn.functional.batch_norm(y0, rm0, rv0, g0, b0, training=train, momentum=0.1, eps=1e-5)
    if slope != 1.0:
        o0 = torch.nn.functional.leaky_relu(o0, slope)
    (o0 * w).sum().backward()
    y1, g1, b1 = (t.cuda().requires_grad_(True) for t in (y, gamma, beta))
    rm1, rv1, nbt = rm.cuda(), rv.cuda(), torch.zeros((), dtype=torch.long).cuda()
    o1 = mpc.ops.bn_act(y1, g1, b1, rm1, rv1, nbt, training=train, slope=slope)
    (o1 * w.cuda()).sum().backward()
    torch.testing.assert_close(o1.detach().cpu(), o0.detach(), rtol=1e-5, atol=2e-6)
    torch.testing.assert_close(rm1.cpu(), rm0, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(rv1.cpu(), rv0, rtol=1e-5, atol=1e-6)
    assert int(nbt) == (1 if train else 0)
    torch.testing.assert_close(y1.grad.cpu(), y0.grad, rtol=1e-4, atol=2e-5)
    scale = float(g0.grad.abs().max())
    torch.testing.assert_close(g1.grad.cpu(), g0.grad, rtol=1e-4, atol=1e-5 * max(1.0, scale))
    torch.testing.assert_close(b1.grad.cpu(), b0.grad, rtol=1e-4, atol=1e-5 * max(1.0, float(b0.grad.abs().max())))


# ---------------------------------------------------------------------------------------------- tcgen05 GEMM
@pytest.mark.parametrize("M,K,N", [(128, 32, 64), (1000, 64, 64), (65536, 64, 64), (4096, 256, 256), (333, 896, 512),
                                   (128, 128, 50), (70000, 192, 64), (8192, 512, 1024), (5, 64, 128), (4096, 64, 16)])
def test_linear_3xtf32(mpc, M, K, N):
    """y = x W^T + b on the tensor cores must be fp32-accurate (the reference is fp32): compared with an fp64
    product, the error has to be of the same order as the library fp32 GEMM's, far below one TF32 pass."""
    g = torch.Generator().manual_seed(M + K + N)
    x = torch.randn(M, K, generator=g).cuda()
    w = (torch.randn(N, K, generator=g) / K ** 0.5).cuda()
    b = torch.randn(N, generator=g).cuda()
    ref = (x.double() @ w.double().t() + b.double())
    y = mpc.ops.linear(x, w, b)
    lib = torch.nn.functional.linear(x, w, b)
    err = (y.double() - ref).abs().max().item()
    err_lib = (lib.double() - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= 4e-6 * scale + 4 * err_lib, (err, err_lib, scale)
    # gradients through the autograd Function
    xg, wg, bg = x.clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    go = torch.randn(M, N, generator=g).cuda()
    (mpc.ops.linear(xg, wg, bg) * go).sum().backward()
    torch.testing.assert_close(xg.grad, (go.double() @ w.double()).float(), rtol=1e-4, atol=1e-4)
    # grad-weight sums M products: compare with fp64 relative to the result's scale, next to the library's error
    ref_w = go.double().t() @ x.double()
    err_w = (wg.grad.double() - ref_w).abs().max().item()
    err_w_lib = ((go.t() @ x).double() - ref_w).abs().max().item()
    assert err_w <= 4e-6 * ref_w.abs().max().item() + 4 * err_w_lib, (err_w, err_w_lib, ref_w.abs().max().item())
    torch.testing.assert_close(bg.grad, go.double().sum(0).float(), rtol=1e-4, atol=1e-5 * M ** 0.5 + 1e-4)


def test_linear_3xtf32_strided_operands(mpc):
    g = torch.Generator().manual_seed(3)
    big = torch.randn(2000, 192, generator=g).cuda()
    x = big[:, 64:128]  # a column slice: row stride 192
    w = torch.randn(96, 64, generator=g).cuda()
    out = torch.empty(2000, 96, device="cuda")
    mpc.ops._tc_gemm(x, w, None, out)
    torch.testing.assert_close(out, (x.double() @ w.double().t()).float(), rtol=1e-5, atol=1e-5)


def test_sample_data_side_fps(mpc, orc):
    """sample(nsample, feature[B,C,N]) of the training scripts: FPS on the xyz channels, gather of all channels."""
    g = torch.Generator().manual_seed(11)
    feat = torch.rand(3, 6, 500, generator=g) * 2 - 1
    torch.manual_seed(5)
    out = mpc.ops.sample(64, feat.cuda())
    torch.manual_seed(5)
    start = torch.randint(0, 500, (3,), dtype=torch.long)
    idx = orc.farthest_point_sample(feat[:, :3].permute(0, 2, 1).contiguous(), 64, start)
    ref = torch.gather(feat, 2, idx.unsqueeze(1).expand(-1, 6, -1))
    assert out.shape == (3, 6, 64) and torch.equal(out.cpu(), ref)


@pytest.mark.parametrize("C", [1, 6, 64, 200])
def test_index_points_bf16_fwd_bwd(mpc, C):
    """bf16 payloads: the forward gather is bit-exact (a byte mover); the backward accumulates in fp32 and rounds once,
    so it matches the fp32 scatter-add of the same bf16 gradient rows to one bf16 rounding (rtol 2e-2 stated in
    SURVEY 8c for bf16 I/O; measured far below)."""
    g = torch.Generator().manual_seed(C)
    pts = torch.randn(3, 300, C, generator=g).to(torch.bfloat16).cuda()
    idx = torch.randint(0, 300, (3, 70, 8), generator=g).cuda()
    p1 = pts.clone().requires_grad_(True)
    out = mpc.ops.index_points(p1, idx)
    assert out.dtype == torch.bfloat16
    ref = pts[torch.arange(3, device="cuda")[:, None, None], idx]
    assert torch.equal(out, ref)
    w = torch.randn(out.shape, generator=g).to(torch.bfloat16).cuda()
    out.backward(w)
    acc = torch.zeros(3, 300, C, device="cuda")
    acc.view(3, 300, C).scatter_add_(1, idx.reshape(3, -1, 1).expand(-1, -1, C), w.float().reshape(3, -1, C))
    torch.testing.assert_close(p1.grad.float(), acc, rtol=2e-2, atol=2e-2)
    torch.testing.assert_close(p1.grad, acc.to(torch.bfloat16), rtol=1e-2, atol=1e-2)


def test_reduction_scratch_bytes(mpc):
    lib = mpc._lib.load()
    assert lib.mpc_reduction_scratch_bytes(64) == (2 * 64 + 2) * 8
