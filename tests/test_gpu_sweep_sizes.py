"""BASELINE config 4 (op sweep on 16K-1M point clouds): the variants of the kernels that only large clouds reach.
Sizes the CPU oracle finishes in seconds are compared bit-exactly; the full 1M-point size is checked through
size-independent properties (exact replay of the first rounds, uniqueness, range, neighbour-list invariants)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def cloud(B, N, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(B, N, 3, generator=g) * 2 - 1


@pytest.fixture()
def grid_fps(mpc):
    """Force the grid-wide (whole-GPU, cooperative) FPS variant for every cloud above 8192 points."""
    lib = mpc._lib.load()
    lib.mpc_debug_set_knob(3, 1)
    yield
    lib.mpc_debug_set_knob(3, 0)


@pytest.mark.parametrize("B,N,npoint", [(1, 8193, 300), (2, 20000, 256), (1, 40000, 500), (1, 70001, 128)])
def test_fps_grid_variant_vs_oracle(mpc, orc, grid_fps, B, N, npoint):
    xyz = cloud(B, N, seed=N)
    start = torch.randint(0, N, (B,), generator=torch.Generator().manual_seed(2))
    ref = orc.farthest_point_sample(xyz, npoint, start)
    out = mpc.ops.farthest_point_sample(xyz.cuda(), npoint, start=start.cuda())
    assert torch.equal(out.cpu(), ref)
    # the barrier words are left clean: a second launch gives the same answer
    out2 = mpc.ops.farthest_point_sample(xyz.cuda(), npoint, start=start.cuda())
    assert torch.equal(out2.cpu(), ref)


def _replay_rounds(xyz, idx, rounds):
    """The reference's arithmetic (R/modules/pointnet2_utils.py:99-108) for the first `rounds` rounds, on the GPU with
    plain torch ops (separately rounded sub / mul / add, strict <, lowest index on ties)."""
    N = xyz.shape[0]
    dist = torch.full((N,), 1e10, device=xyz.device)
    far = int(idx[0])
    for t in range(rounds):
        assert int(idx[t]) == far, "round %d" % t
        d = xyz - xyz[far]
        d = d * d
        d = (d[:, 0] + d[:, 1]) + d[:, 2]
        dist = torch.where(d < dist, d, dist)
        m = dist.max()
        far = int(torch.nonzero(dist == m)[0])


@pytest.mark.parametrize("N,npoint", [(262144, 4096), (1048576, 2048)])
def test_fps_large_cloud_properties(mpc, N, npoint):
    """N = 1M takes the grid-wide variant by size; 262144 the 16-CTA cluster."""
    xyz = cloud(1, N, seed=7).cuda()
    start = torch.tensor([N // 3], device="cuda")
    idx = mpc.ops.farthest_point_sample(xyz, npoint, start=start)[0]
    assert int(idx[0]) == N // 3
    assert int(idx.min()) >= 0 and int(idx.max()) < N
    assert idx.unique().numel() == npoint  # distinct points of a duplicate-free cloud are never sampled twice
    _replay_rounds(xyz[0], idx, 48)
    # farthest-point property at the end: every later sample is at least as close to the earlier ones as the
    # sample before it was (the sequence of selection distances never increases)
    sel = xyz[0, idx]
    probe = sel[-64:]
    diff = probe[:, None, :] - sel[None, :, :]
    d2 = (diff * diff).sum(-1)  # [64, npoint]
    pos = torch.arange(npoint - 64, npoint, device="cuda")[:, None]
    d2 = torch.where(torch.arange(npoint, device="cuda")[None, :] < pos, d2, torch.full_like(d2, 1e10))
    sel_dist = d2.min(dim=1).values  # distance of sample t to the samples before it
    assert bool((sel_dist[1:] <= sel_dist[:-1] * (1 + 1e-5)).all())


def test_knn_million_point_reference_set(mpc):
    """k = 16 neighbours of 4096 queries in a 1M-point cloud: ascending distances, self first, and the same lists as
    torch.topk on the reference's expanded-form distance matrix (pointnet2_utils.py:204-222)."""
    N, S, K = 1048576, 4096, 16
    xyz = cloud(1, N, seed=11).cuda()
    q = xyz[:, :S].contiguous()
    dist, idx = mpc.ops.knn_point(K, xyz, q)
    assert bool((dist[:, :, 1:] >= dist[:, :, :-1]).all())
    assert torch.equal(idx[0, :, 0], torch.arange(S, device="cuda"))
    d = -2 * torch.matmul(q[0, :256].double(), xyz[0].double().t())
    d += (q[0, :256].double() ** 2).sum(-1, keepdim=True) + (xyz[0].double() ** 2).sum(-1)[None]
    ref = d.topk(K, largest=False).indices
    got = idx[0, :256]
    same = (got.sort(dim=1).values == ref.sort(dim=1).values).float().mean()
    assert same > 0.999  # fp32 vs fp64 distances may swap a near-tie at the k-th place


@pytest.fixture()
def tiled_knn(mpc):
    """Force the register-tiled feature-space kNN kernel regardless of the query count."""
    lib = mpc._lib.load()
    lib.mpc_debug_set_knob(4, 2)
    yield
    lib.mpc_debug_set_knob(4, 0)


@pytest.mark.parametrize("B,N,S,C,K", [(2, 300, 200, 64, 8), (1, 1000, 1000, 128, 16), (2, 513, 129, 256, 8),
                                       (3, 2048, 2048, 64, 8), (1, 64, 1, 64, 3), (2, 4100, 700, 64, 9)])
def test_knn_tiled_variant_bit_exact(mpc, orc, tiled_knn, B, N, S, C, K):
    g = torch.Generator().manual_seed(N + S + C)
    ref = torch.randn(B, N, C, generator=g)
    qry = torch.randn(B, S, C, generator=g)
    if S <= N:
        qry[:, : S // 2] = ref[:, : S // 2]  # self-queries: distance ~0 first
    d0, i0 = orc.knn_point(K, ref, qry)
    d1, i1 = mpc.ops.knn_point(K, ref.cuda(), qry.cuda())
    assert torch.equal(i1.cpu(), i0)
    assert torch.equal(d1.cpu(), d0)


def test_knn_tiled_variant_identical_features(mpc, orc, tiled_knn):
    """After a transition many points share one feature vector: exact ties must resolve to the lower index."""
    g = torch.Generator().manual_seed(5)
    base = torch.randn(1, 40, 64, generator=g)
    ref = base[:, torch.randint(0, 40, (600,), generator=g)]
    d0, i0 = orc.knn_point(8, ref, ref[:, :300].contiguous())
    d1, i1 = mpc.ops.knn_point(8, ref.cuda(), ref[:, :300].contiguous().cuda())
    assert torch.equal(i1.cpu(), i0) and torch.equal(d1.cpu(), d0)
