"""mpc_knn3_grid_f32 -- the coordinate-space neighbour search through a uniform grid -- must be BIT-IDENTICAL (indices
and distances) to the brute-force contract of mpc_knn_f32: against the C oracle at sizes it finishes in seconds and
against this library's own brute-force kernel at 24 000 .. 262 144 points; on uniform clouds, flat clouds (all points in
a plane / on a line), clustered clouds, duplicated points (exact distance ties), queries far outside the cloud, and
reference sets barely larger than k."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
REPORT = os.path.join(ROOT, "gpurun_out", "parity_r2.txt")


def report(line):
    os.makedirs(os.path.dirname(REPORT), exist_ok=True)
    with open(REPORT, "a") as f:
        f.write(line + "\n")


def run(mpc, ref, qry, k, grid):
    ops = mpc.ops
    ops.set_knn_grid(grid, min_n=1)
    try:
        d, i = ops._knn_compute(k, ref, qry)
        torch.cuda.synchronize()
        return d, i
    finally:
        ops.set_knn_grid(True, min_n=2048)


def cloud(B, N, seed, kind):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(B, N, 3, generator=g) * 2 - 1
    if kind == "plane":      # z constant: a degenerate bounding box
        x[..., 2] = 0.25
    elif kind == "line":
        x[..., 1:] = x[..., :1] * 0.5
    elif kind == "clusters":  # a few dense blobs far apart + sparse background
        c = torch.rand(B, 6, 3, generator=g) * 20 - 10
        which = torch.randint(0, 6, (B, N), generator=g)
        x = c[torch.arange(B).view(B, 1), which] + 0.01 * torch.randn(B, N, 3, generator=g)
        x[:, ::50] = torch.rand(B, (N + 49) // 50, 3, generator=g) * 40 - 20
    elif kind == "dups":     # every point appears 4 times: exact ties resolved by index
        x = x[:, : (N + 3) // 4].repeat(1, 4, 1)[:, :N]
        x = x[:, torch.randperm(N, generator=g)]
    elif kind == "s3dis":    # room-like: points on 6 planes with jitter
        face = torch.randint(0, 6, (B, N), generator=g)
        ax = face % 3
        val = (face // 3).float() * 2 - 1
        x.scatter_(2, ax.unsqueeze(-1), (val + 0.01 * torch.randn(B, N, generator=g)).unsqueeze(-1))
    return x.contiguous()


@pytest.mark.parametrize("kind", ["uniform", "plane", "line", "clusters", "dups", "s3dis"])
@pytest.mark.parametrize("B,N,S,k", [(2, 3000, 3000, 8), (3, 2049, 700, 8), (1, 5000, 1250, 16), (2, 777, 777, 3),
                                     (1, 4096, 1024, 32), (2, 40, 40, 9), (2, 9, 5, 8)])
def test_knn_grid_bit_exact_vs_oracle(mpc, orc, kind, B, N, S, k):
    ref = cloud(B, N, N + S, kind)
    qry = ref if S == N else cloud(B, S, S, kind)
    if S != N:
        qry[:, ::5] = ref[:, :S][:, ::5]     # some queries coincide with reference points
        qry[:, 1::11] *= 3.0                 # and some lie far outside the reference set's box
    d0, i0 = orc.knn_point(k, ref, qry)
    d1, i1 = run(mpc, ref.cuda(), qry.cuda(), k, grid=True)
    assert torch.equal(i1.cpu(), i0), (kind, B, N, S, k)
    assert torch.equal(d1.cpu(), d0), (kind, B, N, S, k)


@pytest.mark.parametrize("N,S,k", [(24000, 24000, 8), (24000, 12000, 8), (65536, 16384, 16), (262144, 65536, 32)])
def test_knn_grid_equals_brute_force_large(mpc, N, S, k):
    for kind in ("uniform", "s3dis"):
        ref = cloud(2 if N <= 65536 else 1, N, 7, kind).cuda()
        qry = ref if S == N else ref[:, :: N // S].contiguous()
        d0, i0 = run(mpc, ref, qry, k, grid=False)
        d1, i1 = run(mpc, ref, qry, k, grid=True)
        assert torch.equal(i1, i0) and torch.equal(d1, d0), (kind, N, S, k)


def test_knn_grid_timing_report(mpc):
    """Not a pass/fail criterion: device time of the two paths, written to the report."""
    lines = []
    for B, N, S, k in ((8, 24000, 24000, 8), (8, 24000, 12000, 8), (32, 2048, 2048, 8), (1, 262144, 65536, 16),
                       (1, 1048576, 262144, 16)):
        ref = cloud(B, N, 3, "uniform").cuda()
        qry = ref if S == N else ref[:, :: N // S].contiguous()
        t = {}
        for name, grid in (("grid", True), ("brute force", False)):
            if not grid and N > 300000:
                continue
            run(mpc, ref, qry, k, grid)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            run(mpc, ref, qry, k, grid)
            b.record()
            torch.cuda.synchronize()
            t[name] = a.elapsed_time(b)
        lines.append("%d x %d in %d, k = %d: %s" % (B, S, N, k, ", ".join("%s %.3f ms" % kv for kv in t.items())))
    report("coordinate kNN, uniform-grid search vs brute force (both bit-identical): " + "; ".join(lines))
