"""The bucket-pruned FPS variant (csrc/fps_pruned.cu; mid-sized clouds, 1- and 2-CTA) must select exactly the indices
of the plain algorithm: against the C oracle and against this library's plain kernels, on uniform, flat, clustered and
duplicate-heavy clouds, at every size class of the dispatcher (one CTA: <= 3072 / 6144 / 12288 points, two CTAs: <= 24576
with and without a grid cell straddling the split), sampling everything (npoint = N) included."""
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


def fps(mpc, pts, npoint, start, pruned):
    lib = mpc._lib.load()
    lib.mpc_debug_set_knob(6, 1 if pruned else -1)
    try:
        out = mpc.ops.farthest_point_sample(pts, npoint, start=start)
        torch.cuda.synchronize()
        return out
    finally:
        lib.mpc_debug_set_knob(6, 0)


def cloud(B, N, seed, kind):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(B, N, 3, generator=g) * 2 - 1
    if kind == "plane":
        x[..., 2] = -0.5
    elif kind == "clusters":
        c = torch.rand(B, 5, 3, generator=g) * 8 - 4
        which = torch.randint(0, 5, (B, N), generator=g)
        x = c[torch.arange(B).view(B, 1), which] + 0.02 * torch.randn(B, N, 3, generator=g)
    elif kind == "dups":      # every point appears 3 times: exhausting the distinct points leaves zero distances
        x = x[:, : (N + 2) // 3].repeat(1, 3, 1)[:, :N]
        x = x[:, torch.randperm(N, generator=g)]
    elif kind == "onecell":   # 60 % of the points in one tiny blob: a grid cell that straddles the 2-CTA split
        m = torch.rand(B, N, generator=g) < 0.6
        x[m] = 0.3 + 1e-4 * torch.randn(int(m.sum()), 3, generator=g)
    return x.contiguous()


@pytest.mark.parametrize("kind", ["uniform", "plane", "clusters", "dups", "onecell"])
@pytest.mark.parametrize("B,N,npoint", [(3, 1500, 700), (2, 3072, 1536), (2, 5000, 2500), (2, 12000, 3000),
                                        (2, 12288, 500), (2, 13000, 2000), (2, 24000, 4000), (1, 24576, 1000),
                                        (2, 2500, 2500)])
def test_fps_pruned_vs_oracle(mpc, orc, kind, B, N, npoint):
    pts = cloud(B, N, N + npoint, kind)
    start = torch.randint(0, N, (B,), generator=torch.Generator().manual_seed(N))
    ref = orc.farthest_point_sample(pts, npoint, start)
    got = fps(mpc, pts.cuda(), npoint, start.cuda(), pruned=True)
    assert torch.equal(got.cpu(), ref), (kind, B, N, npoint)


@pytest.mark.parametrize("N", [6000, 12000, 24000])
def test_fps_pruned_equals_plain_full_depth(mpc, N):
    """The sampling steps of a 24 000-point block at full depth (npoint = N / 2), 8 clouds."""
    for kind in ("uniform", "clusters"):
        pts = cloud(8, N, 11, kind).cuda()
        start = torch.randint(0, N, (8,), generator=torch.Generator().manual_seed(1)).cuda()
        a = fps(mpc, pts, N // 2, start, pruned=False)
        b = fps(mpc, pts, N // 2, start, pruned=True)
        assert torch.equal(a, b), (kind, N)


def test_fps_pruned_timing_report(mpc):
    """Not a pass/fail criterion: us per round of the two variants, written to the report."""
    lines = []
    for B, N in ((32, 2048), (8, 3000), (8, 6000), (8, 12000), (8, 24000), (1, 24000)):
        pts = cloud(B, N, 5, "uniform").cuda()
        start = torch.zeros(B, dtype=torch.long, device="cuda")
        t = {}
        for name, pruned in (("pruned", True), ("plain", False)):
            fps(mpc, pts, N // 2, start, pruned)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            lib = mpc._lib.load()
            lib.mpc_debug_set_knob(6, 1 if pruned else -1)
            a.record()
            mpc.ops.farthest_point_sample(pts, N // 2, start=start)
            b.record()
            torch.cuda.synchronize()
            lib.mpc_debug_set_knob(6, 0)
            t[name] = 1e3 * a.elapsed_time(b) / (N // 2)
        lines.append("%d x %d -> %d: %s" % (B, N, N // 2, ", ".join("%s %.3f us/round" % kv for kv in t.items())))
    report("FPS, bucket-pruned vs plain kernels (identical indices): " + "; ".join(lines))
