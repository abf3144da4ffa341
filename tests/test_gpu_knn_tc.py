"""mpc_knn_tc_f32 -- the feature-space neighbour search with its distance GEMM on the tensor cores -- must be
BIT-IDENTICAL (indices and distances) to the FP32 brute-force contract: against the C oracle at sizes it finishes in
seconds, against this library's own brute-force kernels at 24 000 points, on random features, on tie-heavy features
(identical rows, as a Markov transition produces them: these take the exact fallback) and on near-duplicates."""
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


def run_tc(mpc, ref, qry, k=8):
    ops = mpc.ops
    ops.knn_tc_debug = []
    try:
        ops.set_knn_tc(True)
        d, i = ops._knn_compute(k, ref, qry)
        torch.cuda.synchronize()
        assert len(ops.knn_tc_debug) == 1, "the search did not take the tensor-core path"
        ws, B, N, S = ops.knn_tc_debug[0]
        head = ws[: (2 * B + 1) * 4].view(torch.int32).cpu()
        fallback = int(head[B:2 * B].sum())
        maxdev = float(head[2 * B:2 * B + 1].view(torch.float32))
        return d, i, fallback, maxdev
    finally:
        ops.knn_tc_debug = None


def run_simt(mpc, ref, qry, k=8):
    ops = mpc.ops
    ops.set_knn_tc(False)
    try:
        return ops._knn_compute(k, ref, qry)
    finally:
        ops.set_knn_tc(True)


def features(B, N, C=64, seed=0, kind="random"):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, N, C, generator=g)
    if kind == "relu":  # post-activation look: non-negative, correlated channels, varying norms
        x = torch.nn.functional.leaky_relu(x @ torch.randn(C, C, generator=g) / 8 + 0.3, 0.2) * (
            0.5 + torch.rand(B, N, 1, generator=g))
    elif kind == "ties":  # a third of the rows share ONE vector, plus duplicate pairs: what a transition leaves behind
        m = torch.rand(B, N, generator=g) < 0.33
        x[m] = x[0, 0].clone()
        x[:, 1::7] = x[:, 0:-1:7][:, : x[:, 1::7].shape[1]].clone()
    elif kind == "near":  # clusters of rows one ulp apart
        base = x[:, : N // 8].repeat(1, 8, 1)[:, :N].clone()
        bump = torch.randint(0, 3, base.shape, generator=g).to(torch.int32) - 1
        x = (base.view(torch.int32) + bump).view(torch.float32)
    return x.contiguous()


@pytest.mark.parametrize("kind", ["random", "relu", "ties", "near"])
@pytest.mark.parametrize("B,N,S", [(2, 3000, 3000), (3, 2049, 700), (1, 4096, 2048)])
def test_knn_tc_bit_exact_vs_oracle(mpc, orc, kind, B, N, S):
    ref = features(B, N, seed=N + S, kind=kind)
    qry = ref if S == N else features(B, S, seed=S, kind=kind)
    if kind == "ties" and S != N:
        qry[:, ::3] = ref[:, :S][:, ::3]  # queries that coincide with tied reference rows
    d0, i0 = orc.knn_point(8, ref, qry)
    rg, qg = ref.cuda(), (qry.cuda() if S != N else None)
    if qg is None:
        qg = rg
    d1, i1, fallback, maxdev = run_tc(mpc, rg, qg)
    assert torch.equal(i1.cpu(), i0), (kind, B, N, S)
    assert torch.equal(d1.cpu(), d0), (kind, B, N, S)
    report("knn_tc %s B=%d N=%d S=%d: bit-exact vs the C oracle; %d of %d queries took the exact fallback; worst "
           "|approx - exact| / (|q|^2 + max|r|^2) = %.3g (eps = 2^-14 = 6.1e-05)" % (kind, B, N, S, fallback, B * S, maxdev))
    assert maxdev < 6.1e-5 / 4, "the filter's error margin is thinner than designed (%.3g)" % maxdev
    if kind in ("random", "relu"):
        assert fallback <= B * S // 100


@pytest.mark.parametrize("ntie", [9, 20, 47, 48, 49, 130])
def test_knn_tc_small_fallback_counts(mpc, orc, ntie):
    """A group of `ntie` identical rows makes exactly those queries undecidable for the filter: up to 48 listed queries per
    cloud go through the one-CTA-per-query kernel, more through the tiled one -- both must reproduce the oracle."""
    ref = features(2, 3000, seed=ntie, kind="random")
    ref[:, 100:100 + ntie] = ref[:, 7:8].clone()
    ref[1, 2000:2003] = ref[1, 7:8].clone()
    d0, i0 = orc.knn_point(8, ref, ref)
    rg = ref.cuda()
    d1, i1, fallback, _ = run_tc(mpc, rg, rg)
    assert fallback >= (2 * ntie if ntie >= 20 else 1)  # (up to 8 ties still fit the 16-entry candidate list)
    assert torch.equal(i1.cpu(), i0) and torch.equal(d1.cpu(), d0)


def test_knn_tc_equals_brute_force_at_24k(mpc):
    """24 000 x 24 000 and 12 000 x 24 000 (the two largest searches of a 24 000-point block), tie-heavy features."""
    ref = features(2, 24000, seed=5, kind="relu").cuda()
    mask = torch.rand(2, 24000, device="cuda") < 0.08
    ref[mask] = ref[0, 1].clone()  # 8 % unreached points share one feature vector
    for S in (24000, 12000):
        qry = ref if S == 24000 else ref[:, ::2].contiguous()
        d0, i0 = run_simt(mpc, ref, qry)
        d1, i1, fallback, maxdev = run_tc(mpc, ref, qry)
        assert torch.equal(i1, i0) and torch.equal(d1, d0)
        report("knn_tc 2 x %d in 24000, 8%% identical rows: bit-exact vs the FP32 brute-force kernel; fallback %d of %d "
               "queries; worst deviation %.3g of the norms" % (S, fallback, 2 * S, maxdev))


def test_knn_tc_timing_report(mpc):
    """Not a pass/fail criterion: device time of the two paths on 8 x 24 000 x 24 000, written to the report."""
    ref = features(8, 24000, seed=6, kind="relu").cuda()
    out = {}
    for name, fn in (("tensor-core filter + exact refinement", run_tc), ("FP32 brute force", run_simt)):
        fn(mpc, ref, ref)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn(mpc, ref, ref)
        b.record()
        torch.cuda.synchronize()
        out[name] = a.elapsed_time(b)
    flop = 8 * 24000 * 24000 * 131 / 1e12
    report("knn 8 x 24000 x 24000, C = 64, k = 8: " + "; ".join("%s %.2f ms (%.1f TFLOP/s of brute-force work)" % (
        k, v, flop / v * 1e3) for k, v in out.items()))
