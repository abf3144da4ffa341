"""bf16-I/O inference path (north_star "FP32 and BF16"; SURVEY.md 8c: indices from fp32 arithmetic, floats within
rtol 2e-2 / atol 2e-2 of the fp32 oracle): the fused tcgen05 block mpc_linear_bf16 (GEMM + BatchNorm-eval affine +
LeakyReLU + residual in the epilogue), the bf16 attention core, and the classifier / part-seg forward in eval() under
ops.bf16_inference() against the fp32 CPU oracle on injected indices.  Measured errors go to the parity report."""
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


@pytest.mark.parametrize("M,K,N", [(100, 64, 64), (128, 128, 256), (16384, 64, 64), (5000, 512, 1024), (4096, 896, 512),
                                   (777, 192, 64), (3000, 256, 40), (130, 1024, 50), (65536, 64, 128)])
@pytest.mark.parametrize("variant", ["plain", "affine_act_res", "f32_out"])
def test_linear_bf16_vs_fp32_of_the_rounded_operands(mpc, M, K, N, variant):
    ops = mpc.ops
    g = torch.Generator().manual_seed(M + K + N)
    x = torch.randn(M, K, generator=g).cuda()
    w = (torch.randn(N, K, generator=g) / K ** 0.5).cuda()
    x16, w16 = ops.to_bf16_rows(x), ops.to_bf16_rows(w)
    assert torch.equal(x16, x.to(torch.bfloat16))  # round-to-nearest-even, like torch
    scale = shift = res = None
    slope = 1.0
    if variant != "plain":
        scale = (torch.rand(N, generator=g) + 0.5).cuda()
        shift = torch.randn(N, generator=g).cuda()
        slope = 0.2
    if variant == "affine_act_res":
        res = torch.randn(M, N, generator=g).cuda().to(torch.bfloat16)
    out = ops.linear_bf16(x16, w16, scale, shift, slope, res, out_f32=(variant == "f32_out"))
    torch.cuda.synchronize()
    ref = x16.double() @ w16.double().t()
    if scale is not None:
        ref = ref * scale.double() + shift.double()
    ref = torch.where(ref > 0, ref, ref * slope)
    if res is not None:
        ref = ref + res.double()
    assert out.shape == (M, N)
    if variant == "f32_out":
        assert out.dtype == torch.float32
        torch.testing.assert_close(out.double(), ref, rtol=2e-5, atol=2e-5)  # fp32 accumulation of exact bf16 products
    else:
        assert out.dtype == torch.bfloat16
        # one bf16 rounding (2^-9 relative) of the fp32 result
        torch.testing.assert_close(out.double(), ref, rtol=4e-3, atol=1e-5)
        assert torch.equal(out, ref.float().to(torch.bfloat16)) or \
            float((out.double() - ref).abs().max()) <= float(ref.abs().max()) * 2 ** -8


def test_attention_bf16_vs_fp32_kernel(mpc):
    ops = mpc.ops
    g = torch.Generator().manual_seed(3)
    B, N, S, C, K = 2, 1000, 400, 128, 8
    q = torch.randn(B, S, C, generator=g).cuda().to(torch.bfloat16)
    kv = torch.randn(B, N, 2 * C, generator=g).cuda().to(torch.bfloat16)
    idx = torch.randint(0, N, (B, S, K), generator=g).cuda()
    ref = ops.AttnFeat.apply(q.float().contiguous(), kv.float().contiguous(), idx)
    import ctypes
    out = torch.empty(B, S, C, dtype=torch.bfloat16, device="cuda")
    P, I = mpc._lib.ptr, ctypes.c_int64
    mpc._lib.call("mpc_attn_feat_fwd_bf16", P(q), I(C), P(kv), ctypes.c_void_p(kv.data_ptr() + 2 * C), I(2 * C), P(idx),
                  P(out), I(B), I(S), I(N), I(K), I(C))
    torch.cuda.synchronize()
    assert torch.equal(out, ref.to(torch.bfloat16))  # same fp32 arithmetic, one rounding at the store


def _inject(ctx):
    return [t for _, t in ctx.tape], [t[:, 0].clone() for k, t in ctx.tape if k == "fps"]


def test_cls_forward_bf16_vs_fp32_oracle(mpc, orc, golden_specs):
    """BASELINE configs[0] (16 x 1024, eval) through the bf16 path, oracle indices injected."""
    P = orc.synthetic_state_dict([tuple(e) for e in golden_specs["cls"]])
    m = mpc.task_models.Model(argparse.Namespace(num_point=1024, return_dist=True, cuda_ops=True, num_class=40))
    m.load_state_dict(P)
    m = m.cuda().eval()
    gen = torch.Generator().manual_seed(100)
    pts = torch.rand(16, 3, 1024, generator=gen) * 2 - 1
    ctx = orc.Ctx(train=False)
    torch.manual_seed(5)
    with torch.no_grad():
        ref = orc.cls_model(P, pts, ctx)
    tape, starts = _inject(ctx)
    mpc._lib.profiler = {"names": {"mpc_linear_bf16", "mpc_attn_feat_fwd_bf16", "mpc_gather_bf16", "mpc_linear_fwd_f32"},
                         "calls": []}
    with mpc.ops.bf16_inference(), mpc.ops.index_tape(inject=tape, fps_starts=starts):
        y = m(pts.cuda())
    names = [c[0] for c in mpc._lib.profiler["calls"]]
    mpc._lib.profiler = None
    # the path really ran in bf16: every shared-MLP block / projection through the bf16 tcgen05 kernel, none through
    # the fp32 one; attention cores and gathers on bf16 rows
    assert names.count("mpc_linear_bf16") >= 40 and names.count("mpc_linear_fwd_f32") == 0, names
    assert names.count("mpc_attn_feat_fwd_bf16") == 10 and names.count("mpc_gather_bf16") >= 10
    with torch.no_grad(), mpc.ops.index_tape(inject=tape, fps_starts=starts):
        y32 = m(pts.cuda())
    err = float((y.float().cpu() - ref).abs().max())
    err32 = float((y32.cpu() - ref).abs().max())
    agree = int((y.argmax(1).cpu() == ref.argmax(1)).sum())
    report("cls 16x1024 eval, bf16 inference path: log-probabilities max abs error %.3g vs the fp32 oracle (fp32 path: "
           "%.3g; tolerance atol 2e-2 + rtol 2e-2), arg-max agreement %d / 16" % (err, err32, agree))
    torch.testing.assert_close(y.float().cpu(), ref, rtol=2e-2, atol=2e-2)
    assert agree >= 15


def test_seg_forward_bf16_vs_fp32_oracle(mpc, orc, golden_specs):
    """Part-seg forward (4 x 2048, eval) through the bf16 path: encoder, Markov transitions, decoder, head."""
    P = orc.synthetic_state_dict([tuple(e) for e in golden_specs["seg"]])
    m = mpc.task_models.get_model(50)
    m.load_state_dict(P)
    m = m.cuda().eval()
    gen = torch.Generator().manual_seed(7)
    B, N = 4, 2048
    xyz = torch.rand(B, 3, N, generator=gen) * 2 - 1
    lab = torch.eye(16)[torch.randint(0, 16, (B,), generator=gen)].unsqueeze(1)
    ctx = orc.Ctx(train=False)
    torch.manual_seed(11)
    with torch.no_grad():
        ref = orc.partseg_model(P, xyz, lab, ctx)
    tape, starts = _inject(ctx)
    with mpc.ops.bf16_inference(), mpc.ops.index_tape(inject=tape, fps_starts=starts):
        y, _ = m(xyz.cuda(), lab.cuda())
    y = y.float().cpu().reshape(ref.shape)
    err = float((y - ref).abs().max())
    scale = float(ref.abs().max())
    agree = float((y.argmax(-1) == ref.argmax(-1)).float().mean())
    report("part-seg 4x2048 eval, bf16 inference path: logits max abs error %.3g (logit range %.3g; tolerance atol "
           "2e-2 * range + rtol 2e-2), per-point arg-max agreement %.4f" % (err, scale, agree))
    torch.testing.assert_close(y, ref, rtol=2e-2, atol=2e-2 * scale)
    assert agree >= 0.97
