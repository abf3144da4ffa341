"""How far is each fp32 implementation (CPU oracle / CUDA path) from the float64 evaluation of the same network on the
same neighbourhoods?  Usage: python scratch/grad_f64.py [B] [N]"""
import importlib
import json
import os
import sys

import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
mpc = importlib.import_module("markov-process-analysis-on-point-cloud_b200")
from oracle import markov_oracle as orc  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
N = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
specs = json.load(open(os.path.join(ROOT, "tests", "golden", "specs.json")))
sd = orc.synthetic_state_dict([tuple(e) for e in specs["seg"]])


def params(dtype):
    return {k: (v.to(dtype) if v.dtype.is_floating_point else v.clone()).clone().requires_grad_(
        v.dtype.is_floating_point and "running" not in k) for k, v in sd.items()}


gen = torch.Generator().manual_seed(7)
xyz = torch.rand(B, 3, N, generator=gen) * 2 - 1
lab = torch.eye(16)[torch.randint(0, 16, (B,), generator=gen)].unsqueeze(1)
tgt = torch.randint(0, 50, (B * N,), generator=gen)

P32 = params(torch.float32)
ctx = orc.Ctx(train=True)
torch.manual_seed(11)
out32 = orc.partseg_model(P32, xyz, lab, ctx)
orc.partseg_loss(out32.reshape(-1, 50), tgt).backward()
tape = [t for _, t in ctx.tape]
starts = [t[:, 0].clone() for k, t in ctx.tape if k == "fps"]

P64 = params(torch.float64)
ctx64 = orc.Ctx(train=True, inject=tape, fps_starts=[s.clone() for s in starts])
out64 = orc.partseg_model(P64, xyz.double(), lab.double(), ctx64)
orc.partseg_loss(out64.reshape(-1, 50), tgt).backward()

m = mpc.task_models.get_model(50)
m.load_state_dict(sd)
m.drop1.p = m.drop2.p = 0.0
m = m.cuda().train()
with mpc.ops.index_tape(inject=tape, fps_starts=starts):
    y, _ = m(xyz.cuda(), lab.cuda())
mpc.task_models.get_loss()(y.reshape(-1, 50), tgt.cuda(), None).backward()
torch.cuda.synchronize()
print("logits: oracle32 vs f64 %.3g, ours vs f64 %.3g" % (float((out32.double() - out64).abs().max()),
                                                        float((y.detach().cpu().double() - out64).abs().max())))
named = dict(m.named_parameters())
rows = []
for k, p in P64.items():
    if not p.requires_grad or p.grad is None:
        continue
    t = p.grad.flatten()
    sc = float(t.abs().max())
    if sc < 1e-3:
        continue
    a = P32[k].grad.double().flatten()
    b = named[k].grad.detach().cpu().double().flatten()
    rows.append((float((b - t).abs().max()) / sc, float((a - t).abs().max()) / sc, float((a - b).abs().max()) / sc, k))
rows.sort(reverse=True)
print("%-52s %10s %10s %10s" % ("parameter (worst by ours-vs-f64)", "ours-f64", "orc32-f64", "ours-orc32"))
for r in rows[:25]:
    print("%-52s %10.3g %10.3g %10.3g" % (r[3], r[0], r[1], r[2]))
print("worst orc32-f64: %.3g; worst ours-f64: %.3g; median ours %.3g, median orc32 %.3g" % (
    max(r[1] for r in rows), rows[0][0], sorted(r[0] for r in rows)[len(rows) // 2],
    sorted(r[1] for r in rows)[len(rows) // 2]))
