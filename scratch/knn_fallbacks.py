"""How many queries of each tensor-core feature-space search of one 24 000-point-block forward take the exact fallback."""
import importlib, sys, torch
sys.path.insert(0, '.')
mpc = importlib.import_module("markov-process-analysis-on-point-cloud_b200")
ops = mpc.ops
torch.manual_seed(0)
m = mpc.task_models.get_model(13).cuda().train()
m.keepHigh.conv7.eval()
g = torch.Generator().manual_seed(1)
B, N = 4, 24000
xyz = (torch.rand(B, 3, N, generator=g) * 2 - 1).cuda()
lab = torch.eye(16)[torch.randint(0, 16, (B,), generator=g)].unsqueeze(1).cuda()
for _ in range(2):
    ops.knn_tc_debug = []
    with torch.no_grad():
        m(xyz, lab)
    torch.cuda.synchronize()
for ws, b, n, s in ops.knn_tc_debug:
    head = ws[: (2 * b + 1) * 4].view(torch.int32).cpu()
    print("search %6d queries in %6d points: fallback per cloud %s (%.1f %%)" % (s, n, head[b:2 * b].tolist(), 100.0 * float(head[b:2 * b].sum()) / (b * s)))
ops.knn_tc_debug = None
