"""Launch each of the non-GEMM hot kernels once or twice at its largest shape in the part-seg step (for ncu)."""
import importlib, sys, torch
sys.path.insert(0, '.')
mpc = importlib.import_module("markov-process-analysis-on-point-cloud_b200")
ops = mpc.ops
torch.manual_seed(0)
B, N = 32, 2048
xyz = (torch.rand(B, N, 3, device="cuda") * 2 - 1)
feat = torch.randn(B, N, 64, device="cuda")
which = sys.argv[1:] or ["knn3", "knn64", "fps", "xyz", "bn", "featattn", "transition"]
for rep in range(2):
    if "knn3" in which:
        d, idx = ops.knn_point(8, xyz, xyz)
        ops.knn_point(8, xyz, xyz[:, :128].contiguous())
    if "knn64" in which:
        ops.knn_point(8, feat, feat)
    if "fps" in which:
        ops.farthest_point_sample(xyz, 1024, start=torch.zeros(B, dtype=torch.long, device="cuda"))
    if "xyz" in which:
        d, idx = ops.knn_point(8, xyz, xyz)
        m = mpc.pointnet2_utils.LocalTrans(3, 64, 8, residual=True).cuda().train()
        y = m(features=xyz, idx=idx, pos=xyz, xyz=True)
        y.sum().backward()
    if "bn" in which:
        lin = mpc.pointnet2_utils.Linear(64, 64, bn=False).cuda().train()
        x = feat.clone().requires_grad_(True)
        lin(x).sum().backward()
    if "featattn" in which:
        d, idx = ops.knn_point(8, xyz, xyz)
        m = mpc.pointnet2_utils.LocalTrans(64, 64, 8, residual=False).cuda().train()
        x = feat.clone().requires_grad_(True)
        m(features=x, idx=idx, pos=xyz).sum().backward()
    if "transition" in which:
        d, idx = ops.knn_point(8, xyz, xyz[:, :1024].contiguous())
        x = feat[:, :1024].clone().requires_grad_(True)
        ops.upsample(x, idx).sum().backward()
torch.cuda.synchronize()
print("done")
