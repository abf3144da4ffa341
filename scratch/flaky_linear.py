import importlib, sys, torch
sys.path.insert(0, '.')
mpc = importlib.import_module("markov-process-analysis-on-point-cloud_b200")
g = torch.Generator().manual_seed(3)
x0 = torch.randn(4, 1000, 64, generator=g).cuda()
lin = mpc.pointnet2_utils.Linear(64, 128, bn=False).cuda().train()
w = torch.randn(4, 1000, 128, generator=g).cuda()
ref_y = None
bad = 0
for it in range(300):
    x = x0.clone().requires_grad_(True)
    y = lin(x)
    (y * w).sum().backward()
    torch.cuda.synchronize()
    if ref_y is None:
        ref_y, ref_g = y.detach().clone(), x.grad.clone()
        z = torch.nn.functional.linear(x0.double(), lin.linear.weight.double(), lin.linear.bias.double())
        z = torch.nn.functional.batch_norm(z.reshape(-1, 128), None, None, lin.norm2.weight.double(), lin.norm2.bias.double(), training=True).reshape(4, 1000, 128)
        z = torch.nn.functional.leaky_relu(z, 0.2).float()
        print("vs fp64 ref:", (ref_y - z).abs().max().item())
    else:
        dy = (y.detach() - ref_y).abs().max().item(); dg = (x.grad - ref_g).abs().max().item()
        if dy > 1e-5 or dg > 1e-4:
            bad += 1
            if bad < 6: print("iter", it, "dy", dy, "dg", dg, "nan", torch.isnan(y).any().item())
print("bad iterations:", bad, "of 299")
# plain GEMM determinism
xx = torch.randn(4000, 64, device="cuda"); ww = torch.randn(128, 64, device="cuda"); bb = torch.randn(128, device="cuda")
r = None; badg = 0
for it in range(300):
    o = torch.empty(4000, 128, device="cuda"); mpc.ops._tc_gemm(xx, ww, bb, o); torch.cuda.synchronize()
    if r is None: r = o.clone()
    elif not torch.equal(o, r): badg += 1
print("gemm nondeterministic iterations:", badg)
