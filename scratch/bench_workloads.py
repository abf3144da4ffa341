"""Extra workloads named by BASELINE.json's metric, outside the contract line of bench.py:
  cls1024 : classifier (Model, 40 classes), 32 clouds x 1024 points per GPU, train step fwd+bwd (config 5 shape)
  sem24k  : generalised part-seg module on 24 000-point blocks, 13 classes, B blocks per GPU, fwd+bwd (config 3 shape)
Same method as bench.py: whole step captured in a CUDA graph, CUDA events per replay, L2 flushed between replays."""
import argparse, importlib, sys, torch
sys.path.insert(0, '.')
mpc = importlib.import_module("markov-process-analysis-on-point-cloud_b200")
ap = argparse.ArgumentParser(); ap.add_argument("workload"); ap.add_argument("--batch", type=int, default=None)
ap.add_argument("--steps", type=int, default=10); a = ap.parse_args()
dev = torch.device("cuda"); gen = torch.Generator().manual_seed(1); torch.manual_seed(0)
if a.workload == "cls1024":
    B, N = a.batch or 32, 1024
    model = mpc.task_models.Model(argparse.Namespace(num_point=N, return_dist=True, cuda_ops=True, num_class=40)).to(dev).train()
    loss_fn = mpc.task_models.SmoothClsLoss()
    xyz = (torch.rand(B, 3, N, generator=gen) * 2 - 1).to(dev); tgt = torch.randint(0, 40, (B,), generator=gen).to(dev)
    sizes = (1024, 512, 256, 128, 64)
    fwd = lambda: loss_fn(model(xyz), tgt)
else:
    B, N = a.batch or 8, 24000
    model = mpc.task_models.get_model(13).to(dev).train()
    loss_fn = mpc.task_models.get_loss()
    xyz = (torch.rand(B, 3, N, generator=gen) * 2 - 1).to(dev)
    lab = torch.eye(16)[torch.randint(0, 16, (B,), generator=gen)].unsqueeze(1).to(dev)
    tgt = torch.randint(0, 13, (B * N,), generator=gen).to(dev)
    sizes = (N, N // 2, N // 4, N // 8)
    fwd = lambda: loss_fn(model(xyz, lab)[0].reshape(-1, 13), tgt, None)
starts = [torch.randint(0, n, (B,), generator=gen).to(dev) for n in sizes]
params = list(model.parameters())
def step():
    for p in params: p.grad = None
    with mpc.ops.index_tape(fps_starts=starts):
        loss = fwd()
    loss.backward(); return loss
side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for _ in range(3): step()
torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g): loss = step()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev); ms = 0.0
for _ in range(a.steps):
    flush.zero_(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize(); ms += e0.elapsed_time(e1)
ms /= a.steps
print("%s: B=%d N=%d  %.3f ms/step  %.1f clouds/s  (%.2f M points/s)  loss %.4f" % (a.workload, B, N, ms, B / ms * 1e3, B * N / ms / 1e3, loss.item()))
