"""The two other workloads BASELINE.json's metric names, outside bench.py's contract line (configs[1]):

  cls1024 (configs[4]): data-parallel TRAINING STEP of the ModelNet40-shaped classifier, 32 clouds x 1024 points per
           GPU (256 at 8 GPUs): gradients reset, forward, label-smoothed loss, backward, one coalesced NCCL all-reduce
           of the fp32 gradients, Adam step (lr 1e-3, weight decay 1e-4: R/tool/train_cls_scanobjectnn.py:206-211;
           torch's fused capturable Adam -- optimiser arithmetic is host-side glue, not part of the point-set path).
  sem24k  (configs[2]): S3DIS-shaped semantic segmentation, 24 000-point blocks through the size-generalised part-seg
           module, 13 classes, 8 blocks in total sharded by batch over the GPUs (8/4/2/1 per GPU), fwd+bwd + all-reduce.

    python bench_workloads.py cls1024|sem24k [--steps K] [--batch B_per_gpu]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 bench_workloads.py ...

Same method as bench.py: the device part of the step is captured once into a CUDA graph, replays are timed with CUDA
events (L2 flushed between them), max over ranks; rank 0 prints one JSON line."""
import argparse
import importlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
PKG = "markov-process-analysis-on-point-cloud_b200"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("workload", choices=["cls1024", "sem24k"])
    ap.add_argument("--batch", type=int, default=None, help="clouds per GPU")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    a = ap.parse_args()
    rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    json_out = os.fdopen(os.dup(1), "w")  # stdout carries the one JSON line only (NCCL prints its banner on fd 1)
    sys.stdout.flush()
    os.dup2(2, 1)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    mpc = importlib.import_module(PKG)
    mpc._lib.load()
    if os.environ.get("MPC_DEFER_WGRAD", "1") == "1":
        mpc.ops.set_defer_wgrad(True)  # gradients are read after backward only (all-reduce / optimiser): see ops._defer_wgrad
    gen = torch.Generator().manual_seed(1 + rank)
    torch.manual_seed(0)
    opt = None
    if a.workload == "cls1024":
        B, N = a.batch or 32, 1024
        args = argparse.Namespace(num_point=N, return_dist=True, cuda_ops=True, num_class=40)
        model = mpc.task_models.Model(args).to(dev).train()
        loss_fn = mpc.task_models.SmoothClsLoss()
        xyz = (torch.rand(B, 3, N, generator=gen) * 2 - 1).to(dev)
        tgt = torch.randint(0, 40, (B,), generator=gen).to(dev)
        sizes = (1024, 512, 256, 128, 64)
        fwd = lambda: loss_fn(model(xyz), tgt)
        scaling, total = "weak", B * world
    else:
        total = 8
        B, N = a.batch or max(1, total // world), 24000
        model = mpc.task_models.get_model(13).to(dev).train()
        loss_fn = mpc.task_models.get_loss()
        xyz = (torch.rand(B, 3, N, generator=gen) * 2 - 1).to(dev)
        lab = torch.eye(16)[torch.randint(0, 16, (B,), generator=gen)].unsqueeze(1).to(dev)
        tgt = torch.randint(0, 13, (B * N,), generator=gen).to(dev)
        sizes = (N, N // 2, N // 4, N // 8)
        fwd = lambda: loss_fn(model(xyz, lab)[0].reshape(-1, 13), tgt, None)
        scaling, total = ("strong" if a.batch is None else "weak"), B * world
    starts = [torch.randint(0, n, (B,), generator=gen).to(dev) for n in sizes]
    params = list(model.parameters())

    def device_step():
        for p in params:
            p.grad = None
        with mpc.ops.index_tape(fps_starts=starts):
            loss = fwd()
        loss.backward()
        return loss

    # warm-up on a side stream (eager), create the optimiser state, then capture
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(max(3, a.warmup)):
            device_step()
            mpc.dist.allreduce_mean_grads(params, world)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    used = [p for p in params if p.grad is not None]  # the reference's constructed-but-unused modules get no gradient
    if a.workload == "cls1024":
        opt = torch.optim.Adam(used, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4, fused=True,
                               capturable=True)
        opt.step()  # materialise the moment buffers outside the graph
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        loss = device_step()
    gopt = None
    if opt is not None:
        grads = [p.grad for p in used]  # static addresses: the graph's private pool hands out the same buffers
        gopt = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gopt):
            opt.step()

    def step():
        g.replay()
        mpc.dist.allreduce_mean_grads(params, world)  # NCCL, outside the graph
        if gopt is not None:
            gopt.replay()

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(2):
        step()
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    ms = 0.0
    for _ in range(a.steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step()
        e1.record()
        torch.cuda.synchronize()
        ms += e0.elapsed_time(e1)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms = float(t[0]) / a.steps
    if rank == 0:
        print(json.dumps({
            "metric": "point clouds/s (%s)" % ("classifier 1024 points, training step with Adam" if opt else
                                               "24 000-point blocks, fwd+bwd"),
            "value": total / ms * 1e3, "unit": "clouds/s", "n_gpus": world, "steps": a.steps, "ms_per_step": ms,
            "scaling": scaling, "dtype": "f32", "data": "synthetic",
            "config": {"workload": a.workload, "clouds_per_gpu": B, "points": N, "points_per_s": total * N / ms * 1e3,
                       "optimizer": "Adam (torch fused, capturable)" if opt else None,
                       "launch": "CUDA graph replay + NCCL all-reduce" if world > 1 else "CUDA graph replay"},
            "last_loss": float(loss.detach())}), file=json_out, flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
