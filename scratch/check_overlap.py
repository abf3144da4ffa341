"""Under torchrun (N >= 2): the overlapped two-part gradient exchange of bench.Step (part 0 all-reduced from a tensor hook
during backward, part 1 after it) yields the same averaged gradients as pack + one all-reduce after backward."""
import importlib, os, sys, torch
sys.path.insert(0, '.')
import bench
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
torch.distributed.init_process_group("nccl", device_id=dev)
mpc = importlib.import_module(bench.PKG)
mpc._lib.load()
mpc.ops.set_defer_wgrad(True)
for key in ("cls1024_train", "partseg2048"):
    wl = bench.Workload(key, world)
    wl.B = 4
    step = bench.Step(wl, mpc, dev, world)
    for m in step.model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    gen = torch.Generator().manual_seed(10 + rank)
    inputs = [t.to(dev) for t in wl.synth(wl.B, gen)]
    starts = [s.to(dev) for s in wl.starts(wl.B, gen)]
    step.device_part(inputs, [s.clone() for s in starts])
    step.ensure_bucket()
    step.device_part(inputs, [s.clone() for s in starts])      # packs
    step.bucket.all_reduce(world)
    torch.cuda.synchronize()
    ref = step.bucket.flat.clone()
    step.overlap, step.comm = True, torch.cuda.Stream()
    step.bucket.flat.zero_()
    step.device_part(inputs, [s.clone() for s in starts])
    torch.cuda.synchronize()
    assert step._exchanged, "the boundary hook did not fire"
    got = step.bucket.flat
    err = float((got - ref).abs().max() / ref.abs().max())
    if rank == 0:
        print("%s: overlapped exchange vs single all-reduce: max abs difference %.3g of the largest gradient; early part %d of "
              "%d elements" % (key, err, step.bucket.split, step.bucket.numel))
    assert err < 1e-4
torch.distributed.destroy_process_group()
