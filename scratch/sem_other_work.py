"""Step time of the sem24k workload at B blocks per GPU WITHOUT its sampling chain (indices fed from a precomputed
pyramid): what the step costs when FPS is fully hidden -- the floor any deeper FPS pipeline could reach."""
import importlib, sys, torch
sys.path.insert(0, '.')
import bench
mpc = importlib.import_module(bench.PKG)
mpc._lib.load()
mpc.ops.set_defer_wgrad(True)
dev = torch.device("cuda")
for B in [int(a) for a in sys.argv[1:]] or [1, 2, 4, 8]:
    wl = bench.Workload("sem24k", 1)
    wl.B = B
    step = bench.Step(wl, mpc, dev, 1)
    gen = torch.Generator().manual_seed(1)
    inputs = [t.to(dev) for t in wl.synth(B, gen)]
    starts = [s.to(dev) for s in wl.starts(B, gen)]
    pyr = mpc.ops.sampling_pyramid(inputs[0].permute(0, 2, 1).contiguous(), wl.fps_npoints, starts)

    def part():
        with mpc.ops.sampled_ahead(pyr):
            return step.device_part(inputs, [])
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            part()
    torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        part()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ts = []
    for _ in range(8):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    print("sem24k, %d block(s) per GPU, sampling chain excluded: %.3f ms per step (median of 8)" % (B, ts[len(ts) // 2]))
    del g, step
    torch.cuda.empty_cache()
