"""Per-shape table of the GEMM launches of one training step of a bench.py workload (default partseg2048; argv[1] =
workload key): every distinct (entry point, M, K, N) is re-issued 24 times back to back in its own CUDA graph over 6
rotating buffer sets and timed with CUDA events.  Prints count, us per launch, algorithmic GB/s and the share of the
step's GEMM time."""
import importlib, sys, torch
sys.path.insert(0, '.')
import bench
mpc = importlib.import_module(bench.PKG)
mpc._lib.load()
dev = torch.device("cuda")
wl = bench.Workload(sys.argv[1] if len(sys.argv) > 1 else "partseg2048", 1)
step = bench.Step(wl, mpc, dev, 1)
gen = torch.Generator().manual_seed(1)
inputs = [t.to(dev) for t in wl.synth(wl.B, gen)]
starts = lambda: [s.to(dev) for s in wl.starts(wl.B, gen)]
for _ in range(2):
    step.device_part(inputs, starts())
torch.cuda.synchronize()
names = {"mpc_linear_fwd_f32", "mpc_linear_dgrad_f32", "mpc_linear_wgrad_f32"}
mpc._lib.profiler = {"names": names, "calls": []}
step.device_part(inputs, starts())
torch.cuda.synchronize()
calls = mpc._lib.profiler["calls"]
mpc._lib.profiler = None
groups = {}
for c in calls:
    name, a, by = c
    v = lambda i: a[i].value if a[i].value is not None else 0
    if name == "mpc_linear_fwd_f32":
        key = (name[11:-4], v(10), v(11), v(12), 1 if v(7) else 0)
    else:
        key = (name[11:-4], v(6), v(7), v(8), 0)
    groups.setdefault(key, []).append(c)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
rows = []
for key, cs in groups.items():
    sets, keep = [], []
    for _ in range(6):
        r, k = bench.rebind_gemm_calls(mpc, cs[:1], dev)
        sets.append(r[0]); keep.append(k)
    seq = [sets[i % 6] for i in range(24)]
    mpc._lib.replay(seq)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(g):
            mpc._lib.replay(seq)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    us = ts[2] * 1e3 / 24
    rows.append((key, len(cs), us, cs[0][2]))
    del g, keep, sets
tot = sum(n * us for _, n, us, _ in rows)
rows.sort(key=lambda r: -r[1] * r[2])
print("%-6s %7s %5s %5s %2s %4s %8s %8s %8s %6s" % ("kind", "M", "K", "N", "st", "cnt", "us", "ideal", "GB/s", "share"))
for (kind, M, K, N, st), n, us, by in rows:
    print("%-6s %7d %5d %5d %2d %4d %8.2f %8.2f %8.0f %5.1f%%" % (kind, M, K, N, st, n, us, by / 6531.9e3, by / us / 1e3,
                                                                 100 * n * us / tot))
print("total %.1f us over %d launches" % (tot, sum(r[1] for r in rows)))
