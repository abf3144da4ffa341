"""Feature-space kNN (k = 8) variants side by side: knob 4 = 1 (knn64 / knn64x2 / generic) vs 2 (register-tiled)."""
import importlib, sys, torch
sys.path.insert(0, '.')
mpc = importlib.import_module("markov-process-analysis-on-point-cloud_b200")
ops = mpc.ops; lib = mpc._lib.load()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timeit(fn, n=5):
    for _ in range(2): fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort(); return ts[len(ts)//2] * 1e3
for (B, S, N, C) in [(32, 2048, 2048, 64), (32, 1024, 2048, 64), (32, 1024, 1024, 64), (32, 512, 1024, 64), (32, 256, 512, 128),
                     (32, 512, 512, 128), (32, 128, 256, 256), (8, 24000, 24000, 64), (8, 12000, 24000, 64), (8, 6000, 12000, 64)]:
    ref = torch.randn(B, N, C, device="cuda"); q = torch.randn(B, S, C, device="cuda")
    fl = B * S * N * (2 * C + 3); row = []
    for knob in (1, 2):
        lib.mpc_debug_set_knob(4, knob)
        t = timeit(lambda: ops.knn_point(8, ref, q)); row.append("%9.1f us %6.2f TF/s" % (t, fl / t / 1e6))
    lib.mpc_debug_set_knob(4, 0)
    print("B=%2d S=%5d N=%5d C=%3d | old %s | tiled %s" % (B, S, N, C, row[0], row[1]), flush=True)
