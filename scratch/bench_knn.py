import importlib, sys, torch
sys.path.insert(0, '.')
mpc = importlib.import_module("markov-process-analysis-on-point-cloud_b200")
ops = mpc.ops
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timeit(fn, n=7):
    for _ in range(2): fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort(); return ts[len(ts)//2] * 1e3
B = 32
for C in (3, 64):
    for (S, N) in [(2048, 2048), (1024, 2048), (1024, 1024), (512, 1024), (512, 512), (256, 512), (128, 2048), (256, 2048), (512, 2048), (128, 256)]:
        ref = torch.randn(B, N, C, device="cuda"); q = torch.randn(B, S, C, device="cuda")
        t = timeit(lambda: ops.knn_point(8, ref, q))
        fl = B * S * N * (2 * C + 3)
        print("knn C=%2d S=%4d N=%4d  %8.1f us   %6.2f TFLOP/s" % (C, S, N, t, fl / t / 1e6))
xyz = torch.rand(B, 2048, 3, device="cuda")
for (N, S) in [(2048, 1024), (1024, 512), (512, 256), (256, 128)]:
    x = xyz[:, :N].contiguous(); st = torch.zeros(B, dtype=torch.long, device="cuda")
    t = timeit(lambda: ops.farthest_point_sample(x, S, start=st))
    print("fps N=%4d S=%4d %8.1f us  %.3f us/iter" % (N, S, t, t / S))
