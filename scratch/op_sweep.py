"""BASELINE config 4: op microbench sweep on single large clouds (B=1; B=8 at 16K), S = N/4 queries, C = 64 features.
Prints one line per (op, N): device time (CUDA events, L2 flushed), algorithmic GB/s or TFLOP/s (SURVEY 8d formulas).
Under torchrun (N GPUs) every rank sweeps its own cloud -- the ops shard by cloud, FPS does not shard inside one -- and
rank 0 prints the slowest rank's time and the aggregate rate (weak scaling)."""
import importlib, os, sys, torch
sys.path.insert(0, '.')
RANK, WORLD = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
if WORLD > 1:
    torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0))))
mpc = importlib.import_module("markov-process-analysis-on-point-cloud_b200")
ops = mpc.ops
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timeit(fn, n=3):
    fn(); ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort(); t = ts[len(ts) // 2]
    if WORLD > 1:  # slowest rank
        tt = torch.tensor([t], device="cuda"); torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX); t = float(tt)
    return t / WORLD  # aggregate rate = WORLD clouds per t
sizes = [int(a) for a in sys.argv[1:]] or [16384, 65536, 262144, 1048576]
_print = print
print = lambda *a, **k: _print(*a, **k) if RANK == 0 else None
if WORLD > 1: print("# %d GPUs, one cloud set per GPU; ms = slowest rank / %d (time per GPU-cloud in aggregate)" % (WORLD, WORLD))
print("%-22s %8s %3s %10s %12s" % ("op", "N", "B", "ms", "rate"))
for N in sizes:
    B = 8 if N <= 16384 else 1
    S, C = N // 4, 64
    g = torch.Generator().manual_seed(N + RANK)
    xyz = (torch.rand(B, N, 3, generator=g) * 2 - 1).cuda()
    feat = torch.randn(B, N, C, generator=g).cuda()
    start = torch.zeros(B, dtype=torch.long, device="cuda")
    t = timeit(lambda: ops.farthest_point_sample(xyz, S, start=start), n=1 if N > 65536 else 3)
    print("%-22s %8d %3d %10.3f %9.2f us/round  (%.2f G point-updates/s)" % ("fps N->N/4", N, B, t, 1e3 * t / S, B * S * N / t / 1e6))
    fidx = ops.farthest_point_sample(xyz, S, start=start)
    sub = ops.index_points(xyz, fidx)
    for K in (16, 32):
        t = timeit(lambda: ops.knn_point(K, xyz, sub), n=1 if N > 65536 else 3)
        print("%-22s %8d %3d %10.3f %9.2f TFLOP/s brute-force-equivalent (grid search: ~7k of N distances per query "
              "evaluated; %.1f M queries/s)" % ("knn k=%d (S=N/4)" % K, N, B, t, B * S * N * 9 / t / 1e9, B * S / t / 1e3))
    _, idx = ops.knn_point(16, xyz, sub)
    t = timeit(lambda: ops.index_points(feat, idx))
    by = B * S * 16 * (8 * C + 8)
    print("%-22s %8d %3d %10.3f %9.1f GB/s" % ("group [S,16,64]", N, B, t, by / t / 1e6))
    subf = ops.index_points(feat, fidx)
    t = timeit(lambda: ops.upsample(subf, idx, n_out=N))
    by = B * ((S + N) * C * 4 + S * 16 * 8)
    print("%-22s %8d %3d %10.3f %9.1f GB/s" % ("transition S->N k=16", N, B, t, by / t / 1e6))
    w = torch.randn(64, 64, device="cuda"); bias = torch.randn(64, device="cuda")
    y = torch.empty(B * N, 64, device="cuda")
    t = timeit(lambda: ops._tc_gemm(feat.view(-1, 64), w, bias, y))
    by = (B * N * 128 + 64 * 64) * 4
    print("%-22s %8d %3d %10.3f %9.1f GB/s" % ("linear 64->64 [N,64]", N, B, t, by / t / 1e6))
