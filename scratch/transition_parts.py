"""Device time of the two halves of the Markov transition (reverse-list build / apply) on one large cloud."""
import importlib, sys, torch
sys.path.insert(0, '.')
mpc = importlib.import_module("markov-process-analysis-on-point-cloud_b200")
ops = mpc.ops
from ctypes import c_int64 as i64
ptr = mpc._lib.ptr
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timeit(fn, n=5):
    fn(); ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]
for N in [int(a) for a in sys.argv[1:]] or [262144, 1048576]:
    B, S, C, K = 1, N // 4, 64, 16
    g = torch.Generator().manual_seed(N)
    xyz = (torch.rand(B, N, 3, generator=g) * 2 - 1).cuda()
    sub = xyz[:, ::4].contiguous()
    _, idx = ops.knn_point(K, xyz, sub)
    pts = torch.randn(B, S, C, generator=g).cuda()
    ws = torch.empty(B * (2 * N + 1) + B * S * K, dtype=torch.int32, device="cuda")
    out = torch.empty(B, N, C, device="cuda"); cnt = torch.empty(B, N, device="cuda")
    tb = timeit(lambda: mpc._lib.call("mpc_transition_csr_build", ptr(idx), ptr(ws), i64(B), i64(S), i64(K), i64(N)))
    ta = timeit(lambda: mpc._lib.call("mpc_transition_csr_apply_f32", ptr(pts), ptr(ws), ptr(out), ptr(cnt), i64(B), i64(S), i64(K), i64(C), i64(N)))
    mpc._lib.load().mpc_debug_set_knob(7, 1)
    to = timeit(lambda: mpc._lib.call("mpc_transition_csr_apply_f32", ptr(pts), ptr(ws), ptr(out), ptr(cnt), i64(B), i64(S), i64(K), i64(C), i64(N)))
    mpc._lib.load().mpc_debug_set_knob(7, 0)
    ts = timeit(lambda: mpc._lib.call("mpc_transition_fwd_f32", ptr(pts), ptr(idx), ptr(out), ptr(cnt), i64(B), i64(S), i64(K), i64(C), i64(N)))
    by = B * ((S + N) * C * 4 + S * K * 8)
    deg = torch.bincount(idx.reshape(-1), minlength=N)
    print("N %d: build %.3f ms, apply (group kernel) %.3f ms = %.0f GB/s, apply (per-thread kernel) %.3f ms, scatter form %.3f ms; "
          "list length mean %.2f max %d" % (N, tb, ta, by / ta / 1e6, to, ts, float(deg.float().mean()), int(deg.max())))
