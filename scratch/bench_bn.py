"""Launch-geometry sweep of the streaming BatchNorm kernels (mpc_bn_act_fwd_sums_f32, mpc_bn_act_bwd_f32) at the
shapes of the part-seg step: 24 launches back to back in a CUDA graph over 6 rotating buffer sets, CUDA events."""
import importlib, sys, ctypes, torch
sys.path.insert(0, '.')
mpc = importlib.import_module("markov-process-analysis-on-point-cloud_b200")
lib = mpc._lib.load()
P, I64, F32 = mpc._lib.ptr, ctypes.c_int64, ctypes.c_float
dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
shapes = [(65536, 64), (32768, 64), (16384, 64), (8192, 128), (4096, 256), (65536, 128), (65536, 512)]


def graph_time(fn_sets, reps=24):
    seq = [fn_sets[i % len(fn_sets)] for i in range(reps)]
    for f in seq[:6]:
        f()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for f in seq:
            f()
    ts = []
    for _ in range(5):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[2] * 1e3 / reps


def make(M, C, kind):
    y = torch.randn(M, C, device=dev); go = torch.randn(M, C, device=dev); out = torch.empty(M, C, device=dev)
    gamma = torch.rand(C, device=dev) + 0.5; beta = torch.randn(C, device=dev)
    mean = y.mean(0); var = y.var(0, unbiased=False)
    stats = torch.empty(2 * C, device=dev); gg = torch.empty(C, device=dev); gb = torch.empty(C, device=dev)
    scratch = torch.zeros(2 * C + 2, dtype=torch.float64, device=dev)
    sums = torch.zeros(2 * C + 2, dtype=torch.float64, device=dev)
    keep = (y, go, out, gamma, beta, mean, var, stats, gg, gb, scratch, sums)
    if kind == "fwd":
        def f(keep=keep):
            # sums are consumed (cleared) by the kernel: statistics are garbage after the first call, timing is not
            mpc._lib.call("mpc_bn_act_fwd_sums_f32", P(y), P(sums), P(gamma), P(beta), F32(1e-5), F32(0.2), P(None),
                          P(out), P(stats), P(None), P(None), P(None), F32(0.1), I64(M), I64(C))
    else:
        def f(keep=keep):
            mpc._lib.call("mpc_bn_act_bwd_f32", P(go), P(y), P(mean), P(var), P(gamma), P(beta), F32(1e-5), F32(0.2),
                          ctypes.c_int(1), P(out), P(gg), P(gb), P(scratch), P(None), I64(0), I64(C), I64(M), I64(C))
    return f


settings = [(0, 0, 0), (8, 0, 0), (8, 0, 2), (8, 0, 4), (4, 0, 4), (16, 0, 4), (0, 2, 0), (0, 4, 0), (0, 1, 0),
            (8, 2, 4), (8, 4, 4)]
print("%-18s" % "knobs(ew,colred,f4)" + "".join("%16s" % ("%dx%d" % s) for s in shapes))
for kind in ("fwd", "bwd"):
    fns = {s: [make(s[0], s[1], kind) for _ in range(6)] for s in shapes}
    for k in settings:
        if kind == "fwd" and k[1]:
            continue
        for i, v in enumerate(k):
            lib.mpc_debug_set_knob(i, v)
        row = [graph_time(fns[s]) for s in shapes]
        print("%s %-14s" % (kind, k) + "".join("%10.1f us   " % t for t in row), flush=True)
    ideal = [(2 if kind == "fwd" else 3) * s[0] * s[1] * 4 / 6531.9e3 for s in shapes]
    print("%s %-14s" % (kind, "ideal") + "".join("%10.1f us   " % t for t in ideal))
for i in range(3):
    lib.mpc_debug_set_knob(i, 0)
