"""Role timeline of knn_tc_kernel (CTA 0 of cloud 0, reference tiles 40..55) through mpc_debug_trace_buffer."""
import importlib, sys, torch, ctypes
sys.path.insert(0, '.')
mpc = importlib.import_module("markov-process-analysis-on-point-cloud_b200")
lib = mpc._lib.load()
g = torch.Generator().manual_seed(6)
x = torch.randn(8, 24000, 64, generator=g)
x = torch.nn.functional.leaky_relu(x @ torch.randn(64, 64, generator=g) / 8 + 0.3, 0.2).cuda().contiguous()
for _ in range(2):
    mpc.ops._knn_compute(8, x, x)
torch.cuda.synchronize()
tr = torch.zeros(4 * 16 * 2, dtype=torch.int64, device="cuda")
lib.mpc_debug_trace_buffer(ctypes.c_void_p(tr.data_ptr()))
mpc.ops._knn_compute(8, x, x)
torch.cuda.synchronize()
lib.mpc_debug_trace_buffer(ctypes.c_void_p(0))
t = tr.cpu().view(4, 16, 2)
t0 = int(t[t > 0].min())
print("tile | TMA: stage free | split: tile landed, split done | MMA: TMEM stage free, split seen (issue) | selection: accumulator ready, done")
for j in range(16):
    r = lambda a, b, c: int(t[a, b, c]) - t0
    print("%4d | %6d | %6d %6d | %6d %6d | %6d %6d" % (40 + j, r(0, j, 0), r(1, j, 0), r(1, j, 1), r(2, j, 0), r(2, j, 1), r(3, j, 0), r(3, j, 1)))
