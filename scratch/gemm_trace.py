import importlib, sys, torch, ctypes
sys.path.insert(0, '.')
mpc = importlib.import_module("markov-process-analysis-on-point-cloud_b200")
lib = mpc._lib.load()
M, K, N = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (65536, 64, 64)
x = torch.randn(M, K, device="cuda"); w = torch.randn(N, K, device="cuda"); b = torch.randn(N, device="cuda")
y = torch.empty(M, N, device="cuda")
for _ in range(3):
    mpc.ops._tc_gemm(x, w, b, y)
torch.cuda.synchronize()
tr = torch.zeros(1024, dtype=torch.int64, device="cuda")
lib.mpc_debug_trace_buffer(ctypes.c_void_p(tr.data_ptr()))
mpc.ops._tc_gemm(x, w, b, y)
torch.cuda.synchronize()
lib.mpc_debug_trace_buffer(ctypes.c_void_p(0))
t = tr.cpu().view(4, 256)
t0 = int(t[t > 0].min())
names = ["producer(start, then after each empty-wait)", "splitter(full-wait done, split done)...", "mma(tmem_empty ok | split-wait done...)", "epilogue(tmem_full ok, done)..."]
for r in range(4):
    v = [int(a) - t0 for a in t[r] if a > 0]
    print(names[r]); print("  ", v[:40])
