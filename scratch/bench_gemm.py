import importlib, sys, torch, ctypes
sys.path.insert(0, '.')
mpc = importlib.import_module("markov-process-analysis-on-point-cloud_b200")
ops = mpc.ops
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timeit(fn, n=10):
    for _ in range(3): fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort(); return ts[len(ts)//2] * 1e3
shapes = [(65536,64,64),(65536,64,128),(65536,192,64),(65536,256,256),(65536,896,512),(32768,64,64),(16384,64,64),(8192,128,128),(4096,256,256),(4096,64,256)]
which = sys.argv[1] if len(sys.argv) > 1 else "all"
for (M,K,N) in shapes:
    x = torch.randn(M,K,device="cuda"); w = torch.randn(N,K,device="cuda"); b = torch.randn(N,device="cuda"); gy = torch.randn(M,N,device="cuda")
    y = torch.empty(M,N,device="cuda"); gw = torch.empty(N,K,device="cuda")
    by = (M*K+M*N+N*K)*4; fl = 2*M*N*K
    if which in ("all","fwd"):
        t = timeit(lambda: ops._tc_gemm(x,w,b,y))
        t2 = timeit(lambda: torch.nn.functional.linear(x,w,b))
        print("fwd   M=%6d K=%4d N=%4d  %8.1f us  %7.1f GB/s  %6.1f TF/s   (cublas fp32 %8.1f us)" % (M,K,N,t,by/t/1e3,fl/t/1e6,t2))
    if which in ("all","wgrad"):
        t = timeit(lambda: mpc._lib.call("mpc_linear_wgrad_f32", mpc._lib.ptr(gy), ctypes.c_int64(N), mpc._lib.ptr(x), ctypes.c_int64(K), mpc._lib.ptr(gw), ctypes.c_int64(K), ctypes.c_int64(M), ctypes.c_int64(K), ctypes.c_int64(N), ctypes.c_int64(0)))
        t2 = timeit(lambda: gy.t().mm(x))
        print("wgrad M=%6d K=%4d N=%4d  %8.1f us  %7.1f GB/s  %6.1f TF/s   (cublas fp32 %8.1f us)" % (M,K,N,t,by/t/1e3,fl/t/1e6,t2))
