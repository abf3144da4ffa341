import importlib, sys, torch, ctypes
sys.path.insert(0, '.')
mpc = importlib.import_module("markov-process-analysis-on-point-cloud_b200")
from importlib import import_module
ops = mpc.ops
torch.manual_seed(0)
def run(M, K, N, pattern):
    if pattern == "rand":
        gy = torch.randn(M, N, device="cuda"); x = torch.randn(M, K, device="cuda")
    else:
        gy = torch.zeros(M, N, device="cuda"); x = torch.zeros(M, K, device="cuda")
        gy[:, 0] = 1.0; gy[:, 5] = 2.0; gy[:, 33 % N] = 3.0
        x[:, :] = torch.arange(K, device="cuda").float()[None, :] + 1
    gw = torch.full((N, K), -7.0, device="cuda")
    mpc._lib.call("mpc_linear_wgrad_f32", mpc._lib.ptr(gy), ctypes.c_int64(N), mpc._lib.ptr(x), ctypes.c_int64(K), mpc._lib.ptr(gw),
                  ctypes.c_int64(K), ctypes.c_int64(M), ctypes.c_int64(K), ctypes.c_int64(N), ctypes.c_int64(0))
    torch.cuda.synchronize()
    ref = gy.double().t() @ x.double()
    print("M,K,N", M, K, N, pattern, "maxerr", (gw.double() - ref).abs().max().item(), "ref max", ref.abs().max().item())
    print(" gw[0,:6]", gw[0, :6].tolist(), "\n ref[0,:6]", ref[0, :6].tolist())
    print(" gw[5,:4]", gw[5, :4].tolist(), " ref", ref[5, :4].tolist())
    print(" gw[1,:4]", gw[1, :4].tolist(), " nnz", int((gw != 0).sum()), "of", gw.numel())
run(32, 32, 64, "pat")
run(32, 32, 64, "rand")
run(128, 32, 64, "pat")
run(4096, 64, 128, "rand")
