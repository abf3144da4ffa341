import importlib, sys, torch
sys.path.insert(0, '.')
mpc = importlib.import_module("markov-process-analysis-on-point-cloud_b200")
lib = mpc._lib.load()
knob = int(sys.argv[1]) if len(sys.argv) > 1 else 2
lib.mpc_debug_set_knob(4, knob)
B, S, N, C = 32, 2048, 2048, 64
ref = torch.randn(B, N, C, device="cuda"); q = torch.randn(B, S, C, device="cuda")
for _ in range(2):
    mpc.ops.knn_point(8, ref, q)
torch.cuda.synchronize()
print("ok")
