"""Kernel timeline of one CUDA-graph replay of bench.py's step (CUPTI through torch.profiler): writes
gpurun_out/timeline_<tag>.csv with start us, duration us, stream, kernel name -- the concurrency picture that a
serialised launch list cannot give."""
import importlib, json, sys, os, torch
sys.path.insert(0, '.')
import bench
from torch.profiler import profile, ProfilerActivity
tag = sys.argv[1] if len(sys.argv) > 1 else "a"
mpc = importlib.import_module(bench.PKG)
mpc._lib.load()
dev = torch.device("cuda")
step = bench.Step(mpc, dev, 1)
gen = torch.Generator().manual_seed(1)
B = bench.B_PER_GPU
xyz, label, target = (t.to(dev) for t in bench.synth_batch(B, gen))
starts = lambda: [s.to(dev) for s in bench.fps_starts(B, gen)]
for _ in range(3):
    step(xyz, label, target, starts())
torch.cuda.synchronize()
g = bench.GraphedStep(step, xyz, label, target, starts())
for _ in range(3):
    g(xyz, label, target, starts())
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        g(xyz, label, target, starts())
    torch.cuda.synchronize()
path = "gpurun_out/timeline_%s.json" % tag
prof.export_chrome_trace(path)
ev = json.load(open(path))["traceEvents"]
rows = [(e["ts"], e["dur"], e.get("args", {}).get("stream", -1), e["name"]) for e in ev
        if e.get("cat") in ("kernel", "gpu_memset", "gpu_memcpy") and "dur" in e]
rows.sort()
# keep the last of the three back-to-back replays (clocks and caches in steady state)
firsts = [r[0] for r in rows if "Memcpy HtoD" in r[3] or "memcpy" in r[3].lower()]
n_per = len(rows) // 3
rows = rows[2 * n_per:]
t0 = rows[0][0]
with open("gpurun_out/timeline_%s.csv" % tag, "w") as f:
    for ts, dur, st, name in rows:
        f.write("%.3f,%.3f,%s,%s\n" % (ts - t0, dur, st, name.replace(",", ";")[:120]))
os.remove(path)
print("kernels", len(rows), "span us", rows[-1][0] + rows[-1][1] - t0)
