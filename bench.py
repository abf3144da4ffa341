"""Benchmark of the point-set hot path: point clouds/s on the workloads BASELINE.json's metric names.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workloads a,b,...] [--profile-ops FILE]

One run measures four workloads with the same method and prints ONE JSON line.  The headline record (top-level keys)
is `sem24k`; the others are complete sub-records under "sub", each with its own value / e2e / roofline /
cpu_baseline / clocks:

  sem24k          BASELINE configs[2]: S3DIS-shaped semantic segmentation, 24 000-point blocks through the
                  size-generalised Markov encoder/decoder, 13 classes, 8 blocks IN TOTAL sharded by batch over the GPUs
                  (8/4/2/1 per GPU: strong scaling), train mode, forward + backward (+ gradient all-reduce for N > 1).
  cls1024_train   configs[4]: data-parallel TRAINING STEP of the ModelNet40-shaped classifier, 32 clouds x 1024 points
                  per GPU (256 at 8 GPUs: weak), forward, label-smoothed loss, backward, gradient all-reduce, Adam.
  cls_fwd         configs[0]: classifier forward, 16 x 1024, eval() -- the reference's own CPU-runnable case.
  cls_fwd_bf16    the same forward through the bf16 inference path (ops.bf16_inference(): a separate dtype "bf16"
                  record; the headline and every other record stay fp32, the reference's precision).
  partseg2048     configs[1]: ShapeNetPart-shaped part segmentation, 32 x 2048 per GPU, forward + backward.

(configs[3], the op sweep, is scratch/op_sweep.py -> profiles/.)

Own arm, per workload:
  value        device-resident inputs; the step is captured once into a CUDA graph and replayed; CUDA events around
               every step, a 256 MiB buffer written between timed steps (L2 flush, outside the events); max over ranks.
               The FPS chain (a latency chain that depends on the coordinates only) is computed one batch ahead: the
               graph of step i samples batch i+1 on its own stream beside the forward/backward of batch i -- one
               sampling chain and one forward/backward per step, results lag (--fps-ahead 0: off; default 2: see GraphedStep).
  e2e          the same step through the public module call with the step's inputs copied from pinned host memory and
               the result (loss, or the logits of the forward-only workload) read back; copies inside the timed region.
  roofline     the dominant kernel of ours in that workload's step (most device time in one instrumented step),
               timed live with CUDA events on the launching stream: GEMM launches re-issued back to back in a CUDA
               graph; everything else bracketed in place during one single-stream eager step.  Algorithmic bytes /
               flops per launch from SURVEY.md 8d; peaks from MEASURED_PEAKS.json (HBM copy, bf16 dense) -- for the
               FP32-SIMT neighbour searches the peak is the FFMA issue limit SMs x 128 lanes x 2 x SM clock (no
               measured figure exists in that file; scratch/ubench/ffma_peak.cu measured 72 of the nominal 74.5
               TFLOP/s on this pool).
  cpu_baseline the CPU oracle (port of the reference path: torch CPU ops + the C restatement) on a bounded sample,
               rank 0 at N = 1 only.
--impl reference: /root/reference is pure Python and does not exist on the GPU box, so the reference arm times the
  oracle port with all host threads on a bounded sample of each workload (sem24k: ONE 24 000-point block per step).
  Under torchrun rank 0 alone runs it.

The sem24k model keeps the 16-channel object-label embedding (`keepHigh.conv7`, B rows) on running statistics:
BatchNorm over a per-GPU batch of ONE block is undefined (nn.BatchNorm1d raises, in the reference too), and the
workload must be the same network at 8, 4, 2 and 1 blocks per GPU.  Everything else runs in train mode.
"""
import argparse
import contextlib
import ctypes
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
PKG = "markov-process-analysis-on-point-cloud_b200"
UNIT = "clouds/s"
HEADLINE = "sem24k"
ORDER = ("sem24k", "cls1024_train", "cls_fwd", "cls_fwd_bf16", "partseg2048")
METRIC = ("point clouds/s (24 000-point blocks fwd+bwd, 8 blocks sharded over the GPUs; sub-records: 1024-point "
          "classifier training step and forward, 2048-point part segmentation fwd+bwd)")


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "sm_max_mhz": 1965.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms while a workload runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def mark(self):
        return time.perf_counter()

    def stop(self):
        if self.proc is None:
            return
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()

    def summary(self, t0=None, t1=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, ln in self.lines:
            if (t0 is not None and ts < t0) or (t1 is not None and ts > t1 + 0.2):
                continue
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax = float(parts[1])
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------
# workloads
# ------------------------------------------------------------------------------------------------------------
class Workload:
    """Static description of one workload: shapes, the model, the synthetic batch, the step's forward and its
    CPU-oracle twin.  `clouds_per_gpu` depends on the world size only for the strong-scaled sem24k."""

    def __init__(self, key, world):
        self.key = key
        self.world = world
        self.adam = False
        self.train = True
        self.eval_blocks = ()
        self.bf16 = False
        if key == "sem24k":
            # (MPC_BENCH_SEM_BLOCKS: experiments only, e.g. 1 block on 1 GPU = the per-GPU load of the 8-GPU run)
            self.total = int(os.environ.get("MPC_BENCH_SEM_BLOCKS", "8"))
            if self.total % world:
                raise SystemExit("bench.py: sem24k shards 8 blocks; --gpus must divide 8")
            self.B, self.N, self.classes, self.scaling = self.total // world, 24000, 13, "strong"
            self.kind, self.cfg_index = "seg", 2
            self.eval_blocks = ("keepHigh.conv7.",)
            self.title = ("S3DIS-shaped semantic segmentation, 24000-point blocks, batch 8 in total sharded by batch "
                          "over the GPUs, full Markov encoder + transition decoder, fwd+bwd (BASELINE configs[2])")
            self.cpu_B, self.cpu_steps = 1, 2
        elif key == "partseg2048":
            self.B, self.N, self.classes, self.scaling = 32, 2048, 50, "weak"
            self.kind, self.cfg_index = "seg", 1
            self.title = ("ShapeNetPart-shaped part segmentation, 32 x 2048 points per GPU, full Markov encoder + "
                          "transition decoder, fwd+bwd (BASELINE configs[1])")
            self.cpu_B, self.cpu_steps = 16, 2
        elif key == "cls1024_train":
            self.B, self.N, self.classes, self.scaling = 32, 1024, 40, "weak"
            self.kind, self.cfg_index = "cls", 4
            self.adam = True
            self.title = ("data-parallel training step of the ModelNet40-shaped classifier, 32 x 1024 points per GPU "
                          "(256 at 8 GPUs), fwd + smoothed loss + bwd + gradient all-reduce + Adam (BASELINE configs[4])")
            self.cpu_B, self.cpu_steps = 32, 3
        elif key in ("cls_fwd", "cls_fwd_bf16"):
            self.B, self.N, self.classes, self.scaling = 16, 1024, 40, "weak"
            self.kind, self.cfg_index = "cls", 0
            self.train = False
            self.bf16 = key.endswith("bf16")
            self.title = ("ModelNet40-shaped classification forward, batch 16 x 1024 points per GPU, eval() "
                          "(BASELINE configs[0])" + (", bf16 inference path (activations in bf16, fp32 accumulation, "
                                                     "indices from fp32 arithmetic)" if self.bf16 else ""))
            self.cpu_B, self.cpu_steps = 16, 5
        else:
            raise SystemExit("bench.py: unknown workload %r" % key)
        self.total_clouds = self.B * world
        if self.kind == "seg":
            self.fps_sizes = (self.N, self.N // 2, self.N // 4, self.N // 8)
        else:
            self.fps_sizes = (1024, 512, 256, 128, 64)
        self.fps_npoints = tuple(n // 2 for n in self.fps_sizes)  # every sampling step halves its state

    # ---- identical in both arms (the driver compares them)
    def config(self):
        return {"workload": self.title, "key": self.key, "points": self.N, "classes": self.classes,
                "clouds_total": self.total_clouds, "mode": "train" if self.train else "eval",
                "optimizer": "Adam lr 1e-3 wd 1e-4" if self.adam else None}

    def synth(self, B, gen):
        """Host tensors of one batch: (xyz [B,3,N], [label [B,1,16]], [target])."""
        xyz = torch.rand(B, 3, self.N, generator=gen) * 2 - 1
        if self.kind == "seg":
            label = torch.eye(16)[torch.randint(0, 16, (B,), generator=gen)].unsqueeze(1)
            target = torch.randint(0, self.classes, (B * self.N,), generator=gen)
            return [xyz, label, target]
        if self.train:
            return [xyz, torch.randint(0, self.classes, (B,), generator=gen)]
        return [xyz]

    def starts(self, B, gen):
        """The FPS start draws of one forward (R/modules/pointnet2_utils.py:96: randint(0, N_state, (B,)) per call)."""
        return [torch.randint(0, n, (B,), generator=gen, dtype=torch.long) for n in self.fps_sizes]

    # ---- own arm
    def build_model(self, mpc, device):
        torch.manual_seed(0)
        if self.kind == "seg":
            model = mpc.task_models.get_model(self.classes).to(device)
            self.loss_fn = mpc.task_models.get_loss()
        else:
            a = argparse.Namespace(num_point=self.N, return_dist=True, cuda_ops=True, num_class=self.classes)
            model = mpc.task_models.Model(a).to(device)
            self.loss_fn = mpc.task_models.SmoothClsLoss()
        model.train(self.train)
        for pre in self.eval_blocks:
            model.get_submodule(pre.rstrip(".")).eval()
        return model

    def forward(self, model, inputs):
        """The step's forward: the loss (train) or the logits (eval)."""
        if self.kind == "seg":
            xyz, label, target = inputs
            out, _ = model(xyz, label)
            return self.loss_fn(out.reshape(-1, self.classes), target, None)
        if self.train:
            return self.loss_fn(model(inputs[0]), inputs[1])
        return model(inputs[0])

    # ---- CPU oracle twin
    def cpu_state(self, mpc):
        torch.manual_seed(0)
        if self.kind == "seg":
            model = mpc.task_models.get_model(self.classes)  # only the source of a random-init state_dict (CPU)
        else:
            a = argparse.Namespace(num_point=self.N, return_dist=True, cuda_ops=True, num_class=self.classes)
            model = mpc.task_models.Model(a)
        return {k: v.clone().requires_grad_(self.train and v.dtype.is_floating_point and "running" not in k)
                for k, v in model.state_dict().items()}

    def cpu_forward(self, orc, P, inputs, starts):
        ctx = orc.Ctx(train=self.train, fps_starts=starts, eval_blocks=self.eval_blocks)
        if self.kind == "seg":
            out = orc.partseg_model(P, inputs[0], inputs[1], ctx)
            return orc.partseg_loss(out.reshape(-1, self.classes), inputs[2])
        if self.train:
            return orc.smooth_cls_loss(orc.cls_model(P, inputs[0], ctx), inputs[1])
        with torch.no_grad():
            return orc.cls_model(P, inputs[0], ctx)


# ------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on host cores, bounded sample
# ------------------------------------------------------------------------------------------------------------
def cpu_reference_run(wl, steps, warmup):
    from oracle import markov_oracle as orc

    orc.build()
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    mpc = importlib.import_module(PKG)
    P = wl.cpu_state(mpc)
    gen = torch.Generator().manual_seed(1)
    B = wl.cpu_B
    inputs = wl.synth(B, gen)
    opt = None
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        for p in P.values():
            p.grad = None
        res = wl.cpu_forward(orc, P, inputs, wl.starts(B, gen))
        if wl.train:
            res.backward()
            if wl.adam:
                if opt is None:
                    opt = torch.optim.Adam([p for p in P.values() if p.grad is not None], lr=1e-3, weight_decay=1e-4)
                opt.step()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    what = "fwd+bwd" + ("+Adam" if wl.adam else "") if wl.train else "fwd (eval)"
    return {"value": B / (ms / 1e3), "ms_per_step": ms, "cores": cores,
            "sample": "%d cloud%s x %d points per step, %s, %d steps after %d warm-up (oracle port: torch CPU ops + C "
                      "restatement, %d threads)" % (B, "" if B == 1 else "s", wl.N, what, steps, warmup, cores)}


def reference_record(wl, steps, warmup, args):
    r = cpu_reference_run(wl, steps, warmup)
    return {"impl": "reference", "metric": METRIC if wl.key == HEADLINE else "point clouds/s (%s)" % wl.key,
            "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": wl.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": wl.config(),
            "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                             "sample": r["sample"]},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}


def run_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    keys = selected(args)
    line = None
    sub = {}
    for key in keys:
        wl = Workload(key, max(1, args.gpus))
        if key == keys[0]:  # the headline honours --steps / --warmup; the sub-records stay short
            line = reference_record(wl, max(1, args.steps), max(1, min(args.warmup, 1)), args)
        else:
            sub[key] = reference_record(wl, min(max(1, args.steps), wl.cpu_steps), 1, args)
    if sub:
        line["sub"] = sub
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------
# own arm
# ------------------------------------------------------------------------------------------------------------
class Step:
    """One step of a workload on one GPU through the drop-in modules: gradients reset to None (autograd then writes
    each gradient straight into a fresh buffer instead of accumulating into a zeroed one), forward, and for training
    workloads loss + backward + the data-parallel gradient exchange (+ Adam)."""

    def __init__(self, wl, mpc, device, world):
        self.wl, self.mpc, self.world = wl, mpc, world
        self.model = wl.build_model(mpc, device)
        self.params = list(self.model.parameters())
        self.bucket = None   # dist.GradBucket, built after the first backward
        self.opt = None
        # Overlapped exchange (--graph-allreduce, N > 1): the parameters of everything that runs AFTER the encoder's last
        # stage in forward (classifier: la5, conv3/4, head = 80 % of the gradient bytes; part-seg: decoder + head) have
        # their gradients complete when backward reaches that stage's output; their part of the flat bucket is
        # all-reduced on a communication stream from a tensor hook there, beside the encoder's backward.
        self.overlap = False
        self.comm = None
        self._armed = False
        enc = ("keepHigh.start.", "keepHigh.la0.", "keepHigh.la1.", "keepHigh.la2.", "keepHigh.la3.", "keepHigh.la4.")
        self._late = {id(p) for n, p in self.model.named_parameters() if n.startswith(enc)}
        if wl.train:
            self.model.keepHigh.la4.register_forward_hook(self._boundary_forward_hook)

    def _boundary_forward_hook(self, module, inputs, output):
        feat = output[0]
        if self.overlap and self._armed and feat.requires_grad:
            feat.register_hook(self._boundary_grad_hook)

    def _boundary_grad_hook(self, grad):
        cur = torch.cuda.current_stream()
        self.comm.wait_stream(cur)
        wg = self.mpc.ops._wgrad_streams.get(cur.device)
        if wg is not None:
            self.comm.wait_stream(wg)  # deferred weight gradients of the early group are still being written there
        with torch.cuda.stream(self.comm):
            self.bucket.pack(0)
            self.bucket.all_reduce(self.world, 0)
        self._fired = True
        return None

    def device_part(self, inputs, starts, pack=True):
        wl = self.wl
        if wl.train:
            for p in self.params:
                p.grad = None
            self._armed = bool(pack and self.overlap and self.bucket is not None)
            self._fired = False
            with self.mpc.ops.index_tape(fps_starts=starts):
                res = wl.forward(self.model, inputs)
            res.backward()
            if self._armed and self._fired:
                # part 0 is on its way on the communication stream; the encoder's gradients follow on this one
                self.bucket.pack(1)
                self.bucket.all_reduce(self.world, 1)
                torch.cuda.current_stream().wait_stream(self.comm)
                self._exchanged = True
            else:
                self._exchanged = False
                if pack and self.bucket is not None and self.world > 1:
                    self.bucket.pack()
            self._armed = False
            return res
        mode = self.mpc.ops.bf16_inference() if wl.bf16 else contextlib.nullcontext()
        with torch.no_grad(), mode, self.mpc.ops.index_tape(fps_starts=starts):
            return wl.forward(self.model, inputs)

    def ensure_bucket(self):
        if self.wl.train and self.bucket is None:
            self.bucket = self.mpc.dist.GradBucket(self.params, early=lambda p: id(p) not in self._late)

    def exchange(self):
        """The path's one exchange step: mean of the gradients over the data-parallel ranks, one all-reduce of the
        flat fp32 bucket (16.5 MB part-seg / 34.1 MB classifier) with the 1/world folded into the reduction."""
        if self.wl.train and self.world > 1:
            if not getattr(self, "_exchanged", False):
                self.bucket.all_reduce(self.world)
            self.bucket.attach()

    def make_optimizer(self):
        used = [p for p in self.params if p.grad is not None]
        self.opt = torch.optim.Adam(used, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4, fused=True,
                                    capturable=True)
        self.opt.step()  # materialise the moment buffers outside any graph

    def __call__(self, inputs, starts):
        res = self.device_part(inputs, starts)
        self.exchange()
        if self.opt is not None:
            self.opt.step()
        return res


class GraphedStep:
    """The same step captured once into a CUDA graph (static input buffers, private memory pool) and replayed:
    the ~1000-1800 kernel launches per step otherwise make the step CPU-launch-bound.  The NCCL all-reduce runs
    between the step graph and the optimiser graph.

    Sampling one batch ahead (default; --no-fps-ahead switches it off): the FPS chain of a forward depends on nothing
    but the input coordinates and is a pure latency chain (22 500 sequential rounds per 24 000-point block), so the graph
    of step i also runs the chain for the batch of step i+1, on its own stream beside the forward / backward of batch i,
    and hands the indices over at the end (ops.sampling_pyramid / ops.sampled_ahead).  Every replay therefore still
    performs exactly one forward(+backward) and one full sampling chain's worth of rounds; what it returns belongs to a
    batch passed to an EARLIER call (`ahead` steps of pipeline latency, primed by the first batch)."""

    def __init__(self, step, inputs, starts, ahead=2, one_graph=False):
        self.step = step
        self.ahead = ahead
        # one_graph: the gradient all-reduce (NCCL, captured) and the optimiser step are part of the step graph -- no host
        # round trip between backward, exchange and optimiser
        self.one_graph = bool(one_graph) and step.wl.train and step.world > 1
        if self.one_graph:
            step.overlap = True
            step.comm = torch.cuda.Stream()
        wl, ops = step.wl, step.mpc.ops
        self.inputs = [t.clone() for t in inputs]
        self.starts = [s.clone() for s in starts]
        if ahead:
            # self.inputs / self.starts receive the batch handed in by a call; self.cur / self.pyr are what the step
            # consumes.  ahead = 1: the whole chain of the handed-in batch runs in this step.  ahead = 2: the chain is
            # cut after its first level (more than half of the rounds: 12 000 of 22 500 on a 24 000-point block); this
            # step runs the first level of the handed-in batch and the remaining levels of the batch handed in one call
            # earlier (self.mid*), so no step waits for more than the longer of the two segments.
            nl = len(wl.fps_npoints)
            self.cur = [t.clone() for t in inputs]
            self.pyr = [t.clone() for t in ops.sampling_pyramid(self._coords(self.cur), wl.fps_npoints, self.starts)]
            self.fps_stream = torch.cuda.Stream(priority=-1)
            if ahead >= 2:
                self.mid = [t.clone() for t in inputs]
                self.mid_starts = [s.clone() for s in starts]
                idx0, base1 = ops.sampling_pyramid(self._coords(self.mid), wl.fps_npoints, self.starts, levels=(0, 1))
                self.mid_idx0, self.mid_base = idx0[0].clone(), base1.clone()
                self.fps_stream2 = torch.cuda.Stream(priority=-1)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):  # warm the allocator / autograd on the capture stream
            for _ in range(2):
                self._part()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.res = self._part()
            if self.one_graph:
                step.exchange()
                if step.opt is not None:
                    step.opt.step()
        self.gopt = None
        if step.opt is not None and not self.one_graph:
            if step.world > 1:
                step.bucket.attach()  # the optimiser reads the averaged gradients from the bucket views
            self.gopt = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.gopt):
                step.opt.step()

    @staticmethod
    def _coords(inputs):
        return inputs[0].permute(0, 2, 1).contiguous()  # [B,3,N] -> [B,N,3], as the models do

    def _part(self):
        if not self.ahead:
            return self.step.device_part(self.inputs, self.starts)
        ops, wl = self.step.mpc.ops, self.step.wl
        nl = len(wl.fps_npoints)
        cur = torch.cuda.current_stream()
        self.fps_stream.wait_stream(cur)
        if self.ahead >= 2:
            self.fps_stream2.wait_stream(cur)
            with torch.cuda.stream(self.fps_stream):   # first level of the batch handed in by this call
                idx0, base1 = ops.sampling_pyramid(self._coords(self.inputs), wl.fps_npoints, self.starts, levels=(0, 1))
            with torch.cuda.stream(self.fps_stream2):  # remaining levels of the batch handed in one call earlier
                rest, _ = ops.sampling_pyramid(None, wl.fps_npoints, self.mid_starts, levels=(1, nl), base=self.mid_base)
            nxt = [self.mid_idx0] + rest
        else:
            with torch.cuda.stream(self.fps_stream):   # the whole chain of the handed-in batch
                nxt = ops.sampling_pyramid(self._coords(self.inputs), wl.fps_npoints, self.starts)
        with ops.sampled_ahead(self.pyr):
            res = self.step.device_part(self.cur, [])
        cur.wait_stream(self.fps_stream)
        if self.ahead >= 2:
            cur.wait_stream(self.fps_stream2)
            for d, s in zip(self.pyr, nxt):
                d.copy_(s)
            for d, s in zip(self.cur, self.mid):
                d.copy_(s)
            self.mid_idx0.copy_(idx0[0])
            self.mid_base.copy_(base1)
            for d, s in zip(self.mid, self.inputs):
                d.copy_(s)
            for d, s in zip(self.mid_starts, self.starts):
                d.copy_(s)
            for t in rest + idx0 + [base1]:
                t.record_stream(cur)
        else:
            for d, s in zip(self.pyr, nxt):
                d.copy_(s)
                s.record_stream(cur)
            for d, s in zip(self.cur, self.inputs):
                d.copy_(s)
        return res

    def __call__(self, inputs, starts):
        for d, s in zip(self.inputs, inputs):
            d.copy_(s, non_blocking=True)
        for d, s in zip(self.starts, starts):
            d.copy_(s, non_blocking=True)
        self.graph.replay()
        if self.one_graph:
            return self.res
        self.step.exchange()
        if self.gopt is not None:
            self.gopt.replay()
        return self.res


GEMM_ENTRIES = ("mpc_linear_fwd_f32", "mpc_linear_dgrad_f32", "mpc_linear_wgrad_f32", "mpc_linear_affine_act_f32",
                "mpc_linear_bf16")
# C-ABI entry point -> the CUDA kernel that does its work (the fp32 GEMM entry points share one kernel)
KERNEL_OF = {n: "linear_3xtf32_kernel" for n in GEMM_ENTRIES[:4]}
KERNEL_OF.update({"mpc_linear_bf16": "linear_bf16_kernel",
                  "mpc_knn3_grid_f32": "knn3_grid_kernel (uniform-grid coordinate search) + grid build",
                  "mpc_knn_f32": "knn3 / knn_tiled / knn64 kernels (FP32 SIMT, by shape)",
                  "mpc_knn_tc_f32": "knn_tc_kernel (tcgen05 filter + exact refinement)",
                  "mpc_fps_f32": "fps kernels (cta / cluster / grid by size)"})


def kernel_table(rows):
    """Aggregate the per-entry-point table by kernel; returns [(kernel, entry points, calls, ms, bytes)] by time."""
    agg = {}
    for name, n, ms, by in rows:
        k = KERNEL_OF.get(name, name)
        e = agg.setdefault(k, [set(), 0, 0.0, 0])
        e[0].add(name)
        e[1] += n
        e[2] += ms
        e[3] += by
    out = [(k, sorted(v[0]), v[1], v[2], v[3]) for k, v in agg.items()]
    out.sort(key=lambda r: -r[3])
    return out


def rebind_gemm_calls(mpc, calls, device):
    """The logged launches point at activations that die with their step; give every logged GEMM launch fresh
    buffers of the same shapes (random contents) so that the launch mix can be replayed on its own."""

    def val(a):
        return a.value if a.value is not None else 0

    def buf(rows, ld):
        return torch.randn(max(int(rows), 1), int(ld), dtype=torch.float32, device=device)

    out, keep = [], []
    P = mpc._lib.ptr
    for name, a, by in calls:
        if name == "mpc_linear_fwd_f32":
            ldx, ldw, ldy, rpg, M, K, N = (val(a[i]) for i in (1, 3, 6, 9, 10, 11, 12))
            x, w, y = buf(M, ldx), buf(N, ldw), buf(M, ldy)
            bias = torch.randn(N, device=device) if val(a[4]) else None
            sc = torch.zeros(2 * N + 2, dtype=torch.float64, device=device) if val(a[7]) else None
            gbias = torch.randn(max(M // max(rpg, 1), 1), N, device=device) if val(a[8]) else None
            keep += [x, w, y, bias, sc, gbias]
            na = (P(x), a[1], P(w), a[3], P(bias), P(y), a[6], P(sc), P(gbias), a[9], a[10], a[11], a[12])
        elif name == "mpc_linear_dgrad_f32":
            ldg, ldw, ldx, M, K, N = (val(a[i]) for i in (1, 3, 5, 6, 7, 8))
            g, w, x = buf(M, ldg), buf(N, ldw), buf(M, ldx)
            zb = torch.empty(max(val(a[10]), 4), dtype=torch.float32, device=device) if val(a[9]) else None
            keep += [g, w, x, zb]
            na = (P(g), a[1], P(w), a[3], P(x), a[5], a[6], a[7], a[8], P(zb), a[10])
        elif name == "mpc_linear_wgrad_f32":
            ldg, ldx, ldw, M, K, N = (val(a[i]) for i in (1, 3, 5, 6, 7, 8))
            g, x, w = buf(M, ldg), buf(M, ldx), buf(N, ldw)
            keep += [g, x, w]
            na = (P(g), a[1], P(x), a[3], P(w), a[5], a[6], a[7], a[8], a[9])
        elif name == "mpc_linear_affine_act_f32":
            ldx, ldw, ldr, ldy, M, K, N = (val(a[i]) for i in (1, 3, 8, 10, 11, 12, 13))
            x, w, y = buf(M, ldx), buf(N, ldw), buf(M, ldy)
            sc, sh = torch.rand(N, device=device) + 0.5, torch.randn(N, device=device)
            res = buf(M, ldr) if val(a[7]) else None
            keep += [x, w, y, sc, sh, res]
            na = (P(x), a[1], P(w), a[3], P(sc), P(sh), a[6], P(res), a[8], P(y), a[10], a[11], a[12], a[13])
        elif name == "mpc_linear_bf16":
            ldx, ldw, ldr, ldo, f32o, M, K, N = (val(a[i]) for i in (1, 3, 8, 10, 11, 12, 13, 14))
            b16 = lambda rows, ld: buf(rows, ld).to(torch.bfloat16)
            x, w = b16(M, ldx), b16(N, ldw)
            y = buf(M, ldo) if f32o else b16(M, ldo)
            sc = torch.rand(N, device=device) + 0.5 if val(a[4]) else None
            sh = torch.randn(N, device=device) if val(a[5]) else None
            res = b16(M, ldr) if val(a[7]) else None
            keep += [x, w, y, sc, sh, res]
            na = (P(x), a[1], P(w), a[3], P(sc), P(sh), a[6], P(res), a[8], P(y), a[10], a[11], a[12], a[13], a[14])
        else:
            return None, None
        out.append((name, na, by))
    return out, keep


def op_table(records):
    rows = []
    for name, recs in records.items():
        ms = sum(a.elapsed_time(b) for a, b, _ in recs)
        by = sum(c for _, _, c in recs)
        rows.append((name, len(recs), ms, by))
    rows.sort(key=lambda r: -r[2])
    return rows


def knn_flops(args):
    """mpc_knn_f32(ref, qry, dist, idx, B, N, S, C, L): B*S*N*(2C+3) flop."""
    B, N, S, C = (a.value for a in args[4:8])
    return B * S * N * (2 * C + 3)


def knn_tc_flops(args):
    """mpc_knn_tc_f32(ref, qry, dist, idx, ws, ws_bytes, B, N, S, C, K): (brute-force flops B*S*N*(2C+3) of the
    reference algorithm, tensor-core flops actually issued = 3 split terms x 2*B*S*N*C)."""
    B, N, S, C = (a.value for a in args[6:10])
    return B * S * N * (2 * C + 3), 3 * 2 * B * S * N * C


def measure_roofline(wl, step, mpc, device, inputs, starts_fn, flush, full, dev_ms_per_step):
    """Roofline record of the dominant kernel of ours in this workload's step (see the module docstring)."""
    pk, pk_kind = peaks()
    name, entries, n_calls, _, _ = kernel_table(full)[0]
    ops = mpc.ops
    if set(entries) <= set(GEMM_ENTRIES):
        # log every launch of the kernel during one eager step, then re-issue exactly those launches (same shapes,
        # fresh buffers) back to back inside a CUDA graph and time the replays on the launching stream
        mpc._lib.profiler = {"names": set(entries), "calls": []}
        step.device_part(inputs, starts_fn(), pack=False)
        torch.cuda.synchronize()
        calls = mpc._lib.profiler["calls"]
        mpc._lib.profiler = None
        calls, keep = rebind_gemm_calls(mpc, calls, device)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            mpc._lib.replay(calls)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        kgraph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(kgraph):
            mpc._lib.replay(calls)
        reps, ms = 5, 0.0
        for _ in range(reps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            kgraph.replay()
            b.record()
            torch.cuda.synchronize()
            ms += a.elapsed_time(b)
        ms /= reps
        by = sum(c[2] for c in calls)
        ideal_ms = 0.0  # every launch at the faster of its two rooflines (HBM bytes; 3 x 2MNK at half the bf16 rate)
        for cname, a, cby in calls:
            dims = [x.value for x in a if isinstance(x, ctypes.c_int64)]
            Mg, Kg, Ng = dims[-3:] if cname in ("mpc_linear_fwd_f32", "mpc_linear_affine_act_f32",
                                                "mpc_linear_bf16") else dims[-4:-1]
            # tensor-pipe floor: 3 split terms at half the bf16 rate (3xTF32), or one bf16 pass
            tens = (2.0 * Mg * Kg * Ng / (pk["bf16_tflops"] * 1e12) if cname == "mpc_linear_bf16"
                    else 3 * 2.0 * Mg * Kg * Ng / (pk["bf16_tflops"] * 0.5e12))
            ideal_ms += 1e3 * max(cby / (pk["hbm_gbs"] * 1e9), tens)
        n = len(calls)
        del kgraph, keep
        achieved = by / 1e9 / (ms / 1e3)
        return {"bound": "hbm", "kernel": name, "entry_points": entries, "achieved": achieved, "peak": pk["hbm_gbs"],
                "unit": "GB/s", "frac": achieved / pk["hbm_gbs"], "traffic": committed_traffic(name, wl.key),
                "peak_kind": pk_kind + " (burst copy)", "launches_per_step": n, "avg_launch_us": 1e3 * ms / max(n, 1),
                "algo_bytes_per_launch": by / max(n, 1), "kernel_ms_per_step": ms,
                "share_of_step": ms / dev_ms_per_step, "frac_mixed_bound": ideal_ms / ms,
                "frac_mixed_bound_how": "sum over the launches of max(bytes / HBM peak, tensor-pipe floor) divided by the "
                                        "measured time; tensor-pipe floor = 3 x 2MNK / (bf16 peak / 2) under the 3xTF32 "
                                        "split (the wide layers are tensor-pipe bound, not HBM bound), 2MNK / bf16 peak "
                                        "for the bf16 kernel",
                "how": "all %d launches of one step re-issued back to back in a CUDA graph, CUDA events around %d "
                       "replays; share_of_step relates that serialised time to the multi-stream graph step" % (n, reps)}
    # anything else: bracket every launch of the kernel in place, during one eager step with the side streams off
    # (so nothing else shares the GPU with it); these launches are 0.1-10 ms each, event overhead is noise
    streams_were = ops._STREAMS_ENABLED
    ops._STREAMS_ENABLED = False
    try:
        reps, ms, by, flops, n = 2, 0.0, 0, 0, 0
        for _ in range(reps):
            flush.zero_()
            mpc._lib.profiler = {"names": set(entries), "records": {}, "keep_args": True}
            step.device_part(inputs, starts_fn(), pack=False)
            torch.cuda.synchronize()
            recs = mpc._lib.profiler["records"]
            argl = mpc._lib.profiler.get("args", {})
            mpc._lib.profiler = None
            n = sum(len(v) for v in recs.values())
            ms += sum(a.elapsed_time(b) for v in recs.values() for a, b, _ in v)
            by = sum(c for v in recs.values() for _, _, c in v)
            flops = sum(knn_flops(a) for a in argl.get("mpc_knn_f32", []))
            tc = [knn_tc_flops(a) for a in argl.get("mpc_knn_tc_f32", [])]
        ms /= reps
    finally:
        ops._STREAMS_ENABLED = streams_were
    how = ("every launch of the kernel bracketed by CUDA events on its stream during %d single-stream eager steps "
           "(average); share_of_step relates that time to the multi-stream graph step" % reps)
    common = {"kernel": name, "entry_points": entries, "launches_per_step": n, "avg_launch_us": 1e3 * ms / max(n, 1),
              "kernel_ms_per_step": ms, "share_of_step": ms / dev_ms_per_step, "how": how,
              "traffic": committed_traffic(name, wl.key)}
    if entries == ["mpc_knn_tc_f32"]:
        issued = sum(t[1] for t in tc)
        tf32_peak = pk["bf16_tflops"] / 2
        achieved = issued / 1e12 / (ms / 1e3)
        rec = {"bound": "tensor", "achieved": achieved, "peak": tf32_peak, "unit": "TFLOP/s",
               "frac": achieved / tf32_peak, "algo_flops_per_launch": issued / max(n, 1),
               "peak_kind": pk_kind + " bf16 dense rate / 2 (kind::tf32 runs at half the bf16 rate)",
               "reference_algorithm_tflops": sum(t[0] for t in tc) / 1e12 / (ms / 1e3),
               "note": "issued tensor-core flops = 3 split terms x 2*B*S*N*C (3xTF32: fp32-level products so that the "
                       "exact refinement only has to look at ~K candidates); the launch also contains the operand "
                       "split pre-pass, the FP32 refinement and the exact fallback for near-tie queries; "
                       "reference_algorithm_tflops counts the reference's brute-force flops B*S*N*(2C+3) instead"}
        rec.update(common)
        return rec
    if entries == ["mpc_knn_f32"]:
        tc = False
        simt_peak = 148 * 128 * 2 * pk.get("sm_max_mhz", 1965.0) * 1e6 / 1e12
        achieved = flops / 1e12 / (ms / 1e3)
        rec = {"bound": "fp32-simt", "achieved": achieved, "peak": simt_peak, "unit": "TFLOP/s",
               "frac": achieved / simt_peak, "algo_flops_per_launch": flops / max(n, 1),
               "peak_kind": "FFMA issue limit 148 SMs x 128 lanes x 2 flop x sm_max_mhz (MEASURED_PEAKS.json holds no "
                            "FP32-SIMT figure; scratch/ubench/ffma_peak.cu measured 72 TFLOP/s on this pool)",
               "note": "brute-force distance flops B*S*N*(2C+3) of the reference algorithm (SURVEY 8d)"}
        if tc:
            rec["note"] += ("; the feature-space searches run their distance GEMM on the tensor cores (3xTF32 filter + "
                            "exact FP32 refinement), so the achieved figure may exceed the SIMT peak")
        rec.update(common)
        return rec
    if entries == ["mpc_fps_f32"]:
        rounds = sum(wl.fps_sizes[1:]) + (wl.fps_sizes[-1] // 2 if wl.kind == "seg" else 32)
        rec = {"bound": "latency", "achieved": 1e3 * ms / max(rounds, 1), "peak": None, "unit": "us/round",
               "frac": None, "note": "sequential chain: one distance update + global arg-max per sampled point; "
                                     "HBM traffic ~0 (SURVEY 8d)"}
        rec.update(common)
        return rec
    achieved = by / 1e9 / (ms / 1e3)
    rec = {"bound": "hbm", "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": achieved / pk["hbm_gbs"],
           "peak_kind": pk_kind + " (burst copy)", "algo_bytes_per_launch": by / max(n, 1)}
    rec.update(common)
    return rec


def committed_traffic(kernel, key):
    """dram__bytes_read + dram__bytes_write per launch of the dominant kernel from the committed ncu launch list of this
    command (profiles/r2_traffic.json: {workload: [{"kernel": name fragment, "dram_bytes_per_launch": ...}, ...]}),
    else null."""
    path = os.path.join(ROOT, "profiles", "r2_traffic.json")
    if not os.path.exists(path):
        return None
    with open(path) as f:
        entries = json.load(f).get(key) or []
    if isinstance(entries, dict):
        entries = [entries]
    for e in entries:
        if e.get("kernel") and e["kernel"] in kernel:
            return e.get("dram_bytes_per_launch")
    return None


def run_workload(wl, args, mpc, device, rank, local, world, sampler):
    gen = torch.Generator().manual_seed(1 + rank)
    B = wl.B
    host = [t.pin_memory() for t in wl.synth(B, gen)]
    inputs = [t.to(device) for t in host]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)  # > 126 MB L2

    def device_starts():
        return [s.pin_memory().to(device, non_blocking=True) for s in wl.starts(B, gen)]

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    t_begin = sampler.mark() if sampler else None
    step = Step(wl, mpc, device, world)
    # ---- warm-up (>= 3), then one fully instrumented step to find the dominant kernel of ours
    W = max(3, args.warmup)
    for i in range(W):
        step.device_part(inputs, device_starts())
        if i == 0:
            step.ensure_bucket()
            if wl.adam:
                step.make_optimizer()
        step.exchange()
    torch.cuda.synchronize()
    # (single-stream: with the side streams on, an entry point's event bracket also measures the kernels of other
    # streams it shares the GPU with, and the ranking would not be the kernels' own device time)
    streams_were = mpc.ops._STREAMS_ENABLED
    mpc.ops._STREAMS_ENABLED = False
    try:
        mpc._lib.profiler = {"names": None, "records": {}}
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        step.device_part(inputs, device_starts(), pack=False)
        ev1.record()
        torch.cuda.synchronize()
        full = op_table(mpc._lib.profiler["records"])
        instrumented_step_ms = ev0.elapsed_time(ev1)
    finally:
        mpc._lib.profiler = None
        mpc.ops._STREAMS_ENABLED = streams_were
    if not args.no_graph and args.fps_ahead:
        # the sampling chain runs beside the step, one or two batches ahead (GraphedStep): it is not what the timed step
        # waits for, so it does not compete for "dominant kernel of the step" (its device time stays in --profile-ops)
        ranked = [r for r in full if r[0] != "mpc_fps_f32"]
    else:
        ranked = full
    if args.profile_ops and rank == 0:
        with open(args.profile_ops, "a") as f:
            f.write("# %s: one instrumented eager step (%.3f ms incl. event overhead); per C-ABI entry point\n"
                    % (wl.key, instrumented_step_ms))
            f.write("%-34s %6s %10s %12s %9s\n" % ("entry point", "calls", "ms", "algo MB", "GB/s"))
            for name, n, ms, by in full:
                f.write("%-34s %6d %10.3f %12.2f %9.1f\n" % (name, n, ms, by / 1e6, by / 1e6 / max(ms, 1e-9)))
            f.write("%-34s %6s %10.3f\n\n" % ("sum of ours", "", sum(r[2] for r in full)))
    kernels_per_step = sum(mpc._lib.KERNELS_PER_CALL[n] * c for n, c, _, _ in full)

    run_step = step
    graphed = False
    if not args.no_graph:
        one_graph = not args.no_graph_allreduce
        try:
            run_step = GraphedStep(step, inputs, device_starts(), ahead=args.fps_ahead, one_graph=one_graph)
            graphed = True
        except Exception as e:  # noqa: BLE001 -- report and fall back
            print("bench.py: CUDA graph capture failed (%s: %s)" % (type(e).__name__, e), file=sys.stderr)
            torch.cuda.synchronize()
            step.overlap = False
            run_step = step
            if one_graph and world > 1:  # second attempt with the exchange outside the graph
                try:
                    run_step = GraphedStep(step, inputs, device_starts(), ahead=args.fps_ahead, one_graph=False)
                    graphed = True
                except Exception as e2:  # noqa: BLE001
                    print("bench.py: capture without the collective failed too (%s: %s); timing eager launches"
                          % (type(e2).__name__, e2), file=sys.stderr)
                    torch.cuda.synchronize()
                    run_step = step
    for _ in range(2):
        run_step(inputs, device_starts())

    # ---- timed region: device-resident inputs, per-step events, L2 flush between steps (outside the events)
    barrier()
    events = []
    t_timed0 = sampler.mark() if sampler else None
    if wl.key == args.ncu_workload:
        torch.cuda.profiler.start()  # ncu --profile-from-start off captures exactly this timed region
    for _ in range(args.steps):
        flush.zero_()
        starts = device_starts()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        run_step(inputs, starts)
        b.record()
        events.append((a, b))
    barrier()
    if wl.key == args.ncu_workload:
        torch.cuda.profiler.stop()
    dev_ms = sum(a.elapsed_time(b) for a, b in events)

    # ---- end to end: pinned host inputs -> H2D -> step -> result D2H, all inside the timed region
    def e2e_once():
        starts = [s.pin_memory() for s in wl.starts(B, gen)]
        res = run_step([t.to(device, non_blocking=True) for t in host],
                       [s.to(device, non_blocking=True) for s in starts])
        return res.detach().to("cpu")  # device -> host read of the step's result (synchronises)

    for _ in range(2):
        e2e_once()
    barrier()
    t_e2e = 0.0
    for _ in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res_host = e2e_once()
        t_e2e += time.perf_counter() - t0
    barrier()
    t_timed1 = sampler.mark() if sampler else None

    # ---- the same step with the sampling chain inside the forward it belongs to (no pipelining across steps), so the line
    # carries both numbers and the pipelined one can be read against it; headline workload only, outside the clock window
    inline_ms = 0.0
    if graphed and args.fps_ahead and wl.key == HEADLINE and world == 1 and not args.no_inline_sampling_arm:
        try:
            plain = GraphedStep(step, inputs, device_starts(), ahead=0, one_graph=getattr(run_step, "one_graph", False))
            for _ in range(2):
                plain(inputs, device_starts())
            barrier()
            pe = []
            for _ in range(args.steps):
                flush.zero_()
                starts = device_starts()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                plain(inputs, starts)
                b.record()
                pe.append((a, b))
            barrier()
            inline_ms = sum(a.elapsed_time(b) for a, b in pe)
            del plain
        except Exception as e:  # noqa: BLE001 -- the comparison arm is optional; the headline number does not depend on it
            print("bench.py: sampling-in-forward arm failed (%s: %s)" % (type(e).__name__, e), file=sys.stderr)
            torch.cuda.synchronize()
            inline_ms = 0.0

    t = torch.tensor([dev_ms, t_e2e * 1e3, inline_ms], dtype=torch.float64, device=device)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    dev_ms, e2e_ms, inline_ms = float(t[0]), float(t[1]), float(t[2])
    clouds = world * B * args.steps
    h2d = sum(t.numel() * t.element_size() for t in host) + sum(8 * B for _ in wl.fps_sizes)
    d2h = res_host.numel() * res_host.element_size()

    # ---- roofline of the dominant kernel (after the timed region: it re-runs eager steps)
    roof = measure_roofline(wl, step, mpc, device, inputs, device_starts, flush, ranked, dev_ms / args.steps) \
        if rank == 0 else None
    barrier()
    rec = None
    if rank == 0:
        par = "single GPU" if world == 1 else (
            "dp%d (batch shards; the flat fp32 gradient bucket all-reduced in two parts, the late layers' part beside the "
            "encoder's backward, captured in the step graph)" % world if (wl.train and getattr(run_step, "one_graph", False))
            else "dp%d (batch shards; one all-reduce of the flat fp32 gradient bucket)" % world if wl.train
            else "dp%d (batch shards; no collective in forward)" % world)
        rec = {
            "metric": METRIC if wl.key == HEADLINE else "point clouds/s (%s)" % wl.key,
            "value": clouds / (dev_ms / 1e3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": W,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": wl.scaling, "vs_baseline": None,
            "dtype": "bf16" if wl.bf16 else "f32", "data": "synthetic", "config": wl.config(),
            "method": {"clouds_per_gpu": B, "parallelism": par,
                       "l2": "256 MiB flush buffer written between timed steps; step working set >> 126 MB L2",
                       "launch": "CUDA graph replay" if graphed else "eager launches",
                       "sampling": ("the FPS chain of batch i+1 runs inside step i beside the forward/backward of batch i "
                                    "(one sampling chain and one forward/backward per step; results lag one step)"
                                    + ("; the chain is cut after its first level and the two segments of consecutive "
                                       "batches run side by side (results lag two steps)" if args.fps_ahead >= 2 else "")
                                    if graphed and args.fps_ahead else "inside the forward"),
                       "points_per_s": clouds * wl.N / (dev_ms / 1e3)},
            "e2e": {"value": clouds / (e2e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": kernels_per_step * args.steps,
            "roofline": roof,
            "clocks": sampler.summary(t_timed0, t_timed1) if sampler else None,
            "last_result": float(res_host.float().reshape(-1)[0]),
        }
        rec["clocks_whole_workload"] = sampler.summary(t_begin, t_timed1) if sampler else None
        if inline_ms > 0:
            rec["method"]["sampling_in_forward_arm"] = {
                "value": clouds / (inline_ms / 1e3), "unit": UNIT, "ms_per_step": inline_ms / args.steps,
                "note": "same step, same graph replay and timing rules, with every FPS chain inside the forward that uses it "
                        "(--fps-ahead 0); measured after the timed region, N=1 only"}
        if world == 1 and not args.no_cpu_baseline:
            r = cpu_reference_run(wl, wl.cpu_steps, 1)
            rec["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                                   "sample": r["sample"]}
    del run_step, step, inputs, flush
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    return rec


def selected(args):
    keys = [k for k in (args.workloads.split(",") if args.workloads else ORDER) if k]
    for k in keys:
        if k not in ORDER:
            raise SystemExit("bench.py: unknown workload %r (choose from %s)" % (k, ", ".join(ORDER)))
    return keys


def run_own(args):
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    # keep stdout to the one JSON line: NCCL prints its version banner there at NCCL_DEBUG=VERSION and =WARN, so file
    # descriptor 1 is pointed at stderr for everything but the final print
    json_out = os.fdopen(os.dup(1), "w")
    sys.stdout.flush()
    os.dup2(2, 1)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=device)
    mpc = importlib.import_module(PKG)
    mpc._lib.load()
    if os.environ.get("MPC_DEFER_WGRAD", "1") == "1":
        mpc.ops.set_defer_wgrad(True)  # gradients are read after backward only (bucket pack / optimiser)
    sampler = None
    if rank == 0:
        sampler = ClockSampler(local)
        sampler.start()
    keys = selected(args)
    line, sub = None, {}
    for key in keys:
        rec = run_workload(Workload(key, world), args, mpc, device, rank, local, world, sampler)
        if rank == 0:
            if key == keys[0]:
                line = rec
            else:
                sub[key] = rec
    if rank == 0:
        sampler.stop()
        if sub:
            line["sub"] = sub
        print(json.dumps(line), file=json_out, flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workloads", default=None, help="comma-separated subset of %s (first = headline record)"
                    % ",".join(ORDER))
    ap.add_argument("--profile-ops", default=None, help="append the per-entry-point device-time tables here")
    ap.add_argument("--ncu-workload", default=HEADLINE, help="workload whose timed region is bracketed by "
                    "cudaProfilerStart/Stop (for ncu --profile-from-start off)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of a CUDA graph replay")
    ap.add_argument("--no-graph-allreduce", action="store_true", help="N > 1: keep the NCCL gradient all-reduce and the "
                    "optimiser step outside the step graph (default: both are captured in it, and the all-reduce of the "
                    "late layers' gradients runs on a communication stream beside the encoder's backward)")
    ap.add_argument("--no-inline-sampling-arm", action="store_true", help="skip the extra N=1 measurement of the headline "
                    "step with the FPS chain inside the forward (reported as method.sampling_in_forward_arm)")
    ap.add_argument("--fps-ahead", type=int, default=2, choices=[0, 1, 2], help="software pipeline depth of the FPS chain "
                    "(see GraphedStep): 0 = inside the forward it belongs to, 1 = one batch ahead, 2 = two batches ahead "
                    "in two segments")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_own(args)


if __name__ == "__main__":
    main()
