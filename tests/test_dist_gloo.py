"""Multi-rank host logic on CPU (gloo, world_size 2): batch sharding and the gradient exchange step."""
import importlib
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
PKG = "markov-process-analysis-on-point-cloud_b200"


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mpc = importlib.import_module(PKG)
    torch.manual_seed(0)
    lin = torch.nn.Linear(5, 3)
    unused = torch.nn.Linear(2, 2)  # never receives a gradient, like the reference's dead sub-modules
    data = torch.arange(40, dtype=torch.float32).reshape(8, 5) / 10.0
    lo, hi = mpc.dist.shard_batch(8, rank, world)
    lin(data[lo:hi]).pow(2).sum().backward()
    local = [p.grad.clone() for p in lin.parameters()]
    # flat-bucket form of the exchange, on clones of the same gradients
    twin = torch.nn.Linear(5, 3)
    twin_unused = torch.nn.Linear(2, 2)
    for p, g in zip(twin.parameters(), local):
        p.grad = g.clone()
    bucket = mpc.dist.GradBucket(list(twin.parameters()) + list(twin_unused.parameters()))
    bucket.exchange(world)
    # two-part exchange (the bias as the "early" group): part 0 first, part 1 later, same result
    twin2 = torch.nn.Linear(5, 3)
    for p, g in zip(twin2.parameters(), local):
        p.grad = g.clone()
    b2 = mpc.dist.GradBucket(list(twin2.parameters()), early=lambda p: p.dim() == 1)
    b2.pack(0)
    b2.all_reduce(world, 0)
    b2.pack(1)
    b2.all_reduce(world, 1)
    b2.attach()
    n = mpc.dist.allreduce_mean_grads(list(lin.parameters()) + list(unused.parameters()), world)
    torch.save({"n": n, "local": local, "avg": [p.grad.clone() for p in lin.parameters()], "shard": (lo, hi),
                "bucket": [p.grad.clone() for p in twin.parameters()], "bucket_numel": bucket.numel,
                "bucket_views": all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(bucket.params, bucket.views)),
                "bucket_skipped_unused": all(p.grad is None for p in twin_unused.parameters()),
                "two_part": [p.grad.clone() for p in twin2.parameters()], "two_part_split": (b2.split, b2.n_early,
                                                                                              b2.params[0].dim())},
               os.path.join(out_dir, "r%d.pt" % rank))
    dist.destroy_process_group()


def test_shard_batch_covers_everything(mpc):
    for B in (1, 7, 8, 32, 33):
        for world in (1, 2, 3, 8):
            spans = [mpc.dist.shard_batch(B, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_gradient_exchange_gloo_world2(tmp_path):
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = (torch.load(os.path.join(str(tmp_path), "r%d.pt" % r)) for r in (0, 1))
    assert r0["shard"] == (0, 4) and r1["shard"] == (4, 8)
    assert r0["n"] == 2 and r1["n"] == 2  # the two parameters with a gradient; the unused module is skipped
    for a0, a1, l0, l1 in zip(r0["avg"], r1["avg"], r0["local"], r1["local"]):
        assert torch.equal(a0, a1)  # every rank ends with the same averaged gradient ...
        torch.testing.assert_close(a0, (l0 + l1) / 2)  # ... the mean of the per-shard gradients
    # the flat bucket gives the same averaged gradients, as views of ONE buffer, unused parameters left out
    for r in (r0, r1):
        assert r["bucket_numel"] == 5 * 3 + 3 and r["bucket_views"] and r["bucket_skipped_unused"]
        for b, a in zip(r["bucket"], r0["avg"]):
            torch.testing.assert_close(b, a)
        # the two-part layout puts the early group (here: the bias, 3 elements) first and exchanges to the same result
        assert r["two_part_split"] == (3, 1, 1)
        for b, a in zip(r["two_part"], r0["avg"]):
            torch.testing.assert_close(b, a)
