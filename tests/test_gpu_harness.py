"""Vote-evaluation loops (R/tool/test_classification.py:114-162, R/tool/test_partseg.py:134-149) on the
GPU: the CUDA-graph-captured forward gives the same votes as eager execution, and the loop reproduces the reference
loop's arithmetic (cumulative in-place rescaling, mean of the per-vote predictions)."""
import argparse

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def synth(orc, specs, name):
    return orc.synthetic_state_dict([tuple(e) for e in specs[name]])


def test_vote_classify_graph_equals_eager(mpc, orc, golden_specs):
    h = mpc.harness
    m = mpc.task_models.Model(argparse.Namespace(num_point=1024, return_dist=True, cuda_ops=True, num_class=40))
    m.load_state_dict(synth(orc, golden_specs, "cls"))
    m = m.cuda().eval()
    gen = torch.Generator().manual_seed(4)
    pts = (torch.rand(4, 1024, 3, generator=gen) * 2 - 1).cuda()
    graphed = h.GraphedForward(m, [pts.permute(0, 2, 1).contiguous()], (1024, 512, 256, 128, 64))
    out = []
    for g in (None, graphed):
        np.random.seed(0)
        torch.manual_seed(0)  # FPS start draws
        out.append(h.vote_classify(m, pts.clone(), vote_num=3, graphed=g).cpu())
    torch.testing.assert_close(out[0], out[1], rtol=1e-4, atol=1e-4)
    # the loop's arithmetic, spelled out like the reference: cumulative rescaling, mean of the votes
    np.random.seed(0)
    torch.manual_seed(0)
    p, pool = pts.clone(), 0
    scale = h.PointcloudScale(0.95, 1.05)
    with torch.no_grad():
        for v in range(3):
            if v > 0:
                p = scale(p)
            pool = pool + m(p.permute(0, 2, 1).contiguous())
    torch.testing.assert_close(out[0], (pool / 3).cpu(), rtol=1e-5, atol=1e-5)
    acc, _ = h.classification_accuracy(out[0], out[0].argmax(1), 40)
    assert acc == 1.0


def test_vote_segment_graph_equals_eager(mpc, orc, golden_specs):
    h = mpc.harness
    m = mpc.task_models.get_model(50)
    m.load_state_dict(synth(orc, golden_specs, "seg"))
    m = m.cuda().eval()
    gen = torch.Generator().manual_seed(5)
    pts = (torch.rand(2, 2048, 3, generator=gen) * 2 - 1).cuda()
    label = torch.tensor([[4], [15]]).cuda()
    onehot = h.to_categorical(label, 16)
    graphed = h.GraphedForward(m, [pts.transpose(2, 1).contiguous(), onehot], (2048, 1024, 512, 256))
    out = []
    for g in (None, graphed):
        np.random.seed(1)
        torch.manual_seed(1)
        out.append(h.vote_segment(m, pts.clone(), label, num_votes=2, graphed=g).cpu())
    assert out[0].shape == (2, 2048, 50)
    # feature-space ties may resolve identically here (same kernels both ways): the two runs are the same arithmetic
    torch.testing.assert_close(out[0], out[1], rtol=1e-4, atol=1e-4)
    target = torch.randint(0, 50, (2, 2048), generator=gen)
    target[0, 0], target[1, 0] = 12, 47
    pred, correct, seen, ious = h.segmentation_metrics(out[0], target)
    assert pred.shape == (2, 2048) and seen == 4096 and len(ious["Chair"]) == 1 and len(ious["Table"]) == 1


def test_sampling_one_batch_ahead_reproduces_the_plain_step(mpc):
    """bench.py's graph-captured step with the FPS chain computed one batch ahead (ops.sampling_pyramid /
    ops.sampled_ahead) returns, one call later, exactly the loss the plain step computes for the same batch and start
    indices -- the sampling indices are the same, so the forward is the same."""
    import argparse
    import bench

    wl = bench.Workload("cls1024_train", 1)
    wl.B = 8
    dev = torch.device("cuda")
    step = bench.Step(wl, mpc, dev, 1)
    step.model.drop1.p = step.model.drop2.p = 0.0  # (the masks would differ from run to run)
    gen = torch.Generator().manual_seed(3)
    batches = [[t.to(dev) for t in wl.synth(wl.B, gen)] for _ in range(3)]
    starts = [[s.to(dev) for s in wl.starts(wl.B, gen)] for _ in range(3)]
    plain = []
    for b, s in zip(batches, starts):
        plain.append(float(step.device_part(b, [x.clone() for x in s]).detach()))
    # sampling_pyramid == the indices the forward computes itself
    rec = []
    with mpc.ops.index_tape(record=rec, fps_starts=[x.clone() for x in starts[0]]):
        wl.forward(step.model, batches[0])
    own = [i for k, i in rec if k == "fps"]
    ahead = mpc.ops.sampling_pyramid(batches[0][0].permute(0, 2, 1).contiguous(), wl.fps_npoints, starts[0])
    assert len(own) == len(ahead) == 5 and all(torch.equal(a, b) for a, b in zip(own, ahead))
    # one batch ahead: call i hands over batch i+1 and returns the result of batch i
    g = bench.GraphedStep(step, batches[0], starts[0], ahead=1)
    got = [float(g(batches[i], starts[i]).detach()) for i in (1, 2, 0)]
    torch.cuda.synchronize()
    np.testing.assert_allclose(got, [plain[0], plain[1], plain[2]], rtol=2e-5)
    # two batches ahead in two segments: results lag two calls (the priming batch is seen twice)
    del g
    g = bench.GraphedStep(step, batches[0], starts[0], ahead=2)
    got = [float(g(batches[i], starts[i]).detach()) for i in (1, 2, 0, 1)]
    torch.cuda.synchronize()
    np.testing.assert_allclose(got, [plain[0], plain[0], plain[1], plain[2]], rtol=2e-5)
