"""Evaluation harness around the hot path: the reference's vote-evaluation loops and its
checkpoint format, on the drop-in models.

R = Markov_Process_Analysis_on_Point_Cloud/ in the reference tree.
  * vote_classify      R/tool/test_classification.py:114-162 (inner vote loop :129-146 and the accuracy bookkeeping)
  * vote_segment       R/tool/test_partseg.py:134-149 (vote loop) and :151-190 (per-category arg-max, accuracy, IoU)
  * PointcloudScale    R/tool/test_classification.py:68-79 == R/tool/test_partseg.py:56-67
  * save_checkpoint / load_checkpoint   R/tool/train_partseg.py:294-307 (the dict the reference's scripts write / read)

The forward of a vote loop runs with one fixed input shape many times, so it is captured once into a CUDA graph
(static input buffers) and replayed per vote; FPS start indices are drawn on the host exactly like the reference
(one torch.randint per sampling step and forward, CPU generator) and copied into the graph's static start buffers.
"""
import numpy as np
import torch

from . import ops

# R/tool/test_partseg.py:17-20
SEG_CLASSES = {'Earphone': [16, 17, 18], 'Motorbike': [30, 31, 32, 33, 34, 35], 'Rocket': [41, 42, 43],
               'Car': [8, 9, 10, 11], 'Laptop': [28, 29], 'Cap': [6, 7], 'Skateboard': [44, 45, 46], 'Mug': [36, 37],
               'Guitar': [19, 20, 21], 'Bag': [4, 5], 'Lamp': [24, 25, 26, 27], 'Table': [47, 48, 49],
               'Airplane': [0, 1, 2, 3], 'Pistol': [38, 39, 40], 'Chair': [12, 13, 14, 15], 'Knife': [22, 23]}
SEG_LABEL_TO_CAT = {label: cat for cat, labels in SEG_CLASSES.items() for label in labels}


class PointcloudScale:
    """Random anisotropic scaling of every cloud: three factors per cloud from np.random.uniform(low, high), drawn in
    the reference's order (one size-3 draw per cloud), applied in place to channels 0..2 of pc [B, N, >=3]."""

    def __init__(self, scale_low=2. / 3., scale_high=3. / 2.):
        self.scale_low = scale_low
        self.scale_high = scale_high

    def __call__(self, pc):
        scales = np.stack([np.random.uniform(low=self.scale_low, high=self.scale_high, size=[3])
                           for _ in range(pc.size(0))])
        pc[:, :, 0:3] *= torch.from_numpy(scales).float().to(pc.device).unsqueeze(1)  # one launch for the batch
        return pc


def to_categorical(y, num_classes):
    """R/tool/test_partseg.py:36-41: one-hot rows on y's device."""
    return torch.eye(num_classes, device=y.device)[y]


class GraphedForward:
    """model.eval() forward for one fixed input signature, captured into a CUDA graph.  `fps_sizes` are the point
    counts the model's sampling steps draw their start index from, in call order."""

    def __init__(self, model, example_inputs, fps_sizes):
        self.model = model.eval()
        self.inputs = [t.clone() for t in example_inputs]
        self.B = self.inputs[0].shape[0]
        self.fps_sizes = tuple(fps_sizes)
        dev = self.inputs[0].device
        self.starts = [torch.zeros(self.B, dtype=torch.long, device=dev) for _ in self.fps_sizes]
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(2):
                self._run()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.out = self._run()

    def _run(self):
        with ops.index_tape(fps_starts=self.starts):
            out = self.model(*self.inputs)
        return out[0] if isinstance(out, tuple) else out

    def __call__(self, *inputs):
        for d, s in zip(self.inputs, inputs):
            d.copy_(s, non_blocking=True)
        for d, n in zip(self.starts, self.fps_sizes):
            d.copy_(torch.randint(0, n, (self.B,), dtype=torch.long), non_blocking=True)  # same draw as the reference
        self.graph.replay()
        return self.out


def _forward(model, graphed, *inputs):
    if graphed is not None:
        return graphed(*inputs)
    with torch.no_grad():
        out = model(*inputs)
    return out[0] if isinstance(out, tuple) else out


def vote_classify(model, points, vote_num=10, pointscale=None, graphed=None):
    """points [B, N, C>=3] on the device (already sampled to the model's point count) -> mean over `vote_num` votes of
    the log-probabilities [B, num_class]; votes after the first see the clouds rescaled IN PLACE, cumulatively, exactly
    like the reference loop (:137-146).  `graphed`: optional GraphedForward built for input [B, C, N]."""
    pointscale = pointscale or PointcloudScale(scale_low=0.95, scale_high=1.05)
    model.eval()
    vote_pool = None
    for v in range(vote_num):
        if v > 0:
            points = pointscale(points)
        pred = _forward(model, graphed, points.permute(0, 2, 1).contiguous())
        vote_pool = pred.clone() if vote_pool is None else vote_pool + pred
    return vote_pool / vote_num


def classification_accuracy(pred, target, num_class):
    """(instance accuracy of the batch, per-class [correct fraction, seen]) as accumulated at :148-153."""
    choice = pred.argmax(1)
    class_acc = np.zeros((num_class, 2))
    tc, cc = target.cpu(), choice.cpu()
    for cat in np.unique(tc.numpy()):
        sel = tc == cat
        class_acc[cat, 0] += float((cc[sel] == tc[sel]).sum()) / float(sel.sum())
        class_acc[cat, 1] += 1
    return float((cc == tc).sum()) / float(tc.numel()), class_acc


def vote_segment(model, points, label, num_part=50, num_classes=16, num_votes=10, pointscale=None, graphed=None):
    """points [B, N, 3] on the device, label [B, 1] object category -> mean over votes of the part logits [B, N,
    num_part] (R/tool/test_partseg.py:138-149)."""
    pointscale = pointscale or PointcloudScale(scale_low=0.95, scale_high=1.05)
    model.eval()
    onehot = to_categorical(label.long(), num_classes)
    if onehot.dim() == 2:
        onehot = onehot.unsqueeze(1)
    vote_pool = None
    for v in range(num_votes):
        if v > 0:
            points = pointscale(points)
        pred = _forward(model, graphed, points.transpose(2, 1).contiguous(), onehot)
        vote_pool = pred.clone() if vote_pool is None else vote_pool + pred
    return vote_pool / num_votes


def segmentation_metrics(seg_pred, target):
    """Per-category restricted arg-max, accuracy and per-shape IoU of one batch (R/tool/test_partseg.py:151-190).
    seg_pred [B, N, 50] (host or device), target [B, N] part labels.  Returns (pred [B,N] as the reference computes it
    -- the arg-max INSIDE the category's part range, without the range offset, :159 --, correct, seen, {cat: [iou]})."""
    logits = seg_pred.detach().cpu().numpy()
    target = target.detach().cpu().numpy()
    B, N = target.shape
    pred = np.zeros((B, N), dtype=np.int32)
    shape_ious = {cat: [] for cat in SEG_CLASSES}
    for i in range(B):
        cat = SEG_LABEL_TO_CAT[int(target[i, 0])]
        pred[i] = np.argmax(logits[i][:, SEG_CLASSES[cat]], 1)  # (sic) no "+ seg_classes[cat][0]" in the reference
        part_ious = []
        for l in SEG_CLASSES[cat]:
            union = np.sum((target[i] == l) | (pred[i] == l))
            part_ious.append(1.0 if union == 0 else np.sum((target[i] == l) & (pred[i] == l)) / float(union))
        shape_ious[cat].append(float(np.mean(part_ious)))
    return pred, int(np.sum(pred == target)), B * N, shape_ious


def save_checkpoint(path, model, optimizer=None, **metrics):
    """The dictionary R/tool/train_partseg.py:297-306 writes ('model_state_dict', 'optimizer_state_dict' + metrics)."""
    state = dict(metrics)
    state['model_state_dict'] = model.state_dict()
    if optimizer is not None:
        state['optimizer_state_dict'] = optimizer.state_dict()
    torch.save(state, path)


def load_checkpoint(path, model, strict=True):
    """Load a checkpoint written by the reference's scripts (or by save_checkpoint) into a drop-in model: the
    state_dict keys and shapes are the reference's (743 classifier / 2189 part-seg)."""
    ckpt = torch.load(path, map_location="cpu", weights_only=False)
    sd = ckpt['model_state_dict'] if isinstance(ckpt, dict) and 'model_state_dict' in ckpt else ckpt
    model.load_state_dict(sd, strict=strict)
    return ckpt
