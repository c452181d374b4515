"""The reference's CPU path of the inference slice, restated with the SAME library calls the reference
makes (TEST INFRASTRUCTURE: bench.py's cpu_baseline / --impl reference legs and tests only).

/root/reference cannot travel to the GPU box, so this is a port ("kind": "port"), not the reference
files themselves; every step names the lines it restates.  It is multi-threaded exactly as the
reference would be with MODEL.DEVICE=cpu: torchvision's CPU roi_pool / nms kernels and ATen ops on all
host threads (torch.set_num_threads).
"""
import time

import torch
import torchvision  # noqa: F401  (registers torch.ops.torchvision.*)
from torchvision.ops.boxes import batched_nms as tv_batched_nms


def pool(features, rois, objectness, spatial_scale, P=7):
    # wsovod/modeling/poolers.py:183-186,277-284 -> torchvision RoIPool (CPU kernel)
    out, _ = torch.ops.torchvision.roi_pool(features, rois, spatial_scale, P, P)
    # wsovod/modeling/roi_heads/roi_heads.py:733-739
    return out * (objectness + 1).view(-1, 1, 1, 1)


def align_probs(x, text, temperature=50.0):
    # wsovod/modeling/class_heads/open_vocabulary_classifier.py:87-102 (projection excluded)
    w = torch.nn.functional.normalize(text.permute(1, 0).contiguous(), p=2, dim=0)
    xn = temperature * torch.nn.functional.normalize(x, p=2, dim=1)
    w = torch.cat([w, w.new_zeros((w.size(0), 1))], dim=1)
    logits = torch.mm(xn, w)
    # roi_heads/fast_rcnn_open_vocabulary.py:1034-1035
    return torch.softmax(logits, dim=-1)


def inference_single_image(boxes, scores, image_shape, score_thresh, nms_thresh, topk):
    # roi_heads/fast_rcnn_open_vocabulary.py:149-217 (class-agnostic boxes)
    valid = torch.isfinite(boxes).all(dim=1) & torch.isfinite(scores).all(dim=1)
    rows = torch.arange(scores.size(0))
    if not valid.all():
        boxes, scores, rows = boxes[valid], scores[valid], rows[valid]
    scores = scores[:, :-1]
    h, w = image_shape
    boxes = torch.stack([boxes[:, 0].clamp(0, w), boxes[:, 1].clamp(0, h),
                         boxes[:, 2].clamp(0, w), boxes[:, 3].clamp(0, h)], dim=-1)
    mask = scores > score_thresh
    inds = mask.nonzero()
    b, s = boxes[inds[:, 0]], scores[mask]
    keep = tv_batched_nms(b.float(), s, inds[:, 1], nms_thresh)   # detectron2.layers.batched_nms
    if topk >= 0:
        keep = keep[:topk]
    return b[keep], s[keep], inds[keep, 1], rows[inds[keep, 0]]


def run_slice(w, images=None, proposals=None):
    """One pass of pool -> align+softmax -> detections over a (sub)sample of workload `w`
    (wsovod_b200.synth.workload dict, CPU tensors).  Returns (seconds, proposals processed)."""
    N = w["N"] if images is None else min(images, w["N"])
    R = w["R"] if proposals is None else min(proposals, w["R"])
    feats = w["features"][:N]
    rois, emb, obj, sizes = [], [], [], []
    for n in range(N):
        r0 = w["offsets"][n]
        rois.append(w["rois"][r0:r0 + R])
        emb.append(w["region_emb"][r0:r0 + R])
        obj.append(w["objectness"][r0:r0 + R])
    rois, emb, obj = torch.cat(rois), torch.cat(emb), torch.cat(obj)
    t0 = time.perf_counter()
    pooled = pool(feats, rois, obj, w["spatial_scale"])
    probs = align_probs(emb, w["text_emb"], w["temperature"])
    dets = []
    for n in range(N):
        sl = slice(n * R, (n + 1) * R)
        hw = (float(w["image_sizes"][n, 0]), float(w["image_sizes"][n, 1]))
        dets.append(inference_single_image(rois[sl, 1:], probs[sl], hw, w["score_thresh"],
                                           w["nms_thresh"], w["topk"]))
    dt = time.perf_counter() - t0
    return dt, N * R, pooled, probs, dets
