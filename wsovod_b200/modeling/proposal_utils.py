"""RPN proposal selection: find_top_rpn_proposals of the reference
(wsovod/modeling/proposal_generator/proposal_utils.py:26-144), batched.

The reference walks the images one by one and calls batched_nms (levels as groups) per image.  Here the
whole batch goes through ONE call of the NMS kernel with (image, level) as the group -- kernel 4's
`batched_nms` returns its survivors sorted by score, so the per-image lists fall out of one stable split.
Same results: per level the `pre_nms_topk` best anchors (the same `sort(descending=True)` call, so ties fall
the same way), non-finite rows dropped (training: FloatingPointError), clip, boxes with a side <=
`min_box_size` dropped, NMS at `nms_thresh`, the first `post_nms_topk` per image.
"""
from typing import List, Tuple

import torch

from .. import ops
from ..structures import Boxes, Instances


def find_top_rpn_proposals(proposals: List[torch.Tensor], pred_objectness_logits: List[torch.Tensor],
                           image_sizes: List[Tuple[int, int]], nms_thresh: float, pre_nms_topk: int,
                           post_nms_topk: int, min_box_size: float, training: bool, nms_fn=None):
    """proposals: L tensors (N, Hi*Wi*A, 4); pred_objectness_logits: L tensors (N, Hi*Wi*A).  Returns N
    Instances with `proposal_boxes`, `objectness_logits`, score-descending.  `nms_fn(boxes, scores, groups,
    thresh) -> kept indices sorted by score` defaults to the CUDA kernel (ops.batched_nms)."""
    nms_fn = ops.batched_nms if nms_fn is None else nms_fn
    N = len(image_sizes)
    dev = proposals[0].device
    rows = torch.arange(N, device=dev)[:, None]
    boxes, scores, levels = [], [], []
    for lvl, (p, s) in enumerate(zip(proposals, pred_objectness_logits)):          # :75-101
        k = min(s.shape[1], pre_nms_topk)
        s_sorted, order = s.sort(descending=True, dim=1)
        boxes.append(p[rows, order[:, :k]])
        scores.append(s_sorted[:, :k])
        levels.append(torch.full((k,), lvl, dtype=torch.int64, device=dev))
    boxes, scores, levels = torch.cat(boxes, 1), torch.cat(scores, 1), torch.cat(levels)   # (N,T,4) (N,T) (T)
    T = levels.numel()
    b, s = boxes.reshape(N * T, 4), scores.reshape(N * T)
    img = torch.arange(N, device=dev).repeat_interleave(T)
    ok = torch.isfinite(b).all(dim=1) & torch.isfinite(s)                            # :115-123
    if training and not bool(ok.all()):
        raise FloatingPointError("Predicted boxes or scores contain Inf/NaN. Training has diverged.")
    hw = torch.tensor(image_sizes, dtype=b.dtype, device=dev)[img]                   # (N*T, 2) = (h, w)
    b = torch.stack((b[:, 0].clamp(min=0).minimum(hw[:, 1]), b[:, 1].clamp(min=0).minimum(hw[:, 0]),
                     b[:, 2].clamp(min=0).minimum(hw[:, 1]), b[:, 3].clamp(min=0).minimum(hw[:, 0])), dim=1)   # :124
    ok &= ((b[:, 2] - b[:, 0]) > min_box_size) & ((b[:, 3] - b[:, 1]) > min_box_size)   # :127-129
    sel = torch.nonzero(ok)[:, 0]
    b, s, img = b[sel], s[sel], img[sel]
    keep = nms_fn(b, s, img * len(proposals) + levels.repeat(N)[sel], nms_thresh)    # :131, all images at once
    keep_img = img[keep]
    out = []
    for n, size in enumerate(image_sizes):
        kn = keep[keep_img == n][:post_nms_topk]                                     # :139, already score-sorted
        res = Instances(size)
        res.proposal_boxes = Boxes(b[kn])
        res.objectness_logits = s[kn]
        out.append(res)
    return out


def find_top_rpn_proposals_group(proposals: List[torch.Tensor], pred_objectness_logits: List[torch.Tensor],
                                 image_sizes: List[Tuple[int, int]], num_anchors: List[int], nms_thresh: float,
                                 pre_nms_topk: int, post_nms_topk: int, min_box_size: float, training: bool,
                                 cpgs=None, cpg_strides=None, nms_fn=None, csc_fn=None):
    """proposal_utils.py:146-362 -- the grouped variant: top-k per (level, ANCHOR) (:197-236), NMS groups
    `level * 1000 + anchor` (:335), optionally the CSC re-weighting `scores * (W + 1)` from per-image class peak
    response maps `cpgs[n]` (H, W) with the rois divided by `cpg_strides[n]` (:272-291); returns Instances with
    `proposal_boxes`, `objectness_logits`, `level_ids`.  One NMS call for the whole batch; `csc_fn(cpgs, labels, preds,
    rois) -> W` defaults to the CUDA kernels (ops.csc)."""
    nms_fn = ops.batched_nms if nms_fn is None else nms_fn
    csc_fn = (lambda c, l, p, r: ops.csc(c, l, p, r, 0.1, True, 1.8)) if csc_fn is None else csc_fn
    N = len(image_sizes)
    dev = proposals[0].device
    rows = torch.arange(N, device=dev)[:, None]
    boxes, scores, levels = [], [], []
    for lvl, (p, s) in enumerate(zip(proposals, pred_objectness_logits)):
        A = num_anchors[lvl]
        hw = int(s.shape[1] / A)
        s = s.view(-1, hw, A)
        p = p.view(-1, hw, A, 4)
        k = min(hw, pre_nms_topk)
        for a in range(A):
            s_sorted, order = s[:, :, a].sort(descending=True, dim=1)
            boxes.append(p[:, :, a, :][rows, order[:, :k]])
            scores.append(s_sorted[:, :k])
            levels.append(torch.full((k,), lvl * 1000 + a, dtype=torch.int64, device=dev))
    boxes, scores, levels = torch.cat(boxes, 1), torch.cat(scores, 1), torch.cat(levels)
    T = levels.numel()
    b, s = boxes.reshape(N * T, 4), scores.reshape(N * T)
    img = torch.arange(N, device=dev).repeat_interleave(T)
    lv = levels.repeat(N)
    ok = torch.isfinite(b).all(dim=1) & torch.isfinite(s)
    if training and not bool(ok.all()):
        raise FloatingPointError("Predicted boxes or scores contain Inf/NaN. Training has diverged.")
    hw = torch.tensor(image_sizes, dtype=b.dtype, device=dev)[img]
    b = torch.stack((b[:, 0].clamp(min=0).minimum(hw[:, 1]), b[:, 1].clamp(min=0).minimum(hw[:, 0]),
                     b[:, 2].clamp(min=0).minimum(hw[:, 1]), b[:, 3].clamp(min=0).minimum(hw[:, 0])), dim=1)
    ok &= ((b[:, 2] - b[:, 0]) > min_box_size) & ((b[:, 3] - b[:, 1]) > min_box_size)
    sel = torch.nonzero(ok)[:, 0]
    b, s, img, lv = b[sel], s[sel], img[sel], lv[sel]
    if isinstance(cpgs, (list, tuple)) and len(cpgs) and isinstance(cpgs[0], torch.Tensor):     # :272-291
        s = s.clone()
        for n in range(N):
            m = torch.nonzero(img == n)[:, 0]
            rois = torch.cat((torch.zeros_like(b[m, :1]), b[m] / cpg_strides[n]), dim=1)
            one = torch.ones((1, 1), dtype=cpgs[n].dtype, device=cpgs[n].device)
            W = csc_fn(cpgs[n].unsqueeze(0).unsqueeze(0), one, one, rois)
            s[m] = s[m] * (W.reshape(-1) + 1.0)
    _, dense = torch.unique(lv, return_inverse=True)
    G = int(dense.max()) + 1 if dense.numel() else 1
    keep = nms_fn(b, s, img * G + dense, nms_thresh)
    keep_img = img[keep]
    out = []
    for n, size in enumerate(image_sizes):
        kn = keep[keep_img == n][:post_nms_topk]
        res = Instances(size)
        res.proposal_boxes = Boxes(b[kn])
        res.objectness_logits = s[kn]
        res.level_ids = lv[kn]
        out.append(res)
    return out
