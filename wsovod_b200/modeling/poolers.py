"""``ROIPooler`` with the reference's surface (wsovod/modeling/poolers.py:119-338): same constructor
arguments, same ``forward(x, box_lists, level_ids=None, ...)``, same pooler-type strings, same
multi-level scatter (x3 rows for ROILoopPool, :306-336).  ``forward`` additionally accepts
``objectness_logits`` (list of per-image tensors): when given, the reference's separate
``box_features * (objectness + 1)`` pass (roi_heads.py:733-739) is folded into the pooling kernel."""
import math
from typing import List

import torch
from torch import nn

from ..layers import ROIAlign, ROILoopPool, RoIPool


def convert_boxes_to_pooler_format(box_lists):
    """poolers.py:81-108: list of N Boxes -> (M,5) = (batch index, x0, y0, x1, y1)"""
    rows = []
    for i, b in enumerate(box_lists):
        t = b.tensor
        rows.append(torch.cat((torch.full_like(t[:, :1], i), t), dim=1))
    return rows[0] if len(rows) == 1 else torch.cat(rows, dim=0)


def assign_boxes_to_levels(box_lists, min_level, max_level, canonical_box_size, canonical_level, valid_range=None):
    """poolers.py:24-71 (FPN eqn. 1, or the explicit size ranges of get_valid_range)"""
    sizes = torch.sqrt(torch.cat([b.area() for b in box_lists]))
    if valid_range is not None:
        lv = torch.full_like(sizes, -1)
        for level, (lo, hi) in enumerate(valid_range):
            lv[torch.ge(sizes, lo) & torch.lt(sizes, hi)] = level
        return lv.to(torch.int64)
    lv = torch.floor(canonical_level + torch.log2(sizes / canonical_box_size + 1e-8))
    return torch.clamp(lv, min=min_level, max=max_level).to(torch.int64) - min_level


def get_valid_range():
    return [[0, 60], [60, 160], [160, 2000]]       # poolers.py:111-116


class ROIPooler(nn.Module):
    def __init__(self, output_size, scales, sampling_ratio, pooler_type, canonical_box_size=224, canonical_level=4,
                 use_range=False):
        super().__init__()
        if isinstance(output_size, int):
            output_size = (output_size, output_size)
        assert len(output_size) == 2 and isinstance(output_size[0], int) and isinstance(output_size[1], int)
        self.output_size = output_size
        if pooler_type == "ROIAlign":
            mk = lambda s: ROIAlign(output_size, spatial_scale=s, sampling_ratio=sampling_ratio, aligned=False)  # noqa: E731
        elif pooler_type == "ROIAlignV2":
            mk = lambda s: ROIAlign(output_size, spatial_scale=s, sampling_ratio=sampling_ratio, aligned=True)  # noqa: E731
        elif pooler_type == "ROIPool":
            mk = lambda s: RoIPool(output_size, spatial_scale=s)  # noqa: E731
        elif pooler_type == "ROILoopPool":
            mk = lambda s: ROILoopPool(output_size, spatial_scale=s)  # noqa: E731
        else:
            raise ValueError("Unknown pooler type: {}".format(pooler_type))   # ROIAlignRotated: not on the path
        self.level_poolers = nn.ModuleList(mk(s) for s in scales)
        min_level, max_level = -(math.log2(scales[0])), -(math.log2(scales[-1]))
        assert math.isclose(min_level, int(min_level)) and math.isclose(max_level, int(max_level)), \
            "Featuremap stride is not power of 2!"
        self.min_level, self.max_level = int(min_level), int(max_level)
        assert 0 <= self.min_level <= self.max_level
        assert canonical_box_size > 0
        self.canonical_level, self.canonical_box_size = canonical_level, canonical_box_size
        self.valid_range = get_valid_range() if use_range else None

    def _merged_levels(self, x):
        """the (n_levels * N, C, H, W) tensor the level maps are consecutive chunks of, or None"""
        p0 = self.level_poolers[0]
        if any(getattr(p, "spatial_scale", None) != p0.spatial_scale for p in self.level_poolers):
            return None
        x0 = x[0]
        if not x0.is_cuda or not x0.is_contiguous() or (x0.requires_grad and torch.is_grad_enabled()):
            return None
        step = x0.numel() * x0.element_size()
        for level, t in enumerate(x):
            if t.shape != x0.shape or t.dtype != x0.dtype or not t.is_contiguous() or t.device != x0.device:
                return None
            if t.data_ptr() != x0.data_ptr() + level * step or t.untyped_storage().data_ptr() != x0.untyped_storage().data_ptr():
                return None
        return torch.as_strided(x0, (len(x) * x0.size(0),) + tuple(x0.shape[1:]), x0.stride(), x0.storage_offset())

    def forward(self, x: List[torch.Tensor], box_lists, level_ids=None, oh_labels_list=None, superpixels=None,
                objectness_logits=None):
        n_levels = len(self.level_poolers)
        assert isinstance(x, list) and isinstance(box_lists, list), "Arguments to pooler must be lists"
        assert len(x) == n_levels, \
            "unequal value, num_level_assignments={}, but x is list of {} Tensors".format(n_levels, len(x))
        assert len(box_lists) == x[0].size(0), \
            "unequal value, x[0] batch dim 0 is {}, but box_list has length {}".format(x[0].size(0), len(box_lists))
        if superpixels is not None:
            raise NotImplementedError("superpixel pooling is not part of the shipped reference poolers")
        if len(box_lists) == 0:
            return torch.zeros((0, x[0].shape[1]) + self.output_size, device=x[0].device, dtype=x[0].dtype)
        rois = convert_boxes_to_pooler_format(box_lists)
        scale = None if objectness_logits is None else torch.cat(list(objectness_logits), dim=0)
        three = isinstance(self.level_poolers[0], ROILoopPool)
        if n_levels == 1:
            return self.level_poolers[0](x[0], rois, scale, 1.0) if scale is not None else self.level_poolers[0](x[0], rois)
        lv = assign_boxes_to_levels(box_lists, self.min_level, self.max_level, self.canonical_box_size,
                                    self.canonical_level, self.valid_range)
        if level_ids is not None:
            lv = torch.cat(list(level_ids)).to(torch.int64)
        merged = self._merged_levels(x)
        if merged is not None:
            # MRRP (roi_heads.py:723-730): the levels are the chunks of ONE batched map at one scale.  The image index of a
            # proposal becomes level * N + image and the whole thing is a single kernel launch whose rows come out in
            # the proposals' own order -- no per-level nonzero (a host sync each), gather and index_put.  A proposal whose
            # level id matches no level keeps the reference's zero rows: its box is moved off the map, where every bin
            # of every stream is empty.
            N = x[0].size(0)
            valid = (lv >= 0) & (lv < n_levels)
            far = torch.full_like(rois[:, 1:], 1.0e6)
            r = torch.cat((rois[:, :1] + (lv.clamp(0, n_levels - 1) * N).to(rois.dtype).unsqueeze(1),
                           torch.where(valid.unsqueeze(1), rois[:, 1:], far)), dim=1)
            pooler = self.level_poolers[0]
            return pooler(merged, r, scale, 1.0) if scale is not None else pooler(merged, r)
        M, C, P = rois.size(0), x[0].shape[1], self.output_size[0]
        out = torch.zeros(((3 * M) if three else M, C, P, P), dtype=x[0].dtype, device=x[0].device)
        for level, pooler in enumerate(self.level_poolers):
            inds = torch.nonzero(lv == level, as_tuple=True)[0]
            r = rois[inds]
            res = pooler(x[level], r, scale[inds], 1.0) if scale is not None else pooler(x[level], r)
            if three:
                inds = torch.cat([inds, inds + M, inds + 2 * M], dim=0)
            out.index_put_((inds,), res)
        return out
