"""Stand-ins for detectron2 / fvcore / clip so the reference's hot-path Python imports VERBATIM.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  detectron2 is an un-pinned, un-vendored
dependency of the reference (requirements.txt:3) and cannot be installed here (no network), so the
few primitives the hot path calls are restated below from their call sites and detectron2's
documented behaviour.  ``install(reference_root)`` registers the fake
modules plus namespace packages for ``wsovod``/``wsovod.modeling``/``wsovod.modeling.roi_heads`` so
that the detectron2-heavy ``__init__``s of the reference do not run, after which e.g.
``import wsovod.modeling.roi_heads.fast_rcnn_open_vocabulary`` executes the reference file itself.

Used by oracle/make_golden.py (in the build container, where /root/reference exists) and -- through the
byte-for-byte copies oracle/build_ref.py ships under the git-ignored oracle/_ref/py/ -- by the drop-in test and
`bench.py --impl reference` on the GPU box.
"""
import importlib.util
import math
import sys
import types
from dataclasses import dataclass
from typing import Optional

import torch
import torchvision
from torch import nn
from torch.nn import functional as F


# ---------------------------------------------------------------- detectron2.config
def configurable(init_func=None, *, from_config=None):
    import functools

    def _called_with_cfg(*args, **kwargs):
        return (len(args) and isinstance(args[0], CfgNode)) or isinstance(kwargs.get("cfg"), CfgNode)

    assert init_func is not None and init_func.__name__ == "__init__"

    @functools.wraps(init_func)
    def wrapped(self, *args, **kwargs):
        if _called_with_cfg(*args, **kwargs):
            explicit = type(self).from_config(*args, **kwargs)
            init_func(self, **explicit)
        else:
            init_func(self, *args, **kwargs)

    return wrapped


class CfgNode(dict):
    """attribute-style nested dict (just enough of yacs for from_config)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


# ---------------------------------------------------------------- detectron2.layers
@dataclass
class ShapeSpec:
    channels: Optional[int] = None
    height: Optional[int] = None
    width: Optional[int] = None
    stride: Optional[int] = None


def cat(tensors, dim=0):
    assert isinstance(tensors, (list, tuple))
    if len(tensors) == 1:
        return tensors[0]
    return torch.cat(tensors, dim)


def nonzero_tuple(x):
    if x.dim() == 0:
        return x.unsqueeze(0).nonzero().unbind(1)
    return x.nonzero().unbind(1)


def batched_nms(boxes, scores, idxs, iou_threshold):
    assert boxes.shape[-1] == 4
    return torchvision.ops.boxes.batched_nms(boxes.float(), scores, idxs, iou_threshold)


def cross_entropy(input, target, *, reduction="mean", **kwargs):
    if target.numel() == 0 and reduction == "mean":
        return input.sum() * 0.0
    return F.cross_entropy(input, target, reduction=reduction, **kwargs)


def _unavailable(*a, **k):
    raise NotImplementedError("not on the region-scoring path")


# ---------------------------------------------------------------- detectron2.structures
class Boxes:
    def __init__(self, tensor):
        if not isinstance(tensor, torch.Tensor):
            tensor = torch.as_tensor(tensor, dtype=torch.float32)
        else:
            tensor = tensor.to(torch.float32)
        if tensor.numel() == 0:
            tensor = tensor.reshape((-1, 4)).to(dtype=torch.float32)
        assert tensor.dim() == 2 and tensor.size(-1) == 4, tensor.size()
        self.tensor = tensor

    def clone(self):
        return Boxes(self.tensor.clone())

    def to(self, device):
        return Boxes(self.tensor.to(device=device))

    def area(self):
        box = self.tensor
        return (box[:, 2] - box[:, 0]) * (box[:, 3] - box[:, 1])

    def clip(self, box_size):
        assert torch.isfinite(self.tensor).all(), "Box tensor contains infinite or NaN!"
        h, w = box_size
        x1 = self.tensor[:, 0].clamp(min=0, max=w)
        y1 = self.tensor[:, 1].clamp(min=0, max=h)
        x2 = self.tensor[:, 2].clamp(min=0, max=w)
        y2 = self.tensor[:, 3].clamp(min=0, max=h)
        self.tensor = torch.stack((x1, y1, x2, y2), dim=-1)

    def nonempty(self, threshold=0.0):
        box = self.tensor
        return ((box[:, 2] - box[:, 0]) > threshold) & ((box[:, 3] - box[:, 1]) > threshold)

    def __getitem__(self, item):
        if isinstance(item, int):
            return Boxes(self.tensor[item].view(1, -1))
        b = self.tensor[item]
        assert b.dim() == 2
        return Boxes(b)

    def __len__(self):
        return self.tensor.shape[0]

    @classmethod
    def cat(cls, boxes_list):
        if len(boxes_list) == 0:
            return cls(torch.empty(0))
        return cls(torch.cat([b.tensor for b in boxes_list], dim=0))

    @property
    def device(self):
        return self.tensor.device


class Instances:
    def __init__(self, image_size, **kwargs):
        self._image_size = image_size
        self._fields = {}
        for k, v in kwargs.items():
            self.set(k, v)

    @property
    def image_size(self):
        return self._image_size

    def __setattr__(self, name, val):
        if name.startswith("_"):
            super().__setattr__(name, val)
        else:
            self.set(name, val)

    def __getattr__(self, name):
        if name == "_fields" or name not in self._fields:
            raise AttributeError("Cannot find field '{}' in the given Instances!".format(name))
        return self._fields[name]

    def set(self, name, value):
        self._fields[name] = value

    def has(self, name):
        return name in self._fields

    def get(self, name):
        return self._fields[name]

    def get_fields(self):
        return self._fields

    def __getitem__(self, item):
        ret = Instances(self._image_size)
        for k, v in self._fields.items():
            ret.set(k, v[item])
        return ret

    def __len__(self):
        for v in self._fields.values():
            return v.__len__()
        raise NotImplementedError("Empty Instances does not support __len__!")


def pairwise_iou(boxes1, boxes2):
    area1, area2 = boxes1.area(), boxes2.area()
    b1, b2 = boxes1.tensor, boxes2.tensor
    wh = torch.min(b1[:, None, 2:], b2[:, 2:]) - torch.max(b1[:, None, :2], b2[:, :2])
    wh.clamp_(min=0)
    inter = wh.prod(dim=2)
    return torch.where(
        inter > 0,
        inter / (area1[:, None] + area2 - inter),
        torch.zeros(1, dtype=inter.dtype, device=inter.device),
    )


# ---------------------------------------------------------------- detectron2.modeling.*
class Matcher:
    def __init__(self, thresholds, labels, allow_low_quality_matches=False):
        thresholds = thresholds[:]
        thresholds.insert(0, -float("inf"))
        thresholds.append(float("inf"))
        self.thresholds, self.labels = thresholds, labels
        assert not allow_low_quality_matches

    def __call__(self, m):
        if m.numel() == 0:
            return (m.new_full((m.size(1),), 0, dtype=torch.int64),
                    m.new_full((m.size(1),), self.labels[0], dtype=torch.int8))
        vals, matches = m.max(dim=0)
        match_labels = matches.new_full(matches.size(), 1, dtype=torch.int8)
        for l, low, high in zip(self.labels, self.thresholds[:-1], self.thresholds[1:]):
            match_labels[(vals >= low) & (vals < high)] = l
        return matches, match_labels


def subsample_labels(labels, num_samples, positive_fraction, bg_label):
    positive = nonzero_tuple((labels != -1) & (labels != bg_label))[0]
    negative = nonzero_tuple(labels == bg_label)[0]
    num_pos = min(positive.numel(), int(num_samples * positive_fraction))
    num_neg = min(negative.numel(), num_samples - num_pos)
    perm1 = torch.randperm(positive.numel(), device=positive.device)[:num_pos]
    perm2 = torch.randperm(negative.numel(), device=negative.device)[:num_neg]
    return positive[perm1], negative[perm2]


class Box2BoxTransform:
    def __init__(self, weights, scale_clamp=math.log(1000.0 / 16)):
        self.weights, self.scale_clamp = weights, scale_clamp

    def get_deltas(self, src_boxes, target_boxes):
        """detectron2 Box2BoxTransform.get_deltas (box_regression.py, restated): (dx, dy, dw, dh) that
        apply_deltas would need to move src onto target"""
        src_widths = src_boxes[:, 2] - src_boxes[:, 0]
        src_heights = src_boxes[:, 3] - src_boxes[:, 1]
        src_ctr_x = src_boxes[:, 0] + 0.5 * src_widths
        src_ctr_y = src_boxes[:, 1] + 0.5 * src_heights
        target_widths = target_boxes[:, 2] - target_boxes[:, 0]
        target_heights = target_boxes[:, 3] - target_boxes[:, 1]
        target_ctr_x = target_boxes[:, 0] + 0.5 * target_widths
        target_ctr_y = target_boxes[:, 1] + 0.5 * target_heights
        wx, wy, ww, wh = self.weights
        dx = wx * (target_ctr_x - src_ctr_x) / src_widths
        dy = wy * (target_ctr_y - src_ctr_y) / src_heights
        dw = ww * torch.log(target_widths / src_widths)
        dh = wh * torch.log(target_heights / src_heights)
        deltas = torch.stack((dx, dy, dw, dh), dim=1)
        assert (src_widths > 0).all().item(), "Input boxes to Box2BoxTransform are not valid!"
        return deltas

    def apply_deltas(self, deltas, boxes):
        deltas = deltas.float()
        boxes = boxes.to(deltas.dtype)
        widths = boxes[:, 2] - boxes[:, 0]
        heights = boxes[:, 3] - boxes[:, 1]
        ctr_x = boxes[:, 0] + 0.5 * widths
        ctr_y = boxes[:, 1] + 0.5 * heights
        wx, wy, ww, wh = self.weights
        dx, dy = deltas[:, 0::4] / wx, deltas[:, 1::4] / wy
        dw, dh = deltas[:, 2::4] / ww, deltas[:, 3::4] / wh
        dw = torch.clamp(dw, max=self.scale_clamp)
        dh = torch.clamp(dh, max=self.scale_clamp)
        pcx = dx * widths[:, None] + ctr_x[:, None]
        pcy = dy * heights[:, None] + ctr_y[:, None]
        pw = torch.exp(dw) * widths[:, None]
        ph = torch.exp(dh) * heights[:, None]
        out = torch.stack((pcx - 0.5 * pw, pcy - 0.5 * ph, pcx + 0.5 * pw, pcy + 0.5 * ph), dim=-1)
        return out.reshape(deltas.shape)


class _Registry:
    def __init__(self):
        self._m = {}

    def register(self, obj=None):
        def deco(o):
            self._m[o.__name__] = o
            return o
        return deco if obj is None else deco(obj)

    def get(self, name):
        return self._m[name]


class _Sink:
    iter = 0

    def put_scalar(self, *a, **k):
        pass

    def put_image(self, *a, **k):
        pass


def get_event_storage():
    return _Sink()


def smooth_l1_loss(input, target, beta, reduction="none"):
    if beta < 1e-5:
        loss = torch.abs(input - target)
    else:
        n = torch.abs(input - target)
        loss = torch.where(n < beta, 0.5 * n ** 2 / beta, n - 0.5 * beta)
    if reduction == "mean":
        return loss.mean() if loss.numel() > 0 else 0.0 * loss.sum()
    if reduction == "sum":
        return loss.sum()
    return loss


class ROIAlign(nn.Module):
    """detectron2.layers.ROIAlign as poolers.py:169-182 constructs it: torchvision's compiled roi_align"""

    def __init__(self, output_size, spatial_scale, sampling_ratio, aligned=True):
        super().__init__()
        self.output_size = (output_size, output_size) if isinstance(output_size, int) else tuple(output_size)
        self.spatial_scale, self.sampling_ratio, self.aligned = spatial_scale, sampling_ratio, aligned

    def forward(self, input, rois):
        assert rois.dim() == 2 and rois.size(1) == 5
        return torchvision.ops.roi_align(input, rois.to(dtype=input.dtype), self.output_size, self.spatial_scale,
                                         self.sampling_ratio, self.aligned)


def default_reference_root():
    """/root/reference in the build container; the shipped copies (oracle/_ref/py) on the GPU box; None if neither"""
    import os
    if os.path.isdir("/root/reference/wsovod"):
        return "/root/reference"
    shipped = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "py")
    return shipped if os.path.isdir(os.path.join(shipped, "wsovod")) else None


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install(reference_root=None):
    """Register the fake modules; afterwards the reference hot-path files import verbatim."""
    if "detectron2" in sys.modules and getattr(sys.modules["detectron2"], "_wsovod_shim", False):
        return
    if reference_root is None:
        reference_root = default_reference_root()
    if reference_root is None:
        raise RuntimeError("no reference tree: neither /root/reference nor oracle/_ref/py (run oracle/build_ref.py)")
    d2 = _mod("detectron2", _wsovod_shim=True)
    d2.__path__ = []
    _mod("detectron2.config", configurable=configurable, CfgNode=CfgNode)
    _mod("detectron2.layers", ShapeSpec=ShapeSpec, cat=cat, nonzero_tuple=nonzero_tuple,
         batched_nms=batched_nms, cross_entropy=cross_entropy, Linear=nn.Linear,
         ciou_loss=_unavailable, diou_loss=_unavailable, ROIAlign=ROIAlign,
         ROIAlignRotated=_unavailable)
    _mod("detectron2.structures", Boxes=Boxes, Instances=Instances, pairwise_iou=pairwise_iou,
         ImageList=object, PolygonMasks=object)
    dm = _mod("detectron2.modeling")
    dm.__path__ = []
    _mod("detectron2.modeling.box_regression", Box2BoxTransform=Box2BoxTransform)
    _mod("detectron2.modeling.matcher", Matcher=Matcher)
    _mod("detectron2.modeling.sampling", subsample_labels=subsample_labels)
    pg = _mod("detectron2.modeling.proposal_generator")
    pg.__path__ = []
    _mod("detectron2.modeling.proposal_generator.proposal_utils",
         add_ground_truth_to_proposals=_unavailable)
    rh = _mod("detectron2.modeling.roi_heads", ROI_HEADS_REGISTRY=_Registry())
    rh.__path__ = []
    _mod("detectron2.modeling.roi_heads.box_head", build_box_head=_unavailable)
    _mod("detectron2.data", MetadataCatalog=object)
    du = _mod("detectron2.utils")
    du.__path__ = []
    _mod("detectron2.utils.events", get_event_storage=get_event_storage)
    _mod("detectron2.utils.comm", get_world_size=lambda: 1, get_rank=lambda: 0,
         is_main_process=lambda: True)
    _mod("detectron2.utils.visualizer", Visualizer=object)
    fv = _mod("fvcore")
    fv.__path__ = []
    _mod("fvcore.nn", smooth_l1_loss=smooth_l1_loss, giou_loss=_unavailable)
    _mod("clip")
    try:
        import cv2  # noqa: F401  (roi_heads.py imports it at module level)
    except Exception:
        _mod("cv2")
    # namespace packages: make `wsovod.modeling.*` importable file by file, skipping the __init__s
    for pkg, sub in (("wsovod", "wsovod"), ("wsovod.modeling", "wsovod/modeling"),
                     ("wsovod.modeling.roi_heads", "wsovod/modeling/roi_heads"),
                     ("wsovod.modeling.class_heads", "wsovod/modeling/class_heads")):
        m = types.ModuleType(pkg)
        m.__path__ = [f"{reference_root}/{sub}"]
        sys.modules[pkg] = m
    # `from wsovod.modeling.class_heads import OpenVocabularyClassifier` (fast_rcnn_open_vocabulary.py:19)
    spec = importlib.util.spec_from_file_location(
        "wsovod.modeling.class_heads.open_vocabulary_classifier",
        f"{reference_root}/wsovod/modeling/class_heads/open_vocabulary_classifier.py")
    ovc = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = ovc
    spec.loader.exec_module(ovc)
    sys.modules["wsovod.modeling.class_heads"].OpenVocabularyClassifier = ovc.OpenVocabularyClassifier
    # wsovod.layers (ROILoopPool) is needed by poolers.py only: the reference's own wrapper over its own compiled
    # extension when oracle/_ref/wsovod_ref_C.so exists (`from wsovod import _C`, roi_loop_pool.py:6), else a
    # placeholder class (the CPU container has no GPU to run it on anyway)
    layers = _mod("wsovod.layers", ROILoopPool=type("ROILoopPool", (nn.Module,), {}))
    layers.__path__ = [f"{reference_root}/wsovod/layers"]
    try:
        import os
        from . import ref
        ext = ref.cuda() if os.path.exists(f"{reference_root}/wsovod/layers/roi_loop_pool.py") else None
        if ext is not None:
            sys.modules["wsovod"]._C = ext
            sys.modules["wsovod._C"] = ext
            spec = importlib.util.spec_from_file_location("wsovod.layers.roi_loop_pool",
                                                          f"{reference_root}/wsovod/layers/roi_loop_pool.py")
            rlp = importlib.util.module_from_spec(spec)
            sys.modules[spec.name] = rlp
            spec.loader.exec_module(rlp)
            layers.ROILoopPool = rlp.ROILoopPool
    except Exception:  # noqa: BLE001  (no libtorch_cuda / no extension: keep the placeholder)
        pass
