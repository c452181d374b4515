"""Level poolers with the reference's module interface, backed by the sm_100a kernels.

``ROILoopPool`` mirrors wsovod/layers/roi_loop_pool.py:38-58 (constructor, forward(input, rois), the
``rois.dim() == 2 and rois.size(1) == 5`` assertion, repr); ``RoIPool`` / ``ROIAlign`` mirror the
torchvision / detectron2 modules that wsovod/modeling/poolers.py:169-186 instantiates.
Each accepts an optional ``row_scale`` so the caller can fold ``box_features * (objectness + 1)``
(roi_heads.py:733-739) into the kernel's store."""
from torch import nn
from torch.nn.modules.utils import _pair

from .. import ops


def roi_loop_pool(input, roi, output_size, spatial_scale):
    """functional form of the reference's ``roi_loop_pool = _ROILoopPool.apply`` (roi_loop_pool.py:35)"""
    return ops.roi_loop_pool(input, roi, spatial_scale, _pair(output_size), with_argmax=True)[0]


class ROILoopPool(nn.Module):
    def __init__(self, output_size, spatial_scale):
        super().__init__()
        self.output_size = output_size
        self.spatial_scale = spatial_scale

    def forward(self, input, rois, row_scale=None, row_scale_bias=0.0):
        """input: NCHW images; rois: Bx5 boxes, first column is the index into N, then xyxy.
        Returns (3B, C, P, P): roi | frame | context blocks."""
        assert rois.dim() == 2 and rois.size(1) == 5
        need_arg = bool(input.requires_grad)
        return ops.roi_loop_pool(input, rois, self.spatial_scale, _pair(self.output_size), row_scale,
                                 row_scale_bias, with_argmax=need_arg)[0]

    def __repr__(self):
        return (self.__class__.__name__ + "(output_size=" + str(self.output_size) + ", spatial_scale="
                + str(self.spatial_scale) + ")")


class RoIPool(nn.Module):
    """torchvision.ops.RoIPool(output_size, spatial_scale)"""

    def __init__(self, output_size, spatial_scale):
        super().__init__()
        self.output_size = output_size
        self.spatial_scale = spatial_scale

    def forward(self, input, rois, row_scale=None, row_scale_bias=0.0):
        assert rois.dim() == 2 and rois.size(1) == 5
        return ops.roi_pool(input, rois, self.spatial_scale, _pair(self.output_size), row_scale, row_scale_bias)[0]

    def __repr__(self):
        return f"{self.__class__.__name__}(output_size={self.output_size}, spatial_scale={self.spatial_scale})"


class ROIAlign(nn.Module):
    """detectron2.layers.ROIAlign(output_size, spatial_scale, sampling_ratio, aligned=True)"""

    def __init__(self, output_size, spatial_scale, sampling_ratio, aligned=True):
        super().__init__()
        self.output_size = output_size
        self.spatial_scale = spatial_scale
        self.sampling_ratio = sampling_ratio
        self.aligned = aligned

    def forward(self, input, rois, row_scale=None, row_scale_bias=0.0):
        assert rois.dim() == 2 and rois.size(1) == 5
        return ops.roi_align(input, rois, self.spatial_scale, _pair(self.output_size), self.sampling_ratio,
                             self.aligned, row_scale, row_scale_bias)

    def __repr__(self):
        return (f"{self.__class__.__name__}(output_size={self.output_size}, spatial_scale={self.spatial_scale}, "
                f"sampling_ratio={self.sampling_ratio}, aligned={self.aligned})")
