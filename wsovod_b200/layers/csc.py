"""``CSC`` / ``CSCConstraint`` with the reference's signatures (wsovod/layers/csc.py:9-143).

``csc_constraint`` is plain tensor arithmetic in the reference (csc.py:102-125) and is kept as such.
``csc_forward`` (wsovod/layers/csc/csc_cuda.cu) is dead code in every shipped config -- its only call
site needs ``cpgs`` that are never produced -- and is outside the region-scoring path
this library accelerates, so ``CSC.forward`` raises instead of silently computing something else."""
import torch
from torch import nn
from torch.autograd import Function
from torch.autograd.function import once_differentiable


def csc(cpgs, labels, preds, rois, tau=0.7, debug_info=False, fg_threshold=0.1, mass_threshold=0.2,
        density_threshold=0.0, area_sqrt=True, context_scale=1.8):
    raise NotImplementedError(
        "wsovod_b200: csc_forward is not on the region-scoring path (unused by every shipped WSOVOD config)")


class CSC(nn.Module):
    def __init__(self, tau=0.7, debug_info=False, fg_threshold=0.1, mass_threshold=0.2, density_threshold=0.0,
                 area_sqrt=True, context_scale=1.8):
        super().__init__()
        self.tau, self.debug_info, self.fg_threshold = tau, debug_info, fg_threshold
        self.mass_threshold, self.density_threshold = mass_threshold, density_threshold
        self.area_sqrt, self.context_scale = area_sqrt, context_scale

    def forward(self, cpgs, labels, preds, rois):
        return csc(cpgs, labels, preds, rois, self.tau, self.debug_info, self.fg_threshold, self.mass_threshold,
                   self.density_threshold, self.area_sqrt, self.context_scale)


class _CSCConstraint(Function):
    @staticmethod
    def forward(ctx, X, W, polar):
        W_ = torch.clamp(W, min=0.0) if polar else torch.clamp(W, max=0.0) * (-1.0)
        ctx.save_for_backward(W_)
        return X * W_

    @staticmethod
    @once_differentiable
    def backward(ctx, dY):
        (W_,) = ctx.saved_tensors
        return dY * W_, None, None


csc_constraint = _CSCConstraint.apply


class CSCConstraint(nn.Module):
    def __init__(self, polar=True):
        super().__init__()
        self.polar = polar

    def forward(self, X, W):
        return csc_constraint(X, W, self.polar)
