"""``CSC`` / ``CSCConstraint`` with the reference's signatures (wsovod/layers/csc.py:9-143).

``csc`` runs ``wsovod_b200_csc_fwd`` (csrc/csc.cu): the reference's ``_C.csc_forward`` (csc_cuda.cu:183-531) as three
stream-ordered launches instead of a host loop with a device synchronisation per (image, class).  Like ``_CSC.apply``
it returns ``(W, PL, NL)`` with ``PL = labels.clone().detach()`` and ``NL = zeros_like(labels)`` (csc.py:25-26,43) and
is not differentiable (:46-49).  ``tau``, ``mass_threshold`` and ``density_threshold`` are accepted and, as upstream,
do not influence the result (their uses are commented out, csc_cuda.cu:424-426,322).  ``csc_constraint`` is plain tensor
arithmetic in the reference (csc.py:102-125) and is kept as such."""
import torch
from torch import nn
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from .. import ops


def csc(cpgs, labels, preds, rois, tau=0.7, debug_info=False, fg_threshold=0.1, mass_threshold=0.2,
        density_threshold=0.0, area_sqrt=True, context_scale=1.8):
    PL = labels.clone().detach()
    NL = torch.zeros(labels.size(), dtype=labels.dtype, device=labels.device)
    with torch.no_grad():
        W = ops.csc(cpgs, labels, preds, rois, fg_threshold, area_sqrt, context_scale)
    return W, PL, NL


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
