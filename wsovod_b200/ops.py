"""Functional ops: torch CUDA tensors in, torch CUDA tensors out, computed by libwsovod_b200.so.

Each op is registered as a torch custom op (``torch.ops.wsovod_b200.*``) with a fake (meta)
implementation and, where the reference op is differentiable, an autograd formula that calls the
matching ``*_bwd`` kernel.  CPU tensors are rejected like the reference's own ops do
(wsovod/layers/ROILoopPool/ROILoopPool.h:62 "Not compiled with CPU support").
"""
import os
from typing import List, Optional, Tuple

import torch

from . import _lib
from ._lib import c_p

IOU_TV_CPU = 0
IOU_TV_CUDA = 1
ALIGN_FP32 = 0
ALIGN_TF32 = 1


def _ptr(t):
    return c_p(0) if t is None else c_p(t.data_ptr())


def _stream(t):
    return c_p(torch.cuda.current_stream(t.device).cuda_stream)


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("wsovod_b200: Not compiled with CPU support (tensor is on %s)" % t.device)
    dev = [t.device for t in ts if t is not None]
    if any(d != dev[0] for d in dev):
        raise RuntimeError("wsovod_b200: all tensors must be on the same GPU")


def _f32c(t):
    if t.dtype != torch.float32:
        raise RuntimeError("wsovod_b200: only float32 is implemented (got %s)" % t.dtype)
    return t.contiguous()


def _workspace(nbytes, device):
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


def _pair(v):
    return (v, v) if isinstance(v, int) else (int(v[0]), int(v[1]))


# Eager fast path.  Every op below is registered as a torch custom op (torch.ops.wsovod_b200.*: schema, fake kernel,
# autograd formula -- what torch.compile / export see).  The dispatcher's Python layers cost 50-100 us per call, which
# is the price of a dozen of these kernels; in plain eager mode the public functions therefore call the op's body
# directly and, where autograd is needed, through a torch.autograd.Function built from the SAME setup / backward
# functions the custom op registers.  WSOVOD_B200_EAGER_FAST=0 routes everything through torch.ops again.
_FAST = os.environ.get("WSOVOD_B200_EAGER_FAST", "1") != "0"
_FAST_FN = {}


def _body(op):
    return op._init_fn


def _fast_function(op, setup, backward):
    body = _body(op)

    class Fn(torch.autograd.Function):
        @staticmethod
        def forward(ctx, *inputs):
            out = body(*inputs)
            setup(ctx, inputs, out)
            return tuple(out) if isinstance(out, (list, tuple)) else out

        @staticmethod
        def backward(ctx, *grads):
            return backward(ctx, *grads)

    Fn.__name__ = "Fast_" + op._opoverload.name().replace("::", "_") if hasattr(op, "_opoverload") else "Fast"
    return Fn


def _call(op, *args):
    """run custom op `op` on `args`: torch.ops dispatch under tracing / compile (or WSOVOD_B200_EAGER_FAST=0), else its
    body directly, wrapped in the op's autograd formula when a tensor argument requires grad"""
    if not _FAST or torch.compiler.is_compiling():
        return op(*args)
    fn = _FAST_FN.get(op)
    if fn is not None and torch.is_grad_enabled() and any(isinstance(a, torch.Tensor) and a.requires_grad for a in args):
        return fn.apply(*args)
    with torch.no_grad():
        return _body(op)(*args)


# ------------------------------------------------------------------------------------------------
# (1) pooling
# ------------------------------------------------------------------------------------------------
@torch.library.custom_op("wsovod_b200::roi_pool", mutates_args=())
def _roi_pool(input: torch.Tensor, rois: torch.Tensor, spatial_scale: float, pooled_h: int,
              pooled_w: int, row_scale: Optional[torch.Tensor], row_scale_bias: float,
              with_argmax: bool) -> Tuple[torch.Tensor, torch.Tensor]:
    _need_cuda(input, rois, row_scale)
    input, rois = _f32c(input), _f32c(rois)
    if rois.dim() != 2 or rois.size(1) != 5 or input.dim() != 4:
        raise RuntimeError("wsovod_b200::roi_pool expects input NCHW and rois (R,5)")
    N, C, H, W = input.shape
    R = rois.size(0)
    rs = None if row_scale is None else _f32c(row_scale)
    with torch.cuda.device(input.device):
        out = torch.empty((R, C, pooled_h, pooled_w), dtype=torch.float32, device=input.device)
        arg = torch.empty((R, C, pooled_h, pooled_w) if with_argmax else (0,), dtype=torch.int32,
                          device=input.device)
        L = _lib.lib()
        ws = _workspace(L.wsovod_b200_roi_pool_workspace(N, R, pooled_h, pooled_w), input.device)
        rc = L.wsovod_b200_roi_pool_fwd(_ptr(input), N, C, H, W, _ptr(rois), R, spatial_scale, pooled_h,
                                        pooled_w, _ptr(rs), row_scale_bias, _ptr(out),
                                        _ptr(arg if with_argmax else None), _ptr(ws), ws.numel(),
                                        _stream(input))
    _lib.check(rc, "roi_pool_fwd")
    return out, arg


@_roi_pool.register_fake
def _(input, rois, spatial_scale, pooled_h, pooled_w, row_scale, row_scale_bias, with_argmax):
    R, C = rois.size(0), input.size(1)
    out = input.new_empty((R, C, pooled_h, pooled_w))
    arg = input.new_empty((R, C, pooled_h, pooled_w) if with_argmax else (0,), dtype=torch.int32)
    return out, arg


@torch.library.custom_op("wsovod_b200::roi_pool_backward", mutates_args=())
def _roi_pool_backward(grad: torch.Tensor, rois: torch.Tensor, argmax: torch.Tensor, N: int, C: int,
                       H: int, W: int, three_way: bool) -> torch.Tensor:
    _need_cuda(grad, rois, argmax)
    if three_way and grad.dtype in _LOOP_DTYPES:
        if rois.dtype != grad.dtype:
            raise RuntimeError("expected scalar type %s but found %s" % (grad.dtype, rois.dtype))
        grad, rois, argmax = grad.contiguous(), rois.contiguous(), argmax.contiguous()
        ph, pw = grad.shape[-2:]
        with torch.cuda.device(grad.device):
            gi = torch.zeros((N, C, H, W), dtype=grad.dtype, device=grad.device)
            rc = _lib.lib().wsovod_b200_roi_loop_pool_dtype_bwd(_LOOP_DTYPES[grad.dtype], _ptr(grad), _ptr(rois), _ptr(argmax),
                                                                rois.size(0), N, C, H, W, ph, pw, _ptr(gi), _stream(grad))
        _lib.check(rc, "roi_loop_pool_dtype_bwd")
        return gi
    grad, rois = _f32c(grad), _f32c(rois)
    argmax = argmax.contiguous()
    R = rois.size(0)
    ph, pw = grad.shape[-2:]
    with torch.cuda.device(grad.device):
        gi = torch.zeros((N, C, H, W), dtype=torch.float32, device=grad.device)
        L = _lib.lib()
        fn = L.wsovod_b200_roi_loop_pool_bwd if three_way else L.wsovod_b200_roi_pool_bwd
        rc = fn(_ptr(grad), _ptr(rois), _ptr(argmax), R, N, C, H, W, ph, pw, _ptr(gi), _stream(grad))
    _lib.check(rc, "roi_pool_bwd")
    return gi


@_roi_pool_backward.register_fake
def _(grad, rois, argmax, N, C, H, W, three_way):
    return grad.new_empty((N, C, H, W))


def _roi_pool_setup(ctx, inputs, output):
    input, rois, spatial_scale, ph, pw, row_scale, bias, with_argmax = inputs
    ctx.shape = tuple(input.shape)
    ctx.has_scale = row_scale is not None
    ctx.bias = bias
    ctx.with_argmax = with_argmax
    ctx.save_for_backward(rois, output[1], row_scale)
    ctx.mark_non_differentiable(output[1])


def _roi_pool_bwd(ctx, grad_out, _grad_arg):
    rois, argmax, row_scale = ctx.saved_tensors
    if not ctx.with_argmax:
        raise RuntimeError("wsovod_b200::roi_pool: backward needs with_argmax=True")
    g = grad_out
    if ctx.has_scale:
        g = g * (row_scale + ctx.bias).view(-1, 1, 1, 1)
    N, C, H, W = ctx.shape
    gi = _call(_roi_pool_backward, g, rois, argmax, N, C, H, W, False)
    return gi, None, None, None, None, None, None, None


_roi_pool.register_autograd(_roi_pool_bwd, setup_context=_roi_pool_setup)
_FAST_FN[_roi_pool] = _fast_function(_roi_pool, _roi_pool_setup, _roi_pool_bwd)


def roi_pool(input, rois, spatial_scale, output_size, row_scale=None, row_scale_bias=0.0,
             with_argmax=None):
    """torchvision.ops.roi_pool semantics; returns (output, argmax) -- argmax is an empty tensor when
    skipped.  ``with_argmax`` defaults to ``input.requires_grad`` (the frozen-backbone configs of the
    reference never read it)."""
    ph, pw = _pair(output_size)
    if with_argmax is None:
        with_argmax = bool(input.requires_grad and torch.is_grad_enabled())
    return _call(_roi_pool, input, rois, float(spatial_scale), ph, pw, row_scale,
                                          float(row_scale_bias), bool(with_argmax))


_LOOP_DTYPES = {torch.float16: _lib.F16, torch.float64: _lib.F64}


def _roi_loop_pool_dtype(input, rois, spatial_scale, pooled_h, pooled_w, row_scale, row_scale_bias, with_argmax):
    """half / double ROILoopPool (ROILoopPool_cuda.cu:294: the reference dispatches float, double and half): rois carry
    the input's dtype as `rois.data_ptr<scalar_t>()` demands; the row scale, which the reference applies as a separate
    multiply in that dtype (roi_heads.py:733-739), stays a separate multiply."""
    if rois.dtype != input.dtype:
        raise RuntimeError("expected scalar type %s but found %s" % (input.dtype, rois.dtype))
    input, rois = input.contiguous(), rois.contiguous()
    N, C, H, W = input.shape
    R = rois.size(0)
    with torch.cuda.device(input.device):
        out = torch.empty((3 * R, C, pooled_h, pooled_w), dtype=input.dtype, device=input.device)
        arg = torch.empty((3 * R, C, pooled_h, pooled_w) if with_argmax else (0,), dtype=torch.int32, device=input.device)
        L = _lib.lib()
        ws = _workspace(L.wsovod_b200_roi_loop_pool_dtype_workspace(N, R, pooled_h, pooled_w), input.device)
        rc = L.wsovod_b200_roi_loop_pool_dtype_fwd(_LOOP_DTYPES[input.dtype], _ptr(input), N, C, H, W, _ptr(rois), R,
                                                   spatial_scale, pooled_h, pooled_w, _ptr(out),
                                                   _ptr(arg if with_argmax else None), _ptr(ws), ws.numel(), _stream(input))
    _lib.check(rc, "roi_loop_pool_dtype_fwd")
    if row_scale is not None:
        s = row_scale.to(input.dtype) + row_scale_bias
        out = out * torch.cat([s, s, s]).view(-1, 1, 1, 1)
    return out, arg


@torch.library.custom_op("wsovod_b200::roi_loop_pool", mutates_args=())
def _roi_loop_pool(input: torch.Tensor, rois: torch.Tensor, spatial_scale: float, pooled_h: int,
                   pooled_w: int, row_scale: Optional[torch.Tensor], row_scale_bias: float,
                   with_argmax: bool) -> Tuple[torch.Tensor, torch.Tensor]:
    _need_cuda(input, rois, row_scale)
    if rois.dim() != 2 or rois.size(1) != 5 or input.dim() != 4:
        raise RuntimeError("wsovod_b200::roi_loop_pool expects input NCHW and rois (R,5)")
    if input.dtype in _LOOP_DTYPES:
        return _roi_loop_pool_dtype(input, rois, spatial_scale, pooled_h, pooled_w, row_scale, row_scale_bias, with_argmax)
    input, rois = _f32c(input), _f32c(rois)
    N, C, H, W = input.shape
    R = rois.size(0)
    rs = None if row_scale is None else _f32c(row_scale)
    with torch.cuda.device(input.device):
        out = torch.empty((3 * R, C, pooled_h, pooled_w), dtype=torch.float32, device=input.device)
        arg = torch.empty((3 * R, C, pooled_h, pooled_w) if with_argmax else (0,), dtype=torch.int32,
                          device=input.device)
        L = _lib.lib()
        ws = _workspace(L.wsovod_b200_roi_loop_pool_workspace(N, R, pooled_h, pooled_w), input.device)
        rc = L.wsovod_b200_roi_loop_pool_fwd(_ptr(input), N, C, H, W, _ptr(rois), R, spatial_scale,
                                             pooled_h, pooled_w, _ptr(rs), row_scale_bias, _ptr(out),
                                             _ptr(arg if with_argmax else None), _ptr(ws), ws.numel(),
                                             _stream(input))
    _lib.check(rc, "roi_loop_pool_fwd")
    return out, arg


@_roi_loop_pool.register_fake
def _(input, rois, spatial_scale, pooled_h, pooled_w, row_scale, row_scale_bias, with_argmax):
    R, C = rois.size(0), input.size(1)
    out = input.new_empty((3 * R, C, pooled_h, pooled_w))
    arg = input.new_empty((3 * R, C, pooled_h, pooled_w) if with_argmax else (0,), dtype=torch.int32)
    return out, arg


def _roi_loop_pool_bwd(ctx, grad_out, _grad_arg):
    rois, argmax, row_scale = ctx.saved_tensors
    if not ctx.with_argmax:
        raise RuntimeError("wsovod_b200::roi_loop_pool: backward needs with_argmax=True")
    g = grad_out
    if ctx.has_scale:
        s = (row_scale + ctx.bias)
        g = g * torch.cat([s, s, s]).view(-1, 1, 1, 1)
    N, C, H, W = ctx.shape
    gi = _call(_roi_pool_backward, g, rois, argmax, N, C, H, W, True)
    return gi, None, None, None, None, None, None, None


_roi_loop_pool.register_autograd(_roi_loop_pool_bwd, setup_context=_roi_pool_setup)
_FAST_FN[_roi_loop_pool] = _fast_function(_roi_loop_pool, _roi_pool_setup, _roi_loop_pool_bwd)


def roi_loop_pool(input, rois, spatial_scale, output_size, row_scale=None, row_scale_bias=0.0,
                  with_argmax=True):
    """wsovod._C.roi_loop_pool_forward semantics: (3R,C,P,P) = roi | frame | context."""
    ph, pw = _pair(output_size)
    return _call(_roi_loop_pool, input, rois, float(spatial_scale), ph, pw, row_scale,
                                               float(row_scale_bias), bool(with_argmax))


@torch.library.custom_op("wsovod_b200::roi_align", mutates_args=())
def _roi_align(input: torch.Tensor, rois: torch.Tensor, spatial_scale: float, pooled_h: int,
               pooled_w: int, sampling_ratio: int, aligned: bool, row_scale: Optional[torch.Tensor],
               row_scale_bias: float) -> torch.Tensor:
    _need_cuda(input, rois, row_scale)
    input, rois = _f32c(input), _f32c(rois)
    if rois.dim() != 2 or rois.size(1) != 5 or input.dim() != 4:
        raise RuntimeError("wsovod_b200::roi_align expects input NCHW and rois (R,5)")
    N, C, H, W = input.shape
    R = rois.size(0)
    rs = None if row_scale is None else _f32c(row_scale)
    with torch.cuda.device(input.device):
        out = torch.empty((R, C, pooled_h, pooled_w), dtype=torch.float32, device=input.device)
        L = _lib.lib()
        ws = _workspace(L.wsovod_b200_roi_align_workspace_hw(N, R, pooled_h, pooled_w, H, W), input.device)
        rc = L.wsovod_b200_roi_align_fwd(_ptr(input), N, C, H, W, _ptr(rois), R, spatial_scale, pooled_h,
                                         pooled_w, sampling_ratio, int(aligned), _ptr(rs), row_scale_bias,
                                         _ptr(out), _ptr(ws), ws.numel(), _stream(input))
    _lib.check(rc, "roi_align_fwd")
    return out


@_roi_align.register_fake
def _(input, rois, spatial_scale, pooled_h, pooled_w, sampling_ratio, aligned, row_scale, row_scale_bias):
    return input.new_empty((rois.size(0), input.size(1), pooled_h, pooled_w))


@torch.library.custom_op("wsovod_b200::roi_align_backward", mutates_args=())
def _roi_align_backward(grad: torch.Tensor, rois: torch.Tensor, spatial_scale: float, sampling_ratio: int, aligned: bool,
                        N: int, C: int, H: int, W: int) -> torch.Tensor:
    _need_cuda(grad, rois)
    grad, rois = _f32c(grad), _f32c(rois)
    ph, pw = grad.shape[-2:]
    with torch.cuda.device(grad.device):
        gi = torch.zeros((N, C, H, W), dtype=torch.float32, device=grad.device)
        rc = _lib.lib().wsovod_b200_roi_align_bwd(_ptr(grad), _ptr(rois), rois.size(0), N, C, H, W, spatial_scale, ph, pw,
                                                  sampling_ratio, int(aligned), _ptr(gi), _stream(grad))
    _lib.check(rc, "roi_align_bwd")
    return gi


@_roi_align_backward.register_fake
def _(grad, rois, spatial_scale, sampling_ratio, aligned, N, C, H, W):
    return grad.new_empty((N, C, H, W))


def _roi_align_setup(ctx, inputs, output):
    input, rois, spatial_scale, ph, pw, sampling_ratio, aligned, row_scale, bias = inputs
    ctx.shape = tuple(input.shape)
    ctx.cfg = (spatial_scale, sampling_ratio, aligned, bias, row_scale is not None)
    ctx.save_for_backward(rois, row_scale)


def _roi_align_bwd(ctx, grad_out):
    rois, row_scale = ctx.saved_tensors
    spatial_scale, sampling_ratio, aligned, bias, has_scale = ctx.cfg
    g = grad_out * (row_scale + bias).view(-1, 1, 1, 1) if has_scale else grad_out
    N, C, H, W = ctx.shape
    gi = _call(_roi_align_backward, g, rois, spatial_scale, sampling_ratio, aligned, N, C, H, W)
    return gi, None, None, None, None, None, None, None, None


_roi_align.register_autograd(_roi_align_bwd, setup_context=_roi_align_setup)
_FAST_FN[_roi_align] = _fast_function(_roi_align, _roi_align_setup, _roi_align_bwd)


def roi_align(input, rois, spatial_scale, output_size, sampling_ratio=0, aligned=False, row_scale=None,
              row_scale_bias=0.0):
    ph, pw = _pair(output_size)
    return _call(_roi_align, input, rois, float(spatial_scale), ph, pw, int(sampling_ratio),
                                           bool(aligned), row_scale, float(row_scale_bias))


# ------------------------------------------------------------------------------------------------
# (2) alignment + MIL
# ------------------------------------------------------------------------------------------------
@torch.library.custom_op("wsovod_b200::align", mutates_args=())
def _align(x: torch.Tensor, classifier: torch.Tensor, temperature: float, norm_weight: int,
           append_background: bool, bias: Optional[torch.Tensor], precision: int,
           want_logits: bool, want_probs: bool) -> Tuple[torch.Tensor, torch.Tensor]:
    _need_cuda(x, classifier, bias)
    x, classifier = _f32c(x), _f32c(classifier)
    if x.dim() != 2 or classifier.dim() != 2 or x.size(1) != classifier.size(1):
        raise RuntimeError("wsovod_b200::align expects x (M,D) and classifier (K,D)")
    if not (want_logits or want_probs):
        raise RuntimeError("wsovod_b200::align: nothing to compute")
    M, D = x.shape
    K = classifier.size(0)
    KO = K + (1 if append_background else 0)
    b = None if bias is None else _f32c(bias)
    with torch.cuda.device(x.device):
        logits = torch.empty((M, KO) if want_logits else (0,), dtype=torch.float32, device=x.device)
        probs = torch.empty((M, KO) if want_probs else (0,), dtype=torch.float32, device=x.device)
        L = _lib.lib()
        ws = _workspace(L.wsovod_b200_align_workspace(M, D, K, precision), x.device)
        rc = L.wsovod_b200_align_fwd(_ptr(x), _ptr(classifier), M, D, K, temperature, int(norm_weight),
                                     int(append_background), _ptr(b), precision,
                                     _ptr(logits if want_logits else None),
                                     _ptr(probs if want_probs else None), _ptr(ws), ws.numel(), _stream(x))
    _lib.check(rc, "align_fwd")
    return logits, probs


@_align.register_fake
def _(x, classifier, temperature, norm_weight, append_background, bias, precision, want_logits, want_probs):
    KO = classifier.size(0) + (1 if append_background else 0)
    return (x.new_empty((x.size(0), KO) if want_logits else (0,)),
            x.new_empty((x.size(0), KO) if want_probs else (0,)))


@torch.library.custom_op("wsovod_b200::align_backward", mutates_args=())
def _align_backward(grad_logits: torch.Tensor, x: torch.Tensor, classifier: torch.Tensor,
                    temperature: float, norm_weight: int, append_background: bool, need_x: bool,
                    need_classifier: bool) -> Tuple[torch.Tensor, torch.Tensor]:
    _need_cuda(grad_logits, x, classifier)
    grad_logits, x, classifier = _f32c(grad_logits), _f32c(x), _f32c(classifier)
    M, D = x.shape
    K = classifier.size(0)
    with torch.cuda.device(x.device):
        gx = torch.empty_like(x) if need_x else x.new_empty((0,))
        gw = torch.empty_like(classifier) if need_classifier else x.new_empty((0,))
        L = _lib.lib()
        ws = _workspace(L.wsovod_b200_align_bwd_workspace(M, D, K), x.device)
        rc = L.wsovod_b200_align_bwd(_ptr(grad_logits), _ptr(x), _ptr(classifier), M, D, K, temperature,
                                     int(norm_weight), int(append_background), _ptr(gx if need_x else None),
                                     _ptr(gw if need_classifier else None), _ptr(ws), ws.numel(), _stream(x))
    _lib.check(rc, "align_bwd")
    return gx, gw


@_align_backward.register_fake
def _(grad_logits, x, classifier, temperature, norm_weight, append_background, need_x, need_classifier):
    return (torch.empty_like(x) if need_x else x.new_empty((0,)),
            torch.empty_like(classifier) if need_classifier else x.new_empty((0,)))


def _align_setup(ctx, inputs, output):
    x, classifier, temperature, norm_weight, append_background, bias, precision, wl, wp = inputs
    ctx.save_for_backward(x, classifier, output[1] if wp else None)
    ctx.cfg = (temperature, norm_weight, append_background, bias is not None, wl, wp)


def _align_bwd(ctx, g_logits, g_probs):
    x, classifier, probs = ctx.saved_tensors
    temperature, norm_weight, append_background, has_bias, wl, wp = ctx.cfg
    g = g_logits if wl and g_logits is not None else None
    if wp and g_probs is not None and probs is not None:
        gp = probs * (g_probs - (g_probs * probs).sum(-1, keepdim=True))   # softmax Jacobian (tiny, MxK)
        g = gp if g is None else g + gp
    if g is None:
        return (None,) * 9
    need_x, need_w = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
    gx = gw = None
    if need_x or need_w:
        gx, gw = _call(_align_backward, g, x, classifier, temperature, norm_weight, append_background,
                                                      bool(need_x), bool(need_w))
        gx, gw = (gx if need_x else None), (gw if need_w else None)
    gb = g.sum().reshape(1) if has_bias else None
    return gx, gw, None, None, None, gb, None, None, None


_align.register_autograd(_align_bwd, setup_context=_align_setup)
_FAST_FN[_align] = _fast_function(_align, _align_setup, _align_bwd)


def align(x, classifier, temperature=50.0, norm_weight=True, append_background=True, bias=None,
          precision=ALIGN_TF32, want_logits=True, want_probs=False):
    """Contraction part of OpenVocabularyClassifier.forward (+ optional fused row softmax).
    Returns (logits, probs); a tensor that was not requested is empty."""
    return _call(_align, x, classifier, float(temperature), int(norm_weight),
                                       bool(append_background), bias, int(precision), bool(want_logits),
                                       bool(want_probs))


def _offsets_tensor(sizes, device):
    off = [0]
    for s in sizes:
        off.append(off[-1] + int(s))
    return torch.tensor(off, dtype=torch.int64, device=device), off


@torch.library.custom_op("wsovod_b200::mil", mutates_args=())
def _mil(cls: torch.Tensor, det: torch.Tensor, offsets: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    _need_cuda(cls, det, offsets)
    cls, det = _f32c(cls), _f32c(det)
    if cls.shape != det.shape or cls.dim() != 2 or offsets.dtype != torch.int64:
        raise RuntimeError("wsovod_b200::mil expects cls/det (M,K) and int64 offsets (N+1)")
    M, K = cls.shape
    N = offsets.numel() - 1
    with torch.cuda.device(cls.device):
        scores = torch.empty_like(cls)
        img = torch.empty((N, K), dtype=torch.float32, device=cls.device)
        L = _lib.lib()
        ws = _workspace(L.wsovod_b200_mil_workspace(M, N, K), cls.device)
        rc = L.wsovod_b200_mil_fwd(_ptr(cls), _ptr(det), _ptr(offsets), M, N, K, _ptr(scores), _ptr(img),
                                   _ptr(ws), ws.numel(), _stream(cls))
    _lib.check(rc, "mil_fwd")
    return scores, img


@_mil.register_fake
def _(cls, det, offsets):
    return torch.empty_like(cls), cls.new_empty((offsets.numel() - 1, cls.size(1)))


@torch.library.custom_op("wsovod_b200::mil_backward", mutates_args=())
def _mil_backward(grad_scores: Optional[torch.Tensor], grad_img: Optional[torch.Tensor], cls: torch.Tensor,
                  det: torch.Tensor, offsets: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    _need_cuda(cls, det, offsets, grad_scores, grad_img)
    cls, det = _f32c(cls), _f32c(det)
    gs = None if grad_scores is None else _f32c(grad_scores)
    gi = None if grad_img is None else _f32c(grad_img)
    M, K = cls.shape
    N = offsets.numel() - 1
    with torch.cuda.device(cls.device):
        gc, gd = torch.zeros_like(cls), torch.zeros_like(det)
        L = _lib.lib()
        ws = _workspace(L.wsovod_b200_mil_workspace(M, N, K), cls.device)
        rc = L.wsovod_b200_mil_bwd(_ptr(gs), _ptr(gi), _ptr(cls), _ptr(det), _ptr(offsets), M, N, K, _ptr(gc),
                                   _ptr(gd), _ptr(ws), ws.numel(), _stream(cls))
    _lib.check(rc, "mil_bwd")
    return gc, gd


@_mil_backward.register_fake
def _(grad_scores, grad_img, cls, det, offsets):
    return torch.empty_like(cls), torch.empty_like(det)


def _mil_setup(ctx, inputs, output):
    cls, det, offsets = inputs
    ctx.save_for_backward(cls, det, offsets, output[1])


def _mil_bwd(ctx, g_scores, g_img):
    cls, det, offsets, img = ctx.saved_tensors
    if g_img is not None:   # clamp(min=1e-6, max=1-1e-6) passes gradient only strictly inside
        g_img = g_img * ((img > 1e-6) & (img < 1.0 - 1e-6)).to(g_img.dtype)
    gc, gd = _call(_mil_backward, g_scores, g_img, cls, det, offsets)
    return gc, gd, None


_mil.register_autograd(_mil_bwd, setup_context=_mil_setup)
_FAST_FN[_mil] = _fast_function(_mil, _mil_setup, _mil_bwd)


def mil(cls, det, offsets):
    """Two-stream MIL scores (M,K) and clamped image-level scores (N,K); offsets: int64 (N+1) on GPU."""
    return _call(_mil, cls, det, offsets)


@torch.library.custom_op("wsovod_b200::align_mil", mutates_args=())
def _align_mil(x: torch.Tensor, classifier: torch.Tensor, det: torch.Tensor, offsets: torch.Tensor, temperature: float,
               norm_weight: int, bias: Optional[torch.Tensor], want_logits: bool) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    _need_cuda(x, classifier, det, offsets, bias)
    x, classifier, det = _f32c(x), _f32c(classifier), _f32c(det)
    if x.dim() != 2 or classifier.dim() != 2 or x.size(1) != classifier.size(1) or det.shape != (x.size(0), classifier.size(0)) \
            or offsets.dtype != torch.int64:
        raise RuntimeError("wsovod_b200::align_mil expects x (M,D), classifier (K,D), det (M,K), int64 offsets (N+1)")
    M, D = x.shape
    K = classifier.size(0)
    N = offsets.numel() - 1
    b = None if bias is None else _f32c(bias)
    with torch.cuda.device(x.device):
        scores = torch.empty((M, K), dtype=torch.float32, device=x.device)
        img = torch.empty((N, K), dtype=torch.float32, device=x.device)
        logits = torch.empty((M, K) if want_logits else (0,), dtype=torch.float32, device=x.device)
        L = _lib.lib()
        ws = _workspace(L.wsovod_b200_align_mil_fused_workspace(M, N, D, K), x.device)
        rc = L.wsovod_b200_align_mil_fused_fwd(_ptr(x), _ptr(classifier), _ptr(det), _ptr(offsets), M, N, D, K, temperature,
                                               int(norm_weight), _ptr(b), _ptr(scores), _ptr(img),
                                               _ptr(logits if want_logits else None), _ptr(ws), ws.numel(), _stream(x))
    _lib.check(rc, "align_mil_fused_fwd")
    return scores, img, logits


@_align_mil.register_fake
def _(x, classifier, det, offsets, temperature, norm_weight, bias, want_logits):
    M, K = det.shape
    return x.new_empty((M, K)), x.new_empty((offsets.numel() - 1, K)), x.new_empty((M, K) if want_logits else (0,))


def _align_mil_setup(ctx, inputs, output):
    x, classifier, det, offsets, temperature, norm_weight, bias, want_logits = inputs
    ctx.save_for_backward(x, classifier, det, offsets, output[1], output[2])
    ctx.cfg = (temperature, norm_weight, bias is not None, want_logits)
    ctx.mark_non_differentiable(output[2])


def _align_mil_bwd(ctx, g_scores, g_img, _g_logits):
    x, classifier, det, offsets, img, logits = ctx.saved_tensors
    temperature, norm_weight, has_bias, want_logits = ctx.cfg
    if not want_logits:
        raise RuntimeError("wsovod_b200::align_mil: backward needs want_logits=True in the forward call")
    if g_scores is None and g_img is None:
        return (None,) * 8
    if g_img is not None:   # clamp(min=1e-6, max=1-1e-6) passes gradient only strictly inside
        g_img = g_img * ((img > 1e-6) & (img < 1.0 - 1e-6)).to(g_img.dtype)
    gc, gd = _call(_mil_backward, g_scores, g_img, logits, det, offsets)
    need_x, need_w = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
    gx = gw = None
    if need_x or need_w:
        gx, gw = _call(_align_backward, gc, x, classifier, temperature, norm_weight, False, bool(need_x),
                                                      bool(need_w))
        gx, gw = (gx if need_x else None), (gw if need_w else None)
    gb = gc.sum().reshape(1) if has_bias else None
    return gx, gw, (gd if ctx.needs_input_grad[2] else None), None, None, None, gb, None


_align_mil.register_autograd(_align_mil_bwd, setup_context=_align_mil_setup)
_FAST_FN[_align_mil] = _fast_function(_align_mil, _align_mil_setup, _align_mil_bwd)


def align_mil(x, classifier, det, offsets, temperature=50.0, norm_weight=True, bias=None, want_logits=None):
    """Fused alignment + MIL (north star kernel 2): scores (M,K) = softmax_k(T normalize(x) normalize(W)^T) *
    per-image softmax_r(det), image scores (N,K); the alignment logits come back too when ``want_logits`` (default:
    whenever autograd will need them).  TF32 contraction, K <= 256."""
    if want_logits is None:
        want_logits = torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in (x, classifier, det, bias))
    s, img, lg = _call(_align_mil, x, classifier, det, offsets, float(temperature), int(norm_weight), bias,
                                                 bool(want_logits))
    return s, img, lg


# ------------------------------------------------------------------------------------------------
# (3) refinement
# ------------------------------------------------------------------------------------------------
@torch.library.custom_op("wsovod_b200::pgt_top1", mutates_args=())
def _pgt_top1(scores: torch.Tensor, boxes: torch.Tensor, offsets: torch.Tensor, gt_classes: torch.Tensor,
              gt_offsets: torch.Tensor, img_scores: torch.Tensor) -> List[torch.Tensor]:
    _need_cuda(scores, boxes, offsets, gt_classes, gt_offsets, img_scores)
    scores, boxes, img_scores = _f32c(scores), _f32c(boxes), _f32c(img_scores)
    gt_classes = gt_classes.to(torch.int64).contiguous()
    M = boxes.size(0)
    N = offsets.numel() - 1
    K = img_scores.size(1)
    G = gt_classes.numel()
    dev = boxes.device
    with torch.cuda.device(dev):
        sb = torch.zeros((G, 4), dtype=torch.float32, device=dev)
        sc = torch.zeros((G,), dtype=torch.int64, device=dev)
        ss = torch.zeros((G,), dtype=torch.float32, device=dev)
        sw = torch.zeros((G,), dtype=torch.float32, device=dev)
        sr = torch.full((G,), -1, dtype=torch.int64, device=dev)
        cnt = torch.zeros((N,), dtype=torch.int64, device=dev)
        rc = _lib.lib().wsovod_b200_pgt_top1(_ptr(scores), scores.size(1), _ptr(boxes), _ptr(offsets),
                                             _ptr(gt_classes), _ptr(gt_offsets), _ptr(img_scores), M, N, K, G,
                                             _ptr(sb), _ptr(sc), _ptr(ss), _ptr(sw), _ptr(sr), _ptr(cnt),
                                             _stream(boxes))
    _lib.check(rc, "pgt_top1")
    return [sb, sc, ss, sw, sr, cnt]


@_pgt_top1.register_fake
def _(scores, boxes, offsets, gt_classes, gt_offsets, img_scores):
    G, N = gt_classes.numel(), offsets.numel() - 1
    i64 = dict(dtype=torch.int64)
    return [boxes.new_empty((G, 4)), boxes.new_empty((G,), **i64), boxes.new_empty((G,)),
            boxes.new_empty((G,)), boxes.new_empty((G,), **i64), boxes.new_empty((N,), **i64)]


def pgt_top1(scores, boxes, offsets, gt_classes, gt_offsets, img_scores):
    sb, sc, ss, sw, sr, cnt = _call(_pgt_top1, scores, boxes, offsets, gt_classes, gt_offsets,
                                                             img_scores)
    return dict(seed_boxes=sb, seed_classes=sc, seed_scores=ss, seed_weights=sw, seed_rows=sr, seed_count=cnt)


@torch.library.custom_op("wsovod_b200::refine_assign", mutates_args=())
def _refine_assign(boxes: torch.Tensor, offsets: torch.Tensor, seed_boxes: torch.Tensor,
                   seed_classes: torch.Tensor, seed_scores: torch.Tensor, seed_weights: torch.Tensor,
                   seed_offsets: torch.Tensor, seed_count: Optional[torch.Tensor], num_classes: int,
                   iou_thresh: float) -> List[torch.Tensor]:
    _need_cuda(boxes, offsets, seed_boxes, seed_classes, seed_scores, seed_weights, seed_offsets, seed_count)
    boxes, seed_boxes = _f32c(boxes), _f32c(seed_boxes)
    seed_scores, seed_weights = _f32c(seed_scores), _f32c(seed_weights)
    seed_classes = seed_classes.to(torch.int64).contiguous()
    M = boxes.size(0)
    N = offsets.numel() - 1
    dev = boxes.device
    with torch.cuda.device(dev):
        midx = torch.empty((M,), dtype=torch.int64, device=dev)
        mlab = torch.empty((M,), dtype=torch.int8, device=dev)
        miou = torch.empty((M,), dtype=torch.float32, device=dev)
        gcls = torch.empty((M,), dtype=torch.int64, device=dev)
        gbox = torch.empty((M, 4), dtype=torch.float32, device=dev)
        gsc = torch.empty((M,), dtype=torch.float32, device=dev)
        gw = torch.empty((M,), dtype=torch.float32, device=dev)
        rc = _lib.lib().wsovod_b200_refine_assign(_ptr(boxes), _ptr(offsets), _ptr(seed_boxes), _ptr(seed_classes),
                                                  _ptr(seed_scores), _ptr(seed_weights), _ptr(seed_offsets),
                                                  _ptr(seed_count), M, N, num_classes, iou_thresh, _ptr(midx),
                                                  _ptr(mlab), _ptr(miou), _ptr(gcls), _ptr(gbox), _ptr(gsc),
                                                  _ptr(gw), _stream(boxes))
    _lib.check(rc, "refine_assign")
    return [midx, mlab, miou, gcls, gbox, gsc, gw]


@_refine_assign.register_fake
def _(boxes, offsets, seed_boxes, seed_classes, seed_scores, seed_weights, seed_offsets, seed_count,
      num_classes, iou_thresh):
    M = boxes.size(0)
    return [boxes.new_empty((M,), dtype=torch.int64), boxes.new_empty((M,), dtype=torch.int8),
            boxes.new_empty((M,)), boxes.new_empty((M,), dtype=torch.int64), boxes.new_empty((M, 4)),
            boxes.new_empty((M,)), boxes.new_empty((M,))]


def refine_assign(boxes, offsets, seed_boxes, seed_classes, seed_scores, seed_weights, seed_offsets,
                  seed_count, num_classes, iou_thresh=0.5):
    r = _call(_refine_assign, boxes, offsets, seed_boxes, seed_classes, seed_scores, seed_weights,
                                            seed_offsets, seed_count, int(num_classes), float(iou_thresh))
    return dict(matched_idx=r[0], matched_label=r[1], matched_iou=r[2], gt_classes=r[3], gt_boxes=r[4],
                gt_scores=r[5], gt_weights=r[6])


# ------------------------------------------------------------------------------------------------
# (3b) weighted refinement losses
# ------------------------------------------------------------------------------------------------
@torch.library.custom_op("wsovod_b200::refine_losses", mutates_args=())
def _refine_losses(logits: torch.Tensor, deltas: Optional[torch.Tensor], gt_classes: torch.Tensor,
                   gt_weights: torch.Tensor, proposal_boxes: Optional[torch.Tensor], gt_boxes: Optional[torch.Tensor],
                   num_classes: int, wx: float, wy: float, ww: float, wh: float,
                   beta: float) -> Tuple[torch.Tensor, torch.Tensor]:
    _need_cuda(logits, deltas, gt_classes, gt_weights, proposal_boxes, gt_boxes)
    logits, gt_weights = _f32c(logits), _f32c(gt_weights)
    gt_classes = gt_classes.to(torch.int64).contiguous()
    if logits.dim() != 2 or gt_classes.numel() != logits.size(0) or gt_weights.numel() != logits.size(0):
        raise RuntimeError("wsovod_b200::refine_losses expects logits (M,K+1), gt_classes (M), gt_weights (M)")
    M, K1 = logits.shape
    dcols = 0
    if deltas is not None:
        deltas, proposal_boxes, gt_boxes = _f32c(deltas), _f32c(proposal_boxes), _f32c(gt_boxes)
        dcols = deltas.size(1) if deltas.dim() == 2 else -1
        if deltas.size(0) != M or proposal_boxes.shape != (M, 4) or gt_boxes.shape != (M, 4):
            raise RuntimeError("wsovod_b200::refine_losses expects deltas (M,4|4K) and boxes (M,4)")
    dev = logits.device
    with torch.cuda.device(dev):
        out = torch.empty((4,), dtype=torch.float32, device=dev)
        lse = torch.empty((M,), dtype=torch.float32, device=dev)
        L = _lib.lib()
        ws = _workspace(L.wsovod_b200_refine_loss_workspace(M), dev)
        rc = L.wsovod_b200_refine_loss_fwd(_ptr(logits), K1, _ptr(gt_classes), _ptr(gt_weights), _ptr(proposal_boxes),
                                           _ptr(gt_boxes), _ptr(deltas), dcols, M, num_classes, wx, wy, ww, wh, beta,
                                           _ptr(out), _ptr(lse), _ptr(ws), ws.numel(), _stream(logits))
    _lib.check(rc, "refine_loss_fwd")
    return out, lse


@_refine_losses.register_fake
def _(logits, deltas, gt_classes, gt_weights, proposal_boxes, gt_boxes, num_classes, wx, wy, ww, wh, beta):
    return logits.new_empty((4,)), logits.new_empty((logits.size(0),))


@torch.library.custom_op("wsovod_b200::refine_losses_backward", mutates_args=())
def _refine_losses_backward(grad_out: torch.Tensor, fwd_out: torch.Tensor, lse: torch.Tensor, logits: torch.Tensor,
                            deltas: Optional[torch.Tensor], gt_classes: torch.Tensor, gt_weights: torch.Tensor,
                            proposal_boxes: Optional[torch.Tensor], gt_boxes: Optional[torch.Tensor],
                            num_classes: int, wx: float, wy: float, ww: float, wh: float,
                            beta: float) -> Tuple[torch.Tensor, torch.Tensor]:
    _need_cuda(grad_out, fwd_out, lse, logits, deltas, gt_classes, gt_weights, proposal_boxes, gt_boxes)
    logits, gt_weights, grad_out = _f32c(logits), _f32c(gt_weights), _f32c(grad_out)
    gt_classes = gt_classes.to(torch.int64).contiguous()
    M, K1 = logits.shape
    dcols = 0
    if deltas is not None:
        deltas, proposal_boxes, gt_boxes = _f32c(deltas), _f32c(proposal_boxes), _f32c(gt_boxes)
        dcols = deltas.size(1)
    dev = logits.device
    with torch.cuda.device(dev):
        gl = torch.empty_like(logits)
        gd = torch.empty_like(deltas) if deltas is not None else logits.new_empty((0,))
        L = _lib.lib()
        rc = L.wsovod_b200_refine_loss_bwd(_ptr(grad_out), _ptr(fwd_out), _ptr(logits), K1, _ptr(lse), _ptr(gt_classes),
                                           _ptr(gt_weights), _ptr(proposal_boxes), _ptr(gt_boxes), _ptr(deltas), dcols,
                                           M, num_classes, wx, wy, ww, wh, beta, _ptr(gl),
                                           _ptr(gd) if deltas is not None else None, _stream(logits))
    _lib.check(rc, "refine_loss_bwd")
    return gl, gd


@_refine_losses_backward.register_fake
def _(grad_out, fwd_out, lse, logits, deltas, gt_classes, gt_weights, proposal_boxes, gt_boxes, num_classes, wx, wy,
      ww, wh, beta):
    return torch.empty_like(logits), (torch.empty_like(deltas) if deltas is not None else logits.new_empty((0,)))


def _refine_losses_setup(ctx, inputs, output):
    logits, deltas, gt_classes, gt_weights, proposal_boxes, gt_boxes, num_classes, wx, wy, ww, wh, beta = inputs
    ctx.has_deltas = deltas is not None
    ctx.consts = (num_classes, wx, wy, ww, wh, beta)
    saved = [output[0], output[1], logits, gt_classes, gt_weights]
    if ctx.has_deltas:
        saved += [deltas, proposal_boxes, gt_boxes]
    ctx.save_for_backward(*saved)


def _refine_losses_bwd(ctx, g_out, g_lse):
    out, lse, logits, gt_classes, gt_weights = ctx.saved_tensors[:5]
    deltas, pb, gb = ctx.saved_tensors[5:] if ctx.has_deltas else (None, None, None)
    gl, gd = _call(_refine_losses_backward, g_out[:2].contiguous(), out, lse, logits, deltas, gt_classes,
                                                          gt_weights, pb, gb, *ctx.consts)
    return (gl, gd if ctx.has_deltas else None) + (None,) * 10


_refine_losses.register_autograd(_refine_losses_bwd, setup_context=_refine_losses_setup)
_FAST_FN[_refine_losses] = _fast_function(_refine_losses, _refine_losses_setup, _refine_losses_bwd)


def refine_losses(logits, deltas, gt_classes, gt_weights, proposal_boxes=None, gt_boxes=None, num_classes=None,
                  box_weights=(10.0, 10.0, 5.0, 5.0), smooth_l1_beta=0.0):
    """InstanceRefinementOutputLayers.losses with cross_entropy_weighted + "smooth_l1_weighted"
    (fast_rcnn_open_vocabulary.py:754-892): returns (loss_cls, loss_box_reg) as 0-d tensors with autograd
    to `logits` and `deltas`; `deltas=None` (refine_reg off) gives loss_box_reg = 0."""
    K = int(num_classes) if num_classes is not None else logits.size(1) - 1
    wx, wy, ww, wh = (float(v) for v in box_weights)
    out, _ = _call(_refine_losses, logits, deltas, gt_classes, gt_weights, proposal_boxes, gt_boxes, K,
                                                 wx, wy, ww, wh, float(smooth_l1_beta))
    return out[0], out[1]


# ------------------------------------------------------------------------------------------------
# (4) NMS
# ------------------------------------------------------------------------------------------------
@torch.library.custom_op("wsovod_b200::batched_nms", mutates_args=())
def _batched_nms(boxes: torch.Tensor, scores: torch.Tensor, groups: torch.Tensor, num_groups: int,
                 iou_thresh: float, iou_mode: int) -> Tuple[torch.Tensor, torch.Tensor]:
    _need_cuda(boxes, scores, groups)
    boxes, scores = _f32c(boxes), _f32c(scores)
    groups = groups.to(torch.int64).contiguous()
    M = boxes.size(0)
    dev = boxes.device
    with torch.cuda.device(dev):
        keep = torch.empty((M,), dtype=torch.int64, device=dev)
        num = torch.zeros((1,), dtype=torch.int64, device=dev)
        L = _lib.lib()
        ws = _workspace(L.wsovod_b200_batched_nms_workspace(M, num_groups), dev)
        rc = L.wsovod_b200_batched_nms(_ptr(boxes), _ptr(scores), _ptr(groups), M, num_groups, iou_thresh,
                                       iou_mode, _ptr(keep), _ptr(num), _ptr(ws), ws.numel(), _stream(boxes))
    _lib.check(rc, "batched_nms")
    return keep, num


@_batched_nms.register_fake
def _(boxes, scores, groups, num_groups, iou_thresh, iou_mode):
    return boxes.new_empty((boxes.size(0),), dtype=torch.int64), boxes.new_empty((1,), dtype=torch.int64)


def batched_nms(boxes, scores, idxs, iou_threshold, iou_mode=IOU_TV_CUDA):
    """detectron2.layers.batched_nms semantics (vanilla per-class strategy on un-offset coordinates).
    `idxs` may hold arbitrary category ids; they are densified here.  Returns kept indices sorted by
    score (one device->host read for the count, like the reference's dynamic-shape result)."""
    _need_cuda(boxes, scores, idxs)
    if boxes.numel() == 0:
        return torch.empty((0,), dtype=torch.int64, device=boxes.device)
    uniq, dense = torch.unique(idxs, return_inverse=True)
    keep, num = _call(_batched_nms, boxes.float(), scores, dense, int(uniq.numel()),
                                                  float(iou_threshold), int(iou_mode))
    return keep[: int(num.item())]


@torch.library.custom_op("wsovod_b200::detections", mutates_args=())
def _detections(probs: torch.Tensor, boxes: torch.Tensor, offsets: torch.Tensor, image_sizes: torch.Tensor,
                max_rows: int, score_thresh: float, nms_thresh: float, topk: int,
                iou_mode: int) -> List[torch.Tensor]:
    _need_cuda(probs, boxes, offsets, image_sizes)
    probs, boxes, image_sizes = _f32c(probs), _f32c(boxes), _f32c(image_sizes)
    M, K1 = probs.shape
    K = K1 - 1
    N = offsets.numel() - 1
    dev = probs.device
    with torch.cuda.device(dev):
        db = torch.empty((N, topk, 4), dtype=torch.float32, device=dev)
        ds = torch.empty((N, topk), dtype=torch.float32, device=dev)
        dc = torch.empty((N, topk), dtype=torch.int64, device=dev)
        dr = torch.empty((N, topk), dtype=torch.int64, device=dev)
        cnt = torch.empty((N,), dtype=torch.int64, device=dev)
        L = _lib.lib()
        ws = _workspace(L.wsovod_b200_detections_workspace(M, N, K, topk), dev)
        rc = L.wsovod_b200_detections(_ptr(probs), _ptr(boxes), _ptr(offsets), _ptr(image_sizes), M, N, K,
                                      max_rows, score_thresh, nms_thresh, topk, iou_mode, _ptr(db), _ptr(ds),
                                      _ptr(dc), _ptr(dr), _ptr(cnt), _ptr(ws), ws.numel(), _stream(probs))
    _lib.check(rc, "detections")
    return [db, ds, dc, dr, cnt]


@_detections.register_fake
def _(probs, boxes, offsets, image_sizes, max_rows, score_thresh, nms_thresh, topk, iou_mode):
    N = offsets.numel() - 1
    i64 = dict(dtype=torch.int64)
    return [probs.new_empty((N, topk, 4)), probs.new_empty((N, topk)), probs.new_empty((N, topk), **i64),
            probs.new_empty((N, topk), **i64), probs.new_empty((N,), **i64)]


def detections(probs, boxes, offsets, image_sizes, max_rows, score_thresh, nms_thresh, topk,
               iou_mode=IOU_TV_CUDA):
    """Fused fast_rcnn_inference tail for class-agnostic boxes, batched over images (padded outputs)."""
    r = _call(_detections, probs, boxes, offsets, image_sizes, int(max_rows), float(score_thresh),
                                         float(nms_thresh), int(topk), int(iou_mode))
    return dict(det_boxes=r[0], det_scores=r[1], det_classes=r[2], det_rows=r[3], det_count=r[4])


# ------------------------------------------------------------------------------------------------
# CSC
# ------------------------------------------------------------------------------------------------
@torch.library.custom_op("wsovod_b200::csc", mutates_args=())
def _csc(cpgs: torch.Tensor, labels: torch.Tensor, preds: torch.Tensor, rois: torch.Tensor, fg_threshold: float,
         area_sqrt: bool, context_scale: float) -> torch.Tensor:
    _need_cuda(cpgs, labels, preds, rois)
    cpgs, labels, preds, rois = _f32c(cpgs), _f32c(labels), _f32c(preds), _f32c(rois)
    # the reference's CAFFE_ENFORCE_EQ shape checks (csc_cuda.cu:364-372)
    if cpgs.dim() != 4 or labels.dim() != 2 or preds.dim() != 2 or rois.dim() != 2 or rois.size(1) != 5 \
            or labels.shape != preds.shape or labels.shape != cpgs.shape[:2]:
        raise RuntimeError("wsovod_b200::csc expects cpgs (B,K,H,W), labels (B,K), preds (B,K), rois (R,5)")
    B, K, H, W = cpgs.shape
    R = rois.size(0)
    with torch.cuda.device(cpgs.device):
        out = torch.empty((R, K), dtype=torch.float32, device=cpgs.device)
        L = _lib.lib()
        ws = _workspace(L.wsovod_b200_csc_workspace(K, H, W), cpgs.device)
        rc = L.wsovod_b200_csc_fwd(_ptr(cpgs), _ptr(labels), _ptr(preds), _ptr(rois), B, K, H, W, R, fg_threshold,
                                   int(area_sqrt), context_scale, _ptr(out), _ptr(ws), ws.numel(), _stream(cpgs))
    _lib.check(rc, "csc_fwd")
    return out


@_csc.register_fake
def _(cpgs, labels, preds, rois, fg_threshold, area_sqrt, context_scale):
    return cpgs.new_empty((rois.size(0), cpgs.size(1)))


def csc(cpgs, labels, preds, rois, fg_threshold=0.1, area_sqrt=True, context_scale=1.8):
    """wsovod._C.csc_forward: W (R,K) from class peak response maps (B,K,H,W), image labels / predictions (B,K) and rois
    (R,5) in map pixels; not differentiable (csc.py:44-47)."""
    return _call(_csc, cpgs, labels, preds, rois, float(fg_threshold), bool(area_sqrt), float(context_scale))
