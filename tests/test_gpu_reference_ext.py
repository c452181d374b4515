"""GPU parity against the reference's OWN CUDA extension (oracle/_ref/wsovod_ref_C.so, compiled unmodified
from /root/reference/wsovod/layers by oracle/build_ref.py): ROILoopPool forward/backward, the one native
op of the reference on this path (wsovod/layers/vision.cpp:10-11)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import ref  # noqa: E402
from wsovod_b200 import ops, synth  # noqa: E402
from wsovod_b200.layers import ROILoopPool  # noqa: E402

DEV = "cuda:0"


def _ext():
    m = ref.cuda()
    if m is None:
        pytest.skip("oracle/_ref/wsovod_ref_C.so not built")
    return m


@pytest.mark.parametrize("N,C,H,W,R", [(2, 8, 60, 80, 600), (1, 3, 100, 152, 400), (3, 5, 23, 37, 200)])
def test_roi_loop_pool_matches_reference_extension(N, C, H, W, R):
    m = _ext()
    g = synth.gen(N * 100 + C)
    feat = synth.features(N, C, H, W, g).to(DEV)              # post-ReLU, like res5
    boxes = [synth.proposals(R, H * 8, W * 8, g) for _ in range(N)]
    rois, _ = synth.rois_from(boxes)
    rois = rois.to(DEV)
    ref_out, ref_arg = m.roi_loop_pool_forward(feat, rois, 1 / 8, 7, 7)
    out, arg = ops.roi_loop_pool(feat, rois, 1 / 8, 7)
    assert torch.equal(out, ref_out)
    assert torch.equal(arg, ref_arg)
    # module form, as poolers.py:187-190 builds it
    assert torch.equal(ROILoopPool((7, 7), 1 / 8)(feat, rois), ref_out)
    # backward (atomics: summation order differs -> tolerance)
    go = torch.randn_like(out)
    ref_gi = m.roi_loop_pool_backward(go, rois, ref_arg, 1 / 8, 7, 7, N, C, H, W)
    gi = torch.ops.wsovod_b200.roi_pool_backward(go, rois, arg, N, C, H, W, True)
    torch.testing.assert_close(gi, ref_gi, rtol=1e-5, atol=1e-4)


@pytest.mark.parametrize("dtype", [torch.float16, torch.float64])
@pytest.mark.parametrize("N,C,H,W,R,scale", [(2, 8, 60, 80, 600, 1 / 8), (1, 3, 100, 152, 400, 1 / 8), (3, 5, 23, 37, 300, 0.1),
                                             (1, 1, 200, 300, 200, 1 / 8), (2, 4, 40, 56, 500, 1 / 16)])
def test_roi_loop_pool_half_and_double_match_reference_extension(dtype, N, C, H, W, R, scale):
    """ROILoopPool_cuda.cu:294,364 dispatch float, double and half; the box arithmetic of the reference's template is
    rounded per dtype (csrc/roi_loop_dtype.cu states the rules).  Values and argmax bit-exact against the reference's
    own kernel on rois quantised to the dtype, scales that are not powers of two included; out-of-image boxes too."""
    m = _ext()
    g = synth.gen(N * 1000 + C + H)
    stride = round(1 / scale)
    feat = synth.features(N, C, H, W, g).to(DEV).to(dtype)
    boxes = [synth.proposals(R, H * stride, W * stride, g) for _ in range(N)]
    rois, _ = synth.rois_from(boxes)
    k = max(rois.size(0) // 10, 1)
    rois[:k, 1:] += torch.randn(k, 4, generator=g) * 150          # out-of-image / inverted boxes
    rois = rois.to(DEV).to(dtype)
    ref_out, ref_arg = m.roi_loop_pool_forward(feat, rois, scale, 7, 7)
    out, arg = ops.roi_loop_pool(feat, rois, scale, 7)
    assert out.dtype == dtype and torch.equal(out, ref_out)
    assert torch.equal(arg, ref_arg)
    out_v, arg_v = ops.roi_loop_pool(feat, rois, scale, 7, with_argmax=False)
    assert arg_v.numel() == 0 and torch.equal(out_v, ref_out)
    assert torch.equal(ROILoopPool((7, 7), scale)(feat, rois), ref_out)
    # row scale: the reference's separate multiply in the same dtype (roi_heads.py:733-739)
    obj = synth.objectness(rois.size(0), g).to(DEV).to(dtype)
    sc, _ = ops.roi_loop_pool(feat, rois, scale, 7, obj, 1.0, False)
    s3 = torch.cat([obj, obj, obj]) + 1
    assert torch.equal(sc, ref_out * s3.view(-1, 1, 1, 1))
    # backward: atomics in T on both sides, the order of the additions differs
    go = torch.randn(out.shape, device=DEV, generator=torch.Generator(DEV).manual_seed(3)).to(dtype)
    ref_gi = m.roi_loop_pool_backward(go, rois, ref_arg, scale, 7, 7, N, C, H, W)
    gi = torch.ops.wsovod_b200.roi_pool_backward(go, rois, arg, N, C, H, W, True)
    assert gi.dtype == dtype
    if dtype == torch.float64:
        torch.testing.assert_close(gi, ref_gi, rtol=1e-12, atol=1e-10)
    else:
        exact = torch.zeros(N * C * H * W, device=DEV, dtype=torch.float64)
        b = rois[:, 0].long().clamp(0, N - 1).repeat(3)
        keep = arg.view(-1) >= 0
        flat = ((b.view(-1, 1, 1, 1) * C + torch.arange(C, device=DEV).view(1, -1, 1, 1)) * (H * W) + arg.long()).view(-1)
        exact.index_add_(0, flat[keep], go.double().view(-1)[keep])
        exact = exact.view(N, C, H, W)
        tol = 0.02 * exact.abs() + 0.25        # ~sqrt(#terms) half roundings of the running sum
        assert bool(((gi.double() - exact).abs() <= tol).all()) and bool(((ref_gi.double() - exact).abs() <= tol).all())
    # autograd through the module
    x = feat.clone().requires_grad_(True)
    y = ROILoopPool((7, 7), scale)(x, rois)
    y.backward(go)
    assert x.grad is not None and x.grad.dtype == dtype
    if dtype == torch.float64:
        torch.testing.assert_close(x.grad, ref_gi, rtol=1e-12, atol=1e-10)


def test_roi_loop_pool_dtype_entry_fp32_equals_main_kernels():
    """the dtype entry point instantiated for float gives the fp32 kernels' results (values and argmax, bit for bit)"""
    from wsovod_b200 import _lib
    from wsovod_b200.ops import _ptr, _stream, _workspace
    g = synth.gen(77)
    N, C, H, W, R = 2, 6, 50, 70, 700
    feat = synth.features(N, C, H, W, g).to(DEV)
    rois, _ = synth.rois_from([synth.proposals(R, H * 8, W * 8, g) for _ in range(N)])
    rois[:50, 1:] += torch.randn(50, 4, generator=g) * 150
    rois = rois.to(DEV)
    out, arg = ops.roi_loop_pool(feat, rois, 1 / 8, 7)
    L = _lib.lib()
    out2, arg2 = torch.empty_like(out), torch.empty_like(arg)
    ws = _workspace(L.wsovod_b200_roi_loop_pool_dtype_workspace(N, rois.size(0), 7, 7), feat.device)
    rc = L.wsovod_b200_roi_loop_pool_dtype_fwd(_lib.F32, _ptr(feat), N, C, H, W, _ptr(rois), rois.size(0), 1 / 8, 7, 7,
                                               _ptr(out2), _ptr(arg2), _ptr(ws), ws.numel(), _stream(feat))
    _lib.check(rc, "roi_loop_pool_dtype_fwd")
    assert torch.equal(out2, out) and torch.equal(arg2, arg)
    with pytest.raises(RuntimeError, match="expected scalar type"):
        ops.roi_loop_pool(feat.half(), rois, 1 / 8, 7)            # rois.data_ptr<scalar_t>() in the reference


def test_reference_extension_rejects_cpu_like_ours():
    m = _ext()
    with pytest.raises(RuntimeError, match="Not compiled with CPU support"):
        m.roi_loop_pool_forward(torch.zeros(1, 1, 8, 8), torch.zeros(1, 5), 0.125, 7, 7)
