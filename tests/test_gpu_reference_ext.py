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


def test_reference_extension_rejects_cpu_like_ours():
    m = _ext()
    with pytest.raises(RuntimeError, match="Not compiled with CPU support"):
        m.roi_loop_pool_forward(torch.zeros(1, 1, 8, 8), torch.zeros(1, 5), 0.125, 7, 7)
