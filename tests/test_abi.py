"""CPU tests of the drop-in boundary: the shared library loads, exports every symbol the header
declares, validates arguments without touching a GPU, and the Python ops refuse CPU tensors."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "wsovod_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(wsovod_b200_\w+)\s*\(", src)))


def test_header_symbols_exported():
    from wsovod_b200 import _lib
    names = _declared()
    assert len(names) >= 25
    L = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/wsovod_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == names, "wsovod_b200/_lib.py must bind exactly the declared ABI"


def test_version_strerror_and_argument_errors():
    from wsovod_b200 import _lib
    L = _lib.lib()
    assert L.wsovod_b200_abi_version() == 1
    assert b"workspace" in L.wsovod_b200_strerror(-3)
    assert L.wsovod_b200_strerror(0) == b"success"
    # argument validation happens before any CUDA call
    z = ctypes.c_void_p(0)
    assert L.wsovod_b200_roi_pool_fwd(z, 1, 4, 8, 8, z, 10, 0.125, 7, 7, z, 0.0, z, z, z, 0, z) == -1
    assert L.wsovod_b200_roi_pool_fwd(z, 1, 4, 8, 8, z, 0, 0.125, 7, 7, z, 0.0, z, z, z, 0, z) == 0   # R == 0
    assert L.wsovod_b200_roi_pool_fwd(z, 1, 4, 40000, 8, z, 10, 0.125, 7, 7, z, 0.0, z, z, z, 0, z) == -1
    assert L.wsovod_b200_batched_nms(z, z, z, -1, 1, 0.3, 0, z, z, z, 0, z) == -1
    assert L.wsovod_b200_align_fwd(z, z, 10, 8, 4, 50.0, 1, 1, z, 7, z, z, z, 0, z) == -1             # bad precision
    assert L.wsovod_b200_roi_pool_workspace(8, 32000, 7, 7) > 32000 * 28 * 2
    assert L.wsovod_b200_launch_count() == 0


def test_ops_reject_cpu_tensors_like_the_reference():
    from wsovod_b200 import ops
    from wsovod_b200.layers import ROILoopPool
    x = torch.zeros(1, 2, 8, 8)
    rois = torch.tensor([[0., 0., 0., 16., 16.]])
    with pytest.raises(RuntimeError, match="Not compiled with CPU support"):
        ops.roi_pool(x, rois, 0.125, 7)
    with pytest.raises(RuntimeError, match="Not compiled with CPU support"):
        ROILoopPool((7, 7), 0.125)(x, rois)
    with pytest.raises(AssertionError):
        ROILoopPool((7, 7), 0.125)(x, rois[:, :4])          # roi_loop_pool.py:50


def test_custom_ops_registered_with_fake_impls():
    from wsovod_b200 import ops  # noqa: F401
    for name in ("roi_pool", "roi_loop_pool", "roi_align", "align", "mil", "pgt_top1", "refine_assign",
                 "batched_nms", "detections"):
        assert hasattr(torch.ops.wsovod_b200, name)
    with torch._subclasses.fake_tensor.FakeTensorMode():
        x = torch.empty(2, 8, 20, 20, device="cuda")
        r = torch.empty(30, 5, device="cuda")
        out, arg = torch.ops.wsovod_b200.roi_pool(x, r, 0.125, 7, 7, None, 0.0, True)
        assert out.shape == (30, 8, 7, 7) and arg.dtype == torch.int32
        out3, _ = torch.ops.wsovod_b200.roi_loop_pool(x, r, 0.125, 7, 7, None, 0.0, True)
        assert out3.shape == (90, 8, 7, 7)


def test_apply_deltas_matches_box2box_restatement():
    """the PyTorch box decode kept outside the kernels == detectron2's Box2BoxTransform (oracle/d2_shim.py)"""
    from oracle.d2_shim import Box2BoxTransform
    from wsovod_b200.modeling.roi_heads import apply_deltas
    g = torch.Generator().manual_seed(3)
    boxes = torch.rand(200, 4, generator=g) * 300
    boxes[:, 2:] += boxes[:, :2]
    deltas = torch.randn(200, 4, generator=g)
    deltas[:5, 2:] = 50.0                      # hits the log(1000/16) clamp
    ref = Box2BoxTransform((10.0, 10.0, 5.0, 5.0)).apply_deltas(deltas, boxes)
    assert torch.equal(apply_deltas(deltas, boxes), ref)
