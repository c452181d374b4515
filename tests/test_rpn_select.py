"""find_top_rpn_proposals (wsovod_b200/modeling/proposal_utils.py) against the reference's own
function (tests/golden/rpn_select.pt).  On the CPU the NMS step is the oracle's batched_nms (torchvision CPU
arithmetic); on the GPU it is the library's kernel."""
import pytest
import torch

import oracle
from wsovod_b200 import ops
from wsovod_b200.modeling.proposal_utils import find_top_rpn_proposals


def _check(res, c):
    assert len(res) == len(c["boxes"])
    for r, gb, gs in zip(res, c["boxes"], c["scores"]):
        assert torch.equal(r.objectness_logits.cpu(), gs)
        assert torch.equal(r.proposal_boxes.tensor.cpu(), gb)


def test_rpn_selection_matches_reference_with_oracle_nms(golden):
    def nms(b, s, groups, thr):
        return oracle.batched_nms(b, s, groups, thr, oracle.IOU_TV_CPU)
    for name, c in golden("rpn_select").items():
        res = find_top_rpn_proposals(c["proposals"], c["logits"], c["image_sizes"], c["nms_thresh"], c["pre"], c["post"],
                                     c["min_box_size"], False, nms_fn=nms)
        _check(res, c)
        assert all(len(r.objectness_logits) <= c["post"] for r in res)
    c = golden("rpn_select")["single"]
    with pytest.raises(FloatingPointError):
        find_top_rpn_proposals(c["proposals"], c["logits"], c["image_sizes"], 0.7, c["pre"], c["post"], 0.0, True, nms_fn=nms)


@pytest.mark.gpu
def test_rpn_selection_matches_reference_on_gpu(golden):
    for name, c in golden("rpn_select").items():
        res = find_top_rpn_proposals([p.cuda() for p in c["proposals"]], [l.cuda() for l in c["logits"]], c["image_sizes"],
                                     c["nms_thresh"], c["pre"], c["post"], c["min_box_size"], False,
                                     nms_fn=lambda b, s, g, t: ops.batched_nms(b, s, g, t, ops.IOU_TV_CPU))
        _check(res, c)


def _check_group(res, c):
    for r, gb, gs, gl in zip(res, c["boxes"], c["scores"], c["level_ids"]):
        assert torch.equal(r.objectness_logits.cpu(), gs)
        assert torch.equal(r.proposal_boxes.tensor.cpu(), gb)
        assert torch.equal(r.level_ids.cpu(), gl)


def test_rpn_group_selection_matches_reference_with_oracle_kernels(golden):
    """find_top_rpn_proposals_group (proposal_utils.py:146-362): per-(level, anchor) groups and the CSC re-weighting"""
    from wsovod_b200.modeling.proposal_utils import find_top_rpn_proposals_group
    for name, c in golden("rpn_group").items():
        res = find_top_rpn_proposals_group(
            c["proposals"], c["logits"], c["image_sizes"], c["num_anchors"], c["nms_thresh"], c["pre"], c["post"], 0.0, False,
            c["cpgs"], c["cpg_strides"], nms_fn=lambda b, s, g, t: oracle.batched_nms(b, s, g, t, oracle.IOU_TV_CPU),
            csc_fn=lambda cp, l, p, r: oracle.csc(cp, l, p, r, 0.1, True, 1.8))
        _check_group(res, c)


@pytest.mark.gpu
def test_rpn_group_selection_on_gpu(golden):
    from wsovod_b200.modeling.proposal_utils import find_top_rpn_proposals_group
    for name, c in golden("rpn_group").items():
        cp = None if c["cpgs"] is None else [m.cuda() for m in c["cpgs"]]
        res = find_top_rpn_proposals_group(
            [p.cuda() for p in c["proposals"]], [l.cuda() for l in c["logits"]], c["image_sizes"], c["num_anchors"],
            c["nms_thresh"], c["pre"], c["post"], 0.0, False, cp, c["cpg_strides"],
            nms_fn=lambda b, s, g, t: ops.batched_nms(b, s, g, t, ops.IOU_TV_CPU))
        _check_group(res, c)
