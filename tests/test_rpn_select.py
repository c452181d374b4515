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
