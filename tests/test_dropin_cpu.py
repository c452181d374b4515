"""The reference's own WSOVODROIHeads._forward_box (roi_heads.py:696-907), imported verbatim through
oracle/d2_shim.py and run on CPU, against the oracle's restatement of the same path composed step by step: pins
the oracle at PATH level (every seam between the per-step restatements), not only per function.  CPU only."""
import os
import sys

import pytest
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import dropin_harness as H  # noqa: E402
import oracle  # noqa: E402
from oracle import d2_shim  # noqa: E402

pytestmark = pytest.mark.skipif(d2_shim.default_reference_root() is None,
                                reason="no reference tree (/root/reference or oracle/_ref/py)")


def _cat_offsets(proposals):
    off = [0]
    for p in proposals:
        off.append(off[-1] + len(p))
    return off


@pytest.mark.parametrize("R,batch", [(300, 4096), (700, 512)])
def test_reference_forward_box_vs_oracle_composition(R, batch):
    mods = H.reference_modules()
    rh = mods[1]
    heads = H.build_reference("cpu", mods=mods, batch_size=batch)
    feats, props, targets, text = H.make_inputs("cpu", R=R)
    losses, _ = H.run_train(heads, rh, feats, props, targets, None, seed=11)
    K = heads.num_classes
    # ---- the same step out of the oracle's functions -------------------------------------------------------
    with torch.no_grad():
        rois = torch.cat([torch.cat([torch.full((len(p), 1), float(i)), p.proposal_boxes.tensor], 1) for i, p in enumerate(props)])
        obj = torch.cat([p.objectness_logits for p in props])
        pooled, _ = oracle.roi_pool(feats["res5"], rois, 1 / 8, 7)
        x = heads.box_head(pooled * (obj + 1).view(-1, 1, 1, 1))                 # roi_heads.py:733-746
        off = _cat_offsets(props)
        scores, img = oracle.mil(heads.object_miner.cls(x), heads.object_miner.det(x), off)
        gt_oh = heads.gt_classes_img_oh
        loss_mil = F.binary_cross_entropy(img, gt_oh, reduction="mean")
        gts = heads.gt_classes_img_int
        goff = [0]
        for g in gts:
            goff.append(goff[-1] + g.numel())
        boxes = rois[:, 1:].contiguous()
        prev = torch.cat([scores, scores.new_zeros(scores.shape[0], 1)], 1)
        sd = oracle.pgt_top1(prev, boxes, off, torch.cat(gts), goff, img)
        a = oracle.refine_assign(boxes, off, sd["seed_boxes"], sd["seed_classes"], sd["seed_scores"], sd["seed_weights"],
                                 goff, sd["seed_count"], K, 0.5)
        cls = a["gt_classes"].clone()
        torch.manual_seed(11)                                                     # subsample_labels, image by image
        for n in range(len(props)):
            sl = slice(off[n], off[n + 1])
            fg, bg = d2_shim.subsample_labels(cls[sl], batch, 1.0, K)
            keep = torch.cat([fg, bg])
            s = torch.full_like(cls[sl], -1)
            s[keep] = cls[sl][keep]
            cls[sl] = s
        r = heads.box_refinery[0]
        logits, _ = oracle.align(r.cls.projection(x), r.cls.class_weight.t().contiguous(), 50.0, 2, True, want_probs=False)
        lc, lb = oracle.refine_losses(logits, r.bbox_pred(x), cls, a["gt_weights"], boxes, a["gt_boxes"], K)
    assert abs(float(losses["loss_cls_object_mining"]) - float(loss_mil)) <= 1e-5 * float(loss_mil)
    assert abs(float(losses["loss_cls_r0"]) - float(lc)) <= 1e-5 * abs(float(lc)) + 1e-7
    assert abs(float(losses["loss_box_reg_r0"]) - float(lb)) <= 1e-5 * abs(float(lb)) + 1e-8
    assert R <= batch or int((cls != -1).sum()) == batch * len(props)            # the sampling branch really ran

    # ---- inference: the reference's probabilities through the oracle's detection tail -----------------------
    inst = H.run_test(heads, feats, props, text)
    with torch.no_grad():
        pred = [r(x, text, True) for r in heads.box_refinery]
        probs = torch.cat(heads.box_refinery[-1].predict_probs_K(pred, props))
        bx = torch.cat(heads.box_refinery[-1].predict_boxes_K(pred, props))
    det = oracle.detections(probs, bx, off, torch.tensor([list(p.image_size) for p in props], dtype=torch.float32),
                            1e-5, 0.3, 100, oracle.IOU_TV_CPU)
    for n, i in enumerate(inst):
        c = int(det["det_count"][n])
        assert c == len(i.scores)
        assert torch.equal(det["det_boxes"][n, :c], i.pred_boxes.tensor) and torch.equal(det["det_scores"][n, :c], i.scores)
        assert torch.equal(det["det_classes"][n, :c], i.pred_classes) and torch.equal(det["det_rows"][n, :c], i.pred_inds)


def test_reference_arm_of_the_bench_is_the_reference_code():
    """bench.py --impl reference / cpu_baseline: oracle/ref_path.py runs the reference's own classes; same numbers as
    the port (oracle/cpu_path.py) and as the oracle's kernels"""
    from oracle import cpu_path, ref_path
    from wsovod_b200 import synth
    w = synth.workload("c1")
    impl, kind = ref_path.get()
    assert kind == "reference" and impl is ref_path
    _, n, pooled, probs, inst = impl.run_slice(w, images=1, proposals=400)
    _, n2, pooled2, probs2, dets2 = cpu_path.run_slice(w, images=1, proposals=400)
    assert n == n2 == 400 and torch.equal(pooled, pooled2) and torch.equal(probs, probs2)
    assert torch.equal(inst[0].scores, dets2[0][1]) and torch.equal(inst[0].pred_inds, dets2[0][3])
    o, _ = oracle.roi_pool(w["features"], w["rois"][:400], w["spatial_scale"], 7)
    assert torch.equal(pooled, o * (w["objectness"][:400] + 1).view(-1, 1, 1, 1))
