"""``WSOVODROIHeads`` / ``WSOVODMixedDatasetsROIHeads``: the reference's ROI-head surface
(wsovod/modeling/roi_heads/roi_heads.py:430-907, :1860-2380) over the batched kernels of this package.

Same constructor keywords, same ``forward(images, features, proposals, data_aware_features, targets, classifier,
append_background, file_names, loaded_proposals)`` and return structure (train ``(proposals, losses)``, test
``(pred_instances, {}, all_scores, all_boxes)``, :648-694).  ``_forward_box`` follows :696-907 step for step; every
per-image Python loop of the reference is one kernel call here:

  pool (+ objectness scale, :727-739)  ->  box_head [PyTorch, out of scope]  ->  object miner (MIL, :761)
  train: MIL loss (:764) -> image-level scores (:766) -> per refinement stage: seeds (get_pgt_top_k / get_pgt_mist,
         :786-806) -> pseudo-label assignment (:808-813) -> alignment logits (:815) -> weighted losses (:817)
  test:  alignment logits + softmax of every stage -> mean -> per-class NMS + top-k (:893-899)

What stays the reference's PyTorch (north star): ``box_head``, the ``cls/det/bbox_pred`` Linears, the projection
MLP, the MIL BCE.  Not mirrored: visualisation hooks (``_vis_*``), SAM box tightening (``sam`` must be None),
``train_on_pred_boxes``, the logging-only ``label_and_sample_proposals`` at head entry (:670, its result is
overwritten per stage).
"""
from typing import Dict, List, Optional

import torch
from torch import nn

from .. import ops
from ..structures import Boxes, Instances
from .roi_heads import get_image_level_gt, get_pgt_top_k, label_proposals_wsl, pgt_candidates


@torch.no_grad()
def get_pgt_mist(prev_pred_boxes, prev_pred_scores, proposals, gt_classes_img_int, pred_class_img_logits, num_classes,
                 top_pro=0.15, thres=0.05, nms_thresh=0.2, iou_mode=ops.IOU_TV_CUDA, nms_fn=None):
    """roi_heads.py:910-1040 without SAM (the MIST seeds): per image and image-level class the top
    ``max(int(n * top_pro), 1)`` proposals with box area > 20 (:1090-1118), rank 0 always and the rest only with
    score >= thres (:1148-1175), then ONE class-agnostic NMS at 0.2 over the image's candidates (:930-939) -- all
    images in one ``batched_nms`` call (group = image).  Returns (targets, flat seeds) like ``get_pgt_top_k``;
    ``gt_weights`` are the seed scores, as the reference's no-SAM branch zips them (:1035-1037)."""
    dev = prev_pred_boxes[0].device
    cs, cb, cc, _ = pgt_candidates(prev_pred_boxes, prev_pred_scores, proposals, gt_classes_img_int, pred_class_img_logits,
                                   top_pro, thres)
    cg = [torch.full((c.numel(),), n, dtype=torch.int64, device=dev) for n, c in enumerate(cs)]
    B, S, C, Gr = torch.cat(cb), torch.cat(cs), torch.cat(cc), torch.cat(cg)
    if nms_fn is None:
        keep, num_keep = torch.ops.wsovod_b200.batched_nms(B, S, Gr, len(proposals), float(nms_thresh), int(iou_mode))
        keep = keep[: int(num_keep.item())]
    else:
        keep = nms_fn(B, S, Gr, nms_thresh)
    # per image, score-descending (batched_nms returns the kept indices of all groups sorted by score)
    kg = Gr[keep]
    order = torch.sort(kg, stable=True).indices
    keep = keep[order]
    counts = torch.bincount(kg, minlength=len(proposals)).tolist()
    targets, off = [], 0
    for n, p in enumerate(proposals):
        k = keep[off:off + counts[n]]
        off += counts[n]
        targets.append(Instances(p.image_size, gt_boxes=Boxes(B[k]), gt_classes=C[k], gt_scores=S[k], gt_weights=S[k]))
    goff = [0]
    for c in counts:
        goff.append(goff[-1] + c)
    seeds = dict(seed_boxes=torch.cat([t.gt_boxes.tensor for t in targets]), seed_classes=torch.cat([t.gt_classes for t in targets]),
                 seed_scores=torch.cat([t.gt_scores for t in targets]), seed_weights=torch.cat([t.gt_weights for t in targets]),
                 seed_offsets=torch.tensor(goff, dtype=torch.int64, device=dev), seed_count=None)
    return targets, seeds


class WSOVODROIHeads(nn.Module):
    def __init__(self, *, num_classes: int, box_in_features: List[str], box_pooler: nn.Module, box_head: nn.Module,
                 object_miner: nn.Module, box_refinery: List[nn.Module], sam=None, train_on_pred_boxes: bool = False,
                 mrrp_on: bool = False, mrrp_num_branch: int = 3, mrrp_fast: bool = False, refine_K: Optional[int] = None,
                 refine_mist: bool = False, refine_reg: Optional[List[bool]] = None, sampling_on: bool = False,
                 iou_thresholds: Optional[List[float]] = None, batch_size_per_images: Optional[List[int]] = None,
                 positive_sample_fractions: Optional[List[float]] = None, cls_agnostic_bbox_known: bool = False,
                 pooler_type: str = "ROIPool", rpn_on: bool = False, metadata: Optional[Dict] = None,
                 output_dir: Optional[str] = None, vis_test: bool = False, vis_period: int = 0, **kwargs):
        super().__init__()
        if sam is not None:
            raise NotImplementedError("SAM box tightening (WSOVOD.BBOX_REFINE) is outside the accelerated path")
        if train_on_pred_boxes:
            raise NotImplementedError("train_on_pred_boxes is not used by the shipped WSOVOD configs")
        self.num_classes = num_classes
        self.in_features = self.box_in_features = box_in_features
        self.box_pooler, self.box_head, self.object_miner = box_pooler, box_head, object_miner
        self.refine_K = len(box_refinery) if refine_K is None else refine_K
        self.box_refinery = list(box_refinery)
        for k in range(self.refine_K):                                       # same module names as :525-526
            self.add_module("box_refinery_{}".format(k), self.box_refinery[k])
        K = max(self.refine_K, 1)
        self.refine_mist = refine_mist
        self.refine_reg = list(refine_reg) if refine_reg is not None else [False] * K
        self.sampling_on = sampling_on
        self.iou_thresholds = list(iou_thresholds) if iou_thresholds is not None else [0.5] * K
        self.batch_size_per_images = list(batch_size_per_images) if batch_size_per_images is not None else [4096] * K
        self.positive_sample_fractions = (list(positive_sample_fractions) if positive_sample_fractions is not None
                                          else [1.0] * K)
        self.mrrp_on, self.mrrp_num_branch, self.mrrp_fast = mrrp_on, mrrp_num_branch, mrrp_fast
        self.cls_agnostic_bbox_known = cls_agnostic_bbox_known
        self.pooler_type, self.rpn_on, self.metadata = pooler_type, rpn_on, metadata
        self.output_dir, self.vis_test, self.vis_period = output_dir, vis_test, vis_period
        self.iter = self.iter_test = self.epoch_test = 0
        self.proposal_targets = None

    # ------------------------------------------------------------------------------------------ :648-694
    def forward(self, images, features, proposals, data_aware_features=None, targets=None, classifier=None,
                append_background=True, file_names=None, loaded_proposals=None):
        self.gt_classes_img, self.gt_classes_img_int, self.gt_classes_img_oh = get_image_level_gt(targets, self.num_classes)
        self.images = images
        if self.training:
            assert targets, "'targets' argument is required during training"
            del targets
            losses = self._forward_box(features, proposals, data_aware_features, classifier, append_background,
                                       file_names=file_names, loaded_proposals=loaded_proposals)
            self.iter += 1
            if self.iter_test > 0:
                self.epoch_test += 1
            self.iter_test = 0
            return proposals, losses
        pred_instances, all_scores, all_boxes = self._forward_box(features, proposals, data_aware_features, classifier,
                                                                  append_background)
        self.iter_test += 1
        return pred_instances, {}, all_scores, all_boxes

    # ------------------------------------------------------------------------------------------ :696-907
    def _forward_box(self, features, proposals, data_aware_features=None, classifier=None, append_background=True,
                     file_names=None, loaded_proposals=None):
        features = [features[f] for f in self.box_in_features]
        if self.mrrp_on:
            features = [ff for f in features for ff in torch.chunk(f, self.mrrp_num_branch)]
        # pooling with `* (objectness_logits + 1)` folded into the store (:727-739)
        box_features = self.box_pooler(
            features, [x.proposal_boxes for x in proposals],
            level_ids=[torch.div(x.level_ids, 1000, rounding_mode="floor") for x in proposals] if self.mrrp_on else None,
            objectness_logits=[x.objectness_logits for x in proposals])
        box_features = self.box_head(box_features)
        if self.pooler_type == "ROILoopPool":
            box_features, frame, context = torch.chunk(box_features, 3, dim=0)
            if data_aware_features is not None:
                box_features, frame, context = (t + data_aware_features for t in (box_features, frame, context))
            predictions = self.object_miner([box_features, frame, context], proposals, context=True)
            del frame, context
        else:
            if data_aware_features is not None:
                box_features = box_features + data_aware_features
            predictions = self.object_miner(box_features, proposals)

        if not self.training:
            if self.refine_K > 0:
                predictions_K = [self.box_refinery[k](box_features, classifier, append_background)
                                 for k in range(self.refine_K)]
                pred_instances, _, all_scores, all_boxes = self.box_refinery[-1].inference(predictions_K, proposals)
            else:
                raise NotImplementedError("refine_K == 0 at test time calls the undefined self.box_predictor upstream (:904)")
            return pred_instances, all_scores, all_boxes

        losses = self.object_miner.losses(predictions, proposals, self.gt_classes_img_oh)
        self.pred_class_img_logits = self.object_miner.predict_probs_img(predictions, proposals).clone().detach()
        prev_pred_scores = [s.detach() for s in self.object_miner.predict_probs(predictions, proposals)]
        prev_pred_boxes = self.object_miner.predict_boxes(predictions, proposals)
        for k in range(self.refine_K):
            if self.refine_mist:
                targets, seeds = get_pgt_mist(prev_pred_boxes, prev_pred_scores, proposals, self.gt_classes_img_int,
                                              self.pred_class_img_logits, self.num_classes)
            else:
                targets, seeds = get_pgt_top_k(prev_pred_boxes, prev_pred_scores, proposals, self.gt_classes_img_int,
                                               self.pred_class_img_logits, self.num_classes, build_targets=False)
            if not self.sampling_on:
                raise NotImplementedError("WSOVOD.SAMPLING.SAMPLING_ON False (detectron2's GT-style sampling, :812-813) is "
                                          "not used by the shipped configs")
            proposals_k, _ = label_proposals_wsl(proposals, seeds, self.num_classes, self.iou_thresholds[k],
                                                 self.batch_size_per_images[k], self.positive_sample_fractions[k])
            predictions_k = self.box_refinery[k](box_features, classifier=classifier, append_background=append_background)
            losses.update(self.box_refinery[k].losses(predictions_k, proposals_k, num_classes=self.num_classes, refine_k=k))
            if k + 1 < self.refine_K or self.rpn_on:       # the next stage's (or the RPN's) seeds come from this stage (:822-828)
                prev_pred_scores = [s.detach() for s in self.box_refinery[k].predict_probs(predictions_k, proposals_k)]
                prev_pred_boxes = [b.detach() for b in self.box_refinery[k].predict_boxes(predictions_k, proposals_k)]
        if self.rpn_on:                                                       # pseudo targets of the RPN (:872-881)
            self.proposal_targets, _ = get_pgt_top_k(prev_pred_boxes, prev_pred_scores, proposals, self.gt_classes_img_int,
                                                     self.pred_class_img_logits, self.num_classes)
        return losses


class WSOVODMixedDatasetsROIHeads(WSOVODROIHeads):
    """roi_heads.py:1860-2380: one MIL head per source dataset (``object_miners[source_id]``, class counts
    ``num_classes_list[source_id]``), shared refinement heads driven by the batch's text embeddings
    (``classifier``).  ``forward`` takes the extra ``source_id`` (:2100-2112), selects the miner / class count of
    the step (:2117-2122, :2223-2242) and runs the shared path; the refinement losses receive the step's class
    count (:2286)."""

    def __init__(self, *, object_miners: List[nn.Module], num_classes_list: List[int], **kwargs):
        kwargs.setdefault("object_miner", object_miners[0])
        kwargs.setdefault("num_classes", num_classes_list[0])
        super().__init__(**kwargs)
        self.object_miners = nn.ModuleList(object_miners)
        self.num_classes_list = list(num_classes_list)

    def forward(self, images, features, proposals, data_aware_features=None, targets=None, classifier=None, source_id=0,
                append_background=True, file_names=None, loaded_proposals=None):
        self.object_miner = self.object_miners[source_id]
        if self.training:
            self.num_classes = self.num_classes_list[source_id]
        return super().forward(images, features, proposals, data_aware_features, targets, classifier, append_background,
                               file_names, loaded_proposals)
