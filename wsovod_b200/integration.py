"""The reference-side binding of INTEGRATION.md section 3, as code: ``patch_reference`` rebinds the symbols of the
IMPORTED reference modules (wsovod.modeling.roi_heads.roi_heads, .fast_rcnn_open_vocabulary, wsovod.modeling.poolers)
to this package's kernels, so the reference's own ``WSOVODROIHeads.forward / _forward_box`` (roi_heads.py:648-907)
runs unmodified on top of them.  ``tests/test_gpu_dropin.py`` executes exactly this against the stock reference.

    import wsovod.modeling.roi_heads.roi_heads as rh
    import wsovod.modeling.roi_heads.fast_rcnn_open_vocabulary as fr
    import wsovod.modeling.poolers as poolers
    undo = wsovod_b200.integration.patch_reference(rh, fr, poolers)      # before build_model(cfg)

What is rebound (reference file:line -> replacement):
  poolers.ROIPooler, rh.ROIPooler (poolers.py:119)                       -> modeling.ROIPooler
  rh / fr.OpenVocabularyClassifier (open_vocabulary_classifier.py:14)    -> modeling.OpenVocabularyClassifier
  fr.ObjectMiningOutputLayers.forward (:318-367)                         -> cls/det Linears + ops.mil
  fr.ObjectMiningOutputLayers.predict_probs_img (:604-618)               -> the image scores of the same kernel pass
  fr.fast_rcnn_inference (:52-96, called by .inference :894-924)         -> modeling.fast_rcnn_inference
  fr.InstanceRefinementOutputLayers.losses (:754-810)                    -> ops.refine_losses for the shipped loss flavour
  rh.WSOVODROIHeads.get_pgt_top_k (:1043-1343), top_k=1 / no SAM         -> ops.pgt_top1
  rh.WSOVODROIHeads.label_and_sample_proposals_wsl (:1722-1825)          -> ops.refine_assign (+ torch RNG sampling)
Calls whose arguments fall outside what the kernels implement (SAM tightening, top_k != 1, other loss types, the
ContextLocNet three-way miner stays on ops.mil too) go to the original reference code: the patch never changes results.
"""
import torch

from . import ops
from .modeling import (OpenVocabularyClassifier, ROIPooler, fast_rcnn_inference, get_pgt_top_k, label_proposals_wsl)
from .modeling.roi_heads import _offsets_tensor


def _offsets(proposals, device):
    off = [0]
    for p in proposals:
        off.append(off[-1] + len(p))
    return _offsets_tensor(off, device)


def _miner_forward(self, x, proposals=None, context=False):
    """fast_rcnn_open_vocabulary.py:318-367 with the per-image softmax loop (:343-354) as one kernel"""
    if context:
        C, D = self.forward_contextlocnet(x)
    else:
        if x.dim() > 2:
            x = torch.flatten(x, start_dim=1)
        C, D = self.cls(x), self.det(x)
    if self.num_classes == 1:                                                  # :338-340
        C = torch.cat((C, torch.zeros_like(C)), dim=1)
        D = torch.cat((D, torch.zeros_like(D)), dim=1)
    off = _offsets_tensor([0, C.shape[0]], C.device) if proposals is None else _offsets(proposals, C.device)
    scores, img = ops.mil(C, D, off)
    if self.num_classes == 1:                                                  # :356-357
        scores, _ = torch.split(scores, 1, dim=1)
        img = None
    self._wsovod_b200_last = (scores, img)
    deltas = torch.zeros(scores.shape[0], self.num_bbox_reg_classes * self.box_dim, dtype=scores.dtype,
                         device=scores.device, requires_grad=False)
    return scores, deltas


def _make_predict_probs_img(original):
    def predict_probs_img(self, predictions, proposals):
        last = getattr(self, "_wsovod_b200_last", None)
        if last is not None and last[0] is predictions[0] and last[1] is not None:
            return last[1]
        return original(self, predictions, proposals)
    return predict_probs_img


def _make_losses(original):
    def losses(self, predictions, proposals, num_classes=None):
        reg = self.refine_reg[self.refine_k]
        plain = (getattr(self, "cross_entropy_weighted", False) and len(proposals) and predictions[0].is_cuda
                 and (not reg or self.box_reg_loss_type == "smooth_l1_weighted"))
        if not plain:
            return original(self, predictions, proposals, num_classes)
        scores, deltas = predictions
        gt_classes = torch.cat([p.gt_classes for p in proposals], dim=0)
        gt_weights = torch.cat([p.gt_weights for p in proposals], dim=0)
        K = num_classes if num_classes else self.num_classes
        k = str(self.refine_k)
        if reg:
            pb = torch.cat([p.proposal_boxes.tensor for p in proposals], dim=0)
            gb = torch.cat([(p.gt_boxes if p.has("gt_boxes") else p.proposal_boxes).tensor for p in proposals], dim=0)
            lc, lb = ops.refine_losses(scores, deltas, gt_classes, gt_weights, pb, gb, K, self.box2box_transform.weights,
                                       self.smooth_l1_beta)
            out = {"loss_cls_r" + k: lc, "loss_box_reg_r" + k: lb}
        else:
            lc, _ = ops.refine_losses(scores, None, gt_classes, gt_weights, num_classes=K)
            out = {"loss_cls_r" + k: lc}
        return {n: v * self.loss_weight.get(n, 1.0) for n, v in out.items()}
    return losses


def _make_get_pgt_top_k(original):
    def get_pgt(self, prev_pred_boxes, prev_pred_scores, proposals, top_k=1, thres=0, need_instance=True,
                need_weight=True, sam=None, file_names=None):
        plain = (top_k == 1 and thres == 0 and need_instance and need_weight and sam is None
                 and prev_pred_boxes[0].dim() == 2 and prev_pred_boxes[0].size(1) == 4 and prev_pred_boxes[0].is_cuda)
        if not plain:
            return original(self, prev_pred_boxes, prev_pred_scores, proposals, top_k=top_k, thres=thres,
                            need_instance=need_instance, need_weight=need_weight, sam=sam, file_names=file_names)
        targets, seeds = get_pgt_top_k(prev_pred_boxes, prev_pred_scores, proposals, self.gt_classes_img_int,
                                       self.pred_class_img_logits, self.num_classes)
        self._wsovod_b200_seeds = (targets, seeds)
        return targets
    return get_pgt


def _make_label(original):
    def label(self, k, proposals, targets, suffix=""):
        cached = getattr(self, "_wsovod_b200_seeds", None)
        if self.proposal_append_gt or self.cls_agnostic_bbox_known or cached is None or cached[0] is not targets:
            return original(self, k, proposals, targets, suffix=suffix)
        m = self.proposal_matchers[k]
        thr = [t for t in m.thresholds if t not in (float("inf"), -float("inf"))]
        if len(thr) != 1 or list(m.labels) != [0, 1]:
            return original(self, k, proposals, targets, suffix=suffix)
        out, _ = label_proposals_wsl(proposals, cached[1], self.num_classes, thr[0], self.batch_size_per_images[k],
                                     self.positive_sample_fractions[k])
        return out
    return label


def patch_reference(rh=None, fr=None, poolers=None, precision=ops.ALIGN_TF32):
    """Rebind the imported reference modules (any of them may be None).  ``precision`` is the contraction type the
    substituted OpenVocabularyClassifier uses (ALIGN_FP32 for 1e-5 parity, ALIGN_TF32 within the stated 5e-2 logit
    tolerance).  Returns a callable that restores every binding."""
    saved = []

    def rebind(obj, name, value):
        saved.append((obj, name, getattr(obj, name)))
        setattr(obj, name, value)

    class _Classifier(OpenVocabularyClassifier):
        def __init__(self, *a, **kw):
            kw.setdefault("precision", precision)
            super().__init__(*a, **kw)

    for mod in (rh, poolers):
        if mod is not None and hasattr(mod, "ROIPooler"):
            rebind(mod, "ROIPooler", ROIPooler)
    for mod in (rh, fr):
        if mod is not None and hasattr(mod, "OpenVocabularyClassifier"):
            rebind(mod, "OpenVocabularyClassifier", _Classifier)
    if fr is not None:
        rebind(fr.ObjectMiningOutputLayers, "forward", _miner_forward)
        rebind(fr.ObjectMiningOutputLayers, "predict_probs_img",
               _make_predict_probs_img(fr.ObjectMiningOutputLayers.predict_probs_img))
        rebind(fr, "fast_rcnn_inference", fast_rcnn_inference)
        rebind(fr.InstanceRefinementOutputLayers, "losses", _make_losses(fr.InstanceRefinementOutputLayers.losses))
    if rh is not None:
        for cls_name in ("WSOVODROIHeads", "WSOVODMixedDatasetsROIHeads"):
            c = getattr(rh, cls_name, None)
            if c is not None:
                rebind(c, "get_pgt_top_k", _make_get_pgt_top_k(c.get_pgt_top_k))
                rebind(c, "label_and_sample_proposals_wsl", _make_label(c.label_and_sample_proposals_wsl))

    def undo():
        while saved:
            obj, name, value = saved.pop()
            setattr(obj, name, value)
    return undo
