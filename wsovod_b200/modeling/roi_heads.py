"""Host-side mirror of the reference's output-layer / ROI-head methods on the region-scoring path.

Same names, argument meaning and return structure as
wsovod/modeling/roi_heads/fast_rcnn_open_vocabulary.py (``ObjectMiningOutputLayers``,
``InstanceRefinementOutputLayers``, ``fast_rcnn_inference``) and the pseudo-label methods of
``WSOVODROIHeads`` in wsovod/modeling/roi_heads/roi_heads.py (``get_image_level_gt``, ``get_pgt_top_k``,
``label_and_sample_proposals_wsl``) -- but each per-image Python loop of the reference is ONE batched
kernel call here.  The Linear layers (cls/det/bbox_pred, projection) stay PyTorch (out of scope).
Proposals are duck-typed: anything with ``len()``, ``.proposal_boxes.tensor`` and ``.image_size``.
"""
import math
from typing import List, Tuple

import torch
from torch import nn

from .. import ops
from ..structures import Boxes, Instances


def apply_deltas(deltas, boxes, weights=(10.0, 10.0, 5.0, 5.0), scale_clamp=math.log(1000.0 / 16)):
    """detectron2 Box2BoxTransform.apply_deltas (call sites fast_rcnn_open_vocabulary.py:987-1017); box
    decoding is outside the accelerated path and stays plain PyTorch, op for op."""
    deltas = deltas.float()
    boxes = boxes.to(deltas.dtype)
    widths = boxes[:, 2] - boxes[:, 0]
    heights = boxes[:, 3] - boxes[:, 1]
    ctr_x = boxes[:, 0] + 0.5 * widths
    ctr_y = boxes[:, 1] + 0.5 * heights
    wx, wy, ww, wh = weights
    dx, dy = deltas[:, 0::4] / wx, deltas[:, 1::4] / wy
    dw = torch.clamp(deltas[:, 2::4] / ww, max=scale_clamp)
    dh = torch.clamp(deltas[:, 3::4] / wh, max=scale_clamp)
    pcx = dx * widths[:, None] + ctr_x[:, None]
    pcy = dy * heights[:, None] + ctr_y[:, None]
    pw = torch.exp(dw) * widths[:, None]
    ph = torch.exp(dh) * heights[:, None]
    out = torch.stack((pcx - 0.5 * pw, pcy - 0.5 * ph, pcx + 0.5 * pw, pcy + 0.5 * ph), dim=-1)
    return out.reshape(deltas.shape)


_OFFSET_CACHE = {}


def _offsets_tensor(off, device):
    """int64 device tensor of a host offsets list.  A pageable host->device copy synchronises the stream, and the
    same per-image counts recur step after step (4000 proposals / image), so the tensors are cached."""
    key = (tuple(off), str(device))
    t = _OFFSET_CACHE.get(key)
    if t is None:
        if len(_OFFSET_CACHE) >= 64:
            _OFFSET_CACHE.clear()
        t = _OFFSET_CACHE[key] = torch.tensor(off, dtype=torch.int64, device=device)
    return t


def _sizes_tensor(image_shapes, device):
    key = (tuple((float(h), float(w)) for h, w in image_shapes), str(device), "hw")
    t = _OFFSET_CACHE.get(key)
    if t is None:
        if len(_OFFSET_CACHE) >= 64:
            _OFFSET_CACHE.clear()
        t = _OFFSET_CACHE[key] = torch.tensor([list(k) for k in key[0]], dtype=torch.float32, device=device).reshape(-1, 2)
    return t


def _offsets(proposals, device):
    sizes = [len(p) for p in proposals]
    off = [0]
    for s in sizes:
        off.append(off[-1] + s)
    return _offsets_tensor(off, device), sizes


# ------------------------------------------------------------------------------------------------
class ObjectMiningOutputLayers(nn.Module):
    """MIL head (fast_rcnn_open_vocabulary.py:220-618).  ``cls``/``det`` are Linear layers unless a
    ``class_head`` (an OpenVocabularyClassifier) is supplied -- the fused "alignment + MIL" variant of
    the commented-out line roi_heads.py:588-589."""

    def __init__(self, input_size, num_classes, class_head=None, loss_weight=None, mean_loss=True):
        super().__init__()
        self.num_classes = num_classes
        self.loss_weight = loss_weight or {}
        self.mean_loss = mean_loss
        self.det = nn.Linear(input_size, num_classes)
        nn.init.xavier_uniform_(self.det.weight)
        nn.init.constant_(self.det.bias, 0)
        if class_head is None:
            self.cls = nn.Linear(input_size, num_classes)
            nn.init.xavier_uniform_(self.cls.weight)
            nn.init.constant_(self.cls.bias, 0)
        else:
            self.cls = class_head

    def forward(self, x, proposals=None, context=False):
        """-> (scores (M,K), proposal_deltas (M,4) zeros); scores = softmax(C,1) * per-image softmax(D,0)"""
        if context:                                           # ContextLocNet variant (:369-390)
            xr, xf, xc = x[:]
            C = self.cls(torch.flatten(xr, 1))
            D = self.det(torch.flatten(xf, 1)) - self.det(torch.flatten(xc, 1))
        else:
            if x.dim() > 2:
                x = torch.flatten(x, start_dim=1)
            fused = self._fused(x, proposals)
            if fused is not None:
                self._last = fused
                return fused[0], torch.zeros(fused[0].shape[0], 4, dtype=fused[0].dtype, device=fused[0].device)
            C, D = self.cls(x), self.det(x)
        scores, img = self.score(C, D, proposals)
        self._last = (scores, img)          # the image-level scores come out of the same kernel pass
        deltas = torch.zeros(scores.shape[0], 4, dtype=scores.dtype, device=scores.device)
        return scores, deltas

    def _fused(self, x, proposals):
        """the class-head variant (roi_heads.py:588-590) through the fused alignment + MIL kernel: `cls` is an
        OpenVocabularyClassifier on its tensor-core path with at most 256 concepts and stored weights"""
        from .class_heads import OpenVocabularyClassifier
        c = self.cls
        if not (isinstance(c, OpenVocabularyClassifier) and c.precision == ops.ALIGN_TF32 and x.is_cuda
                and 1 < self.num_classes <= 256):
            return None
        off = _offsets_tensor([0, x.shape[0]], x.device) if proposals is None else _offsets(proposals, x.device)[0]
        scores, img, _ = ops.align_mil(c.projection(x), c._stored_kd(), self.det(x), off, c.norm_temperature,
                                       2 if c.norm_weight else 0, c.cls_bias if c.use_bias else None)
        return scores, img

    @staticmethod
    def score(C, D, proposals=None):
        if proposals is None:
            off = _offsets_tensor([0, C.shape[0]], C.device)
        else:
            off, _ = _offsets(proposals, C.device)
        return ops.mil(C, D, off)                             # (scores, clamped image-level scores)

    def losses(self, predictions, proposals, gt_classes_img_oh):
        """:392-427 -- image-level BCE of the MIL head (plain PyTorch: N x K numbers; the image-level scores come
        out of the MIL kernel with their autograd formula): {"loss_cls_object_mining": ...}"""
        img = self.predict_probs_img(predictions, proposals)
        assert gt_classes_img_oh.dim() == 2 and img.dim() == 2
        if self.mean_loss:
            loss = nn.functional.binary_cross_entropy(img.float(), gt_classes_img_oh.float(), reduction="mean")
        else:
            loss = nn.functional.binary_cross_entropy(img.float(), gt_classes_img_oh.float(),
                                                      reduction="sum") / (1.0 * gt_classes_img_oh.size(0))
        return {"loss_cls_object_mining": loss * self.loss_weight.get("loss_cls_object_mining", 1.0)}

    def predict_probs_img(self, predictions, proposals):
        """:604-618 -- clamp(sum over the image's proposals of scores, 1e-6, 1-1e-6)"""
        scores, _ = predictions
        last = getattr(self, "_last", None)
        if last is not None and last[0] is scores:
            return last[1]
        sizes = [len(p) for p in proposals]                   # scores edited by the caller: plain reduction
        sums = torch.cat([s.sum(dim=0, keepdim=True) for s in scores.split(sizes, dim=0)], dim=0)
        return torch.clamp(sums, min=1e-6, max=1.0 - 1e-6)

    def predict_probs(self, predictions, proposals):
        """:580-602 -- scores with a zero background column, split per image"""
        scores, _ = predictions
        probs = torch.cat((scores, scores.new_zeros(scores.shape[0], 1)), 1)
        return probs.split([len(p) for p in proposals], dim=0)

    def predict_boxes(self, predictions, proposals):
        return [p.proposal_boxes.tensor for p in proposals]    # :567 (early return of the reference)


class InstanceRefinementOutputLayers(nn.Module):
    """Refinement head (fast_rcnn_open_vocabulary.py:621-1058): class_head logits (+ optional bbox_pred)."""

    def __init__(self, input_size, num_classes, class_head, test_score_thresh=0.0, test_nms_thresh=0.5,
                 test_topk_per_image=100, refine_reg=False, box_dim=4, iou_mode=ops.IOU_TV_CUDA,
                 bbox_reg_weights=(10.0, 10.0, 5.0, 5.0)):
        super().__init__()
        self.num_classes = num_classes
        self.cls = class_head
        self.refine_reg = refine_reg
        if refine_reg:
            self.bbox_pred = nn.Linear(input_size, box_dim)
            nn.init.normal_(self.bbox_pred.weight, std=0.001)
            nn.init.constant_(self.bbox_pred.bias, 0)
        self.test_score_thresh, self.test_nms_thresh = test_score_thresh, test_nms_thresh
        self.test_topk_per_image = test_topk_per_image
        self.iou_mode = iou_mode
        self.bbox_reg_weights = bbox_reg_weights

    def forward(self, x, classifier=None, append_background=True):
        if x.dim() > 2:
            x = torch.flatten(x, start_dim=1)
        scores = self.cls(x, classifier, append_background=append_background)
        deltas = self.bbox_pred(x) if self.refine_reg else torch.zeros(scores.shape[0], 4, dtype=scores.dtype,
                                                                       device=scores.device)
        return scores, deltas

    def predict_probs(self, predictions, proposals):
        scores, _ = predictions
        return torch.softmax(scores, dim=-1).split([len(p) for p in proposals], dim=0)    # :1034-1036

    def predict_probs_K(self, predictions, proposals):
        probs = torch.zeros_like(predictions[0][0])
        for s, _ in predictions:
            probs += torch.softmax(s, dim=-1)
        return (probs / len(predictions)).split([len(p) for p in proposals], dim=0)       # :1052-1058

    def predict_boxes(self, predictions, proposals):
        """:964-985 -- class-agnostic boxes: apply_deltas(proposal_deltas, proposal_boxes)"""
        if not len(proposals):
            return []
        _, deltas = predictions
        boxes = torch.cat([p.proposal_boxes.tensor for p in proposals], dim=0)
        return apply_deltas(deltas, boxes, self.bbox_reg_weights).split([len(p) for p in proposals])

    def predict_boxes_K(self, predictions, proposals):
        """:987-1017 -- mean of the heads' deltas, then apply_deltas"""
        if not len(proposals):
            return []
        deltas = torch.zeros_like(predictions[0][1])
        for _, d in predictions:
            deltas += d
        deltas = deltas / len(predictions)
        boxes = torch.cat([p.proposal_boxes.tensor for p in proposals], dim=0)
        return apply_deltas(deltas, boxes, self.bbox_reg_weights).split([len(p) for p in proposals])

    def losses(self, predictions, proposals, num_classes=None, refine_k=0, smooth_l1_beta=0.0, loss_weight=None):
        """:754-810 with cross_entropy_weighted and BBOX_REG_LOSS_TYPE "smooth_l1_weighted" (the shipped
        configs): {"loss_cls_r<k>", "loss_box_reg_r<k>"} from one fused kernel (ops.refine_losses);
        proposals carry proposal_boxes, gt_classes, gt_weights and (for the box term) gt_boxes"""
        scores, deltas = predictions
        gt_classes = torch.cat([p.gt_classes for p in proposals], dim=0)
        gt_weights = torch.cat([p.gt_weights for p in proposals], dim=0)
        out = {}
        if self.refine_reg:
            pboxes = torch.cat([p.proposal_boxes.tensor for p in proposals], dim=0)
            gboxes = torch.cat([(p.gt_boxes if p.has("gt_boxes") else p.proposal_boxes).tensor for p in proposals], dim=0)
            lc, lb = ops.refine_losses(scores, deltas, gt_classes, gt_weights, pboxes, gboxes,
                                       self.num_classes if num_classes is None else num_classes, self.bbox_reg_weights,
                                       smooth_l1_beta)
            out["loss_box_reg_r" + str(refine_k)] = lb
        else:
            lc, _ = ops.refine_losses(scores, None, gt_classes, gt_weights)
        out["loss_cls_r" + str(refine_k)] = lc
        lw = loss_weight or {}
        return {k: v * lw.get(k, 1.0) for k, v in out.items()}

    def inference(self, predictions, proposals):
        """:894-924 -- predictions: (scores, deltas) or a list of them (one per refinement head)"""
        if isinstance(predictions[0], tuple):
            scores = self.predict_probs_K(predictions, proposals)
            boxes = self.predict_boxes_K(predictions, proposals)
        else:
            scores = self.predict_probs(predictions, proposals)
            boxes = self.predict_boxes(predictions, proposals)
        shapes = [p.image_size for p in proposals]
        return fast_rcnn_inference(boxes, scores, shapes, self.test_score_thresh, self.test_nms_thresh,
                                   self.test_topk_per_image, iou_mode=self.iou_mode)


# ------------------------------------------------------------------------------------------------
def fast_rcnn_inference(boxes: List[torch.Tensor], scores: List[torch.Tensor], image_shapes: List[Tuple[int, int]],
                        score_thresh: float, nms_thresh: float, topk_per_image: int, iou_mode=ops.IOU_TV_CUDA):
    """fast_rcnn_open_vocabulary.py:52-96: -> (instances, kept_indices, all_scores, all_boxes), all images in
    three kernel launches.  One device->host read (the per-image detection counts) sizes the outputs,
    as the reference's boolean indexing does implicitly."""
    if len(boxes) == 0:
        return [], [], [], []
    if boxes[0].shape[1] != 4:
        raise NotImplementedError("class-specific box regression is not used by WSOVOD (CLS_AGNOSTIC boxes)")
    if topk_per_image < 0:
        raise NotImplementedError("topk_per_image < 0: filter, then call wsovod_b200.ops.batched_nms")
    dev = scores[0].device
    sizes = [int(s.shape[0]) for s in scores]
    off = [0]
    for s in sizes:
        off.append(off[-1] + s)
    probs = scores[0] if len(scores) == 1 else torch.cat(list(scores), 0)
    bx = boxes[0] if len(boxes) == 1 else torch.cat(list(boxes), 0)
    r = ops.detections(probs, bx, _offsets_tensor(off, dev),
                       _sizes_tensor(image_shapes, dev),
                       max(sizes), score_thresh, nms_thresh, topk_per_image, iou_mode)
    counts = r["det_count"].tolist()
    instances, kept = [], []
    for n, c in enumerate(counts):
        inst = Instances(tuple(image_shapes[n]))
        inst.pred_boxes = Boxes(r["det_boxes"][n, :c])
        inst.scores = r["det_scores"][n, :c]
        inst.pred_classes = r["det_classes"][n, :c]
        inst.pred_inds = r["det_rows"][n, :c]
        instances.append(inst)
        rows = r["det_rows"][n, :c]
        # the reference's kept index counts only rows that survive the finite filter (:178-182)
        valid = torch.isfinite(boxes[n]).all(1) & torch.isfinite(scores[n]).all(1)
        if not bool(valid.all()):
            rows = (torch.cumsum(valid.to(torch.int64), 0) - 1)[rows]
        kept.append(rows)
    return instances, kept, [s.unsqueeze(0) for s in scores], [b.unsqueeze(0) for b in boxes]


def fast_rcnn_inference_single_image(boxes, scores, image_shape, score_thresh, nms_thresh, topk_per_image,
                                     iou_mode=ops.IOU_TV_CUDA):
    """:149-217"""
    inst, kept, s, b = fast_rcnn_inference([boxes], [scores], [image_shape], score_thresh, nms_thresh,
                                           topk_per_image, iou_mode)
    return inst[0], kept[0], s[0], b[0]


# ------------------------------------------------------------------------------------------------
@torch.no_grad()
def get_image_level_gt(targets, num_classes):
    """roi_heads.py:159-174 (tiny host-side op, kept in PyTorch): sorted unique classes + one-hot"""
    if targets is None:
        return None, None, None
    gt = [torch.unique(t.gt_classes, sorted=True) for t in targets]
    gt_int = [g.to(torch.int64) for g in gt]
    oh = torch.cat([torch.zeros((1, num_classes), dtype=torch.float, device=g.device).scatter_(1, g.unsqueeze(0), 1)
                    for g in gt_int], dim=0)
    return gt, gt_int, oh


@torch.no_grad()
def pgt_candidates(prev_pred_boxes, prev_pred_scores, proposals, gt_classes_img_int, pred_class_img_logits, top_k=1, thres=0):
    """The general selection of roi_heads.py:1043-1207 (any ``top_k`` / ``thres``; the callers that leave the default
    are ``get_pgt_mist`` with top_k = 0.15, thres = 0.05 and the visualisation hooks, :919-928,1442-1449), in PyTorch:
    per image the boxes of area > 20 (:1090-1111), per image-level class the ``top_k`` best (an integer count, a
    fraction of the surviving proposals, :1114-1125), rank 0 always and the others only with score >= thres
    (:1148-1175), flattened rank-major like ``masked_select`` / ``reshape(-1)`` do; the fallback seed for an image without
    a candidate (:1181-1207).  Returns per-image lists (scores, boxes, classes, weights); weights are the image-level
    score of the class (:1140-1146)."""
    dev = prev_pred_boxes[0].device
    sizes = [len(p) for p in proposals]
    scores = prev_pred_scores.split(sizes, 0) if isinstance(prev_pred_scores, torch.Tensor) else list(prev_pred_scores)
    out_s, out_b, out_c, out_w = [], [], [], []
    for n, (b, s, gt) in enumerate(zip(prev_pred_boxes, scores, gt_classes_img_int)):
        G = gt.numel()
        if b.dim() == 2 and b.size(1) == 4:
            b = b.unsqueeze(1).expand(b.size(0), G, 4)                       # class-agnostic boxes (:1060-1064)
        else:
            b = b.reshape(b.size(0), -1, 4)[:, gt]                            # per-class boxes (:1066-1069,1087-1090)
        s = s[:, gt]
        if G > 0:
            keep = ((b[:, :, 2] - b[:, :, 0]) * (b[:, :, 3] - b[:, :, 1])) > 20      # (num, G)
            # the reference's masked_select(...).view(-1, G, 4) needs the same number of survivors in every column
            if not bool((keep == keep[:, :1]).all()):
                raise RuntimeError("shape '[-1, %d, 4]' is invalid: the area filter keeps different rows per class" % G)
            b, s = b[keep[:, 0]], s[keep[:, 0]]
        num = b.size(0)
        if G == 0:
            out_s.append(torch.ones(1, dtype=s.dtype, device=dev))
            out_b.append(torch.tensor([[-10000.0, -10000.0, 10000.0, 10000.0]], dtype=b.dtype, device=dev))
            out_c.append(torch.zeros(1, dtype=gt.dtype, device=dev))
            out_w.append(torch.ones(1, dtype=pred_class_img_logits.dtype, device=dev))
            continue
        if top_k >= 1:
            k = min(num, int(top_k))
        elif 0 < top_k < 1:
            k = max(int(num * top_k), 1)
        else:
            k = min(num, 1)
        if k > num:
            # upstream asks topk for max(int(0 * top_k), 1) = 1 row of an empty matrix (:1116-1125) and fails
            raise RuntimeError("selected index k out of range (image %d has no proposal with box area > 20)" % n)
        v, i = torch.topk(s, k, dim=0)                                        # (k, G)
        m = v.ge(thres) if thres > 0 else torch.ones_like(v, dtype=torch.bool)
        if k > 0:
            m[0] = True
        bb = torch.gather(b, 0, i.unsqueeze(2).expand(k, G, 4))
        w = pred_class_img_logits[n:n + 1, gt].expand(k, G)
        cs, cb, cc, cw = v[m], bb[m], gt.unsqueeze(0).expand(k, G)[m], w[m]
        if cs.numel() == 0:                                                   # :1181-1207
            cs = torch.ones(1, dtype=s.dtype, device=dev)
            cb = torch.tensor([[-10000.0, -10000.0, 10000.0, 10000.0]], dtype=b.dtype, device=dev)
            cc = torch.zeros(1, dtype=gt.dtype, device=dev)
            cw = torch.ones(1, dtype=w.dtype, device=dev)
        out_s.append(cs); out_b.append(cb); out_c.append(cc); out_w.append(cw)
    return out_s, out_b, out_c, out_w


@torch.no_grad()
def get_pgt_top_k(prev_pred_boxes, prev_pred_scores, proposals, gt_classes_img_int, pred_class_img_logits,
                  num_classes, build_targets=True, top_k=1, thres=0):
    """roi_heads.py:1043-1343 with need_weight=True, sam=None.  The default (top_k=1, thres=0: one seed per image-level
    class, what every ``_forward_box`` call asks for, :801,872) is one kernel for the batch; any other ``top_k`` /
    ``thres`` goes through ``pgt_candidates``.  Returns (targets: list[Instances{gt_boxes, gt_classes, gt_scores,
    gt_weights}], flat seeds).  ``build_targets=False`` returns (None, seeds): slicing the per-image Instances needs
    the seed counts on the host (one device->host read); the assignment kernel only takes the flat seeds."""
    dev = prev_pred_boxes[0].device
    if top_k != 1 or thres > 0:
        cs, cb, cc, cw = pgt_candidates(prev_pred_boxes, prev_pred_scores, proposals, gt_classes_img_int,
                                        pred_class_img_logits, top_k, thres)
        goff = [0]
        for c in cs:
            goff.append(goff[-1] + c.numel())
        seeds = dict(seed_boxes=torch.cat(cb), seed_classes=torch.cat(cc), seed_scores=torch.cat(cs), seed_weights=torch.cat(cw),
                     seed_offsets=torch.tensor(goff, dtype=torch.int64, device=dev), seed_count=None)
        if not build_targets:
            return None, seeds
        targets = [Instances(p.image_size, gt_boxes=Boxes(b), gt_classes=c, gt_scores=s, gt_weights=w)
                   for p, s, b, c, w in zip(proposals, cs, cb, cc, cw)]
        return targets, seeds
    off, sizes = _offsets(proposals, dev)
    scores = prev_pred_scores if isinstance(prev_pred_scores, torch.Tensor) else torch.cat(list(prev_pred_scores), 0)
    boxes = torch.cat([b.reshape(-1, 4) for b in prev_pred_boxes], 0)
    gsz = [int(g.numel()) for g in gt_classes_img_int]
    goff = [0]
    for s in gsz:
        goff.append(goff[-1] + s)
    goff_t = _offsets_tensor(goff, dev)
    seeds = ops.pgt_top1(scores, boxes, off, torch.cat(list(gt_classes_img_int)).to(dev), goff_t,
                         pred_class_img_logits)
    seeds["seed_offsets"] = goff_t
    if not build_targets:
        return None, seeds
    counts = seeds["seed_count"].tolist()
    targets = []
    for n, p in enumerate(proposals):
        a, c = goff[n], counts[n]
        targets.append(Instances(p.image_size, gt_boxes=Boxes(seeds["seed_boxes"][a:a + c]),
                                 gt_classes=seeds["seed_classes"][a:a + c], gt_scores=seeds["seed_scores"][a:a + c],
                                 gt_weights=seeds["seed_weights"][a:a + c]))
    return targets, seeds


@torch.no_grad()
def label_proposals_wsl(proposals, seeds, num_classes, iou_threshold=0.5, batch_size_per_image=4096,
                        positive_fraction=1.0, skip_noop_sampling=False):
    """label_and_sample_proposals_wsl + _sample_proposals_wsl (roi_heads.py:1566-1610,1722-1825): IoU ->
    argmax -> label -> class / gathered seed box, score, loss weight in one kernel for all images.  The
    random subsampling to BATCH_SIZE_PER_IMAGE stays in PyTorch and follows detectron2's subsample_labels call
    for call (two randperm draws per image, :1597-1602), so labels AND the torch RNG stream match the reference
    under the same seed.  ``skip_noop_sampling`` drops the draws for images where they cannot change a label
    (at most batch_size_per_image proposals and positive_fraction >= 1): same labels, RNG not advanced."""
    dev = seeds["seed_boxes"].device
    off, sizes = _offsets(proposals, dev)
    boxes = torch.cat([p.proposal_boxes.tensor for p in proposals], 0)
    a = ops.refine_assign(boxes, off, seeds["seed_boxes"], seeds["seed_classes"], seeds["seed_scores"],
                          seeds["seed_weights"], seeds["seed_offsets"], seeds["seed_count"], num_classes, iou_threshold)
    out = []
    o = [0]
    for n in sizes:
        o.append(o[-1] + n)
    for n, p in enumerate(proposals):
        sl = slice(o[n], o[n + 1])
        cls = a["gt_classes"][sl]
        if not (skip_noop_sampling and sizes[n] <= batch_size_per_image and positive_fraction >= 1.0):
            pos = torch.nonzero((cls != -1) & (cls != num_classes), as_tuple=True)[0]
            neg = torch.nonzero(cls == num_classes, as_tuple=True)[0]
            num_pos = min(pos.numel(), int(batch_size_per_image * positive_fraction))
            num_neg = min(neg.numel(), batch_size_per_image - num_pos)
            perm1 = torch.randperm(pos.numel(), device=dev)[:num_pos]
            perm2 = torch.randperm(neg.numel(), device=dev)[:num_neg]
            if num_pos < pos.numel() or num_neg < neg.numel():
                keep = torch.cat([pos[perm1], neg[perm2]])
                sampled = torch.full_like(cls, -1)
                sampled[keep] = cls[keep]
                cls = sampled
        q = Instances(p.image_size, proposal_boxes=p.proposal_boxes, gt_classes=cls,
                      gt_boxes=Boxes(a["gt_boxes"][sl]), gt_scores=a["gt_scores"][sl], gt_weights=a["gt_weights"][sl])
        if hasattr(p, "has") and p.has("objectness_logits"):
            q.objectness_logits = p.objectness_logits
        out.append(q)
    return out, a
