"""Generate tests/golden/*.pt by running the REFERENCE code on seeded synthetic inputs.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):

    python -m oracle.make_golden            # rewrites tests/golden/*.pt

What runs verbatim from /root/reference (through oracle/d2_shim.py):
  * wsovod/modeling/class_heads/open_vocabulary_classifier.py  OpenVocabularyClassifier.forward
    (projection replaced by nn.Identity(): the MLP is out of scope, the contraction is what we pin)
  * wsovod/modeling/roi_heads/fast_rcnn_open_vocabulary.py  ObjectMiningOutputLayers.forward /
    predict_probs_img / predict_probs, InstanceRefinementOutputLayers.predict_probs,
    fast_rcnn_inference
  * wsovod/modeling/roi_heads/roi_heads.py  WSOVODROIHeads.get_pgt_top_k,
    label_and_sample_proposals_wsl, _sample_proposals_wsl, get_image_level_gt
  * wsovod/modeling/roi_heads/fast_rcnn_open_vocabulary.py  InstanceRefinementOutputLayers.losses /
    softmax_cross_entropy_loss / box_reg_loss (weighted flavours) with autograd for the gradients
  * wsovod/data/detection_utils.py  unique_boxes, transform_proposals (the two functions are compiled from
    the file's AST: importing the module would pull detectron2.data / PIL machinery they do not use)
  * wsovod/modeling/proposal_generator/proposal_utils.py  find_top_rpn_proposals (from the AST as well)
and from the installed torchvision 0.26 (the reference's un-vendored dependency):
  * torch.ops.torchvision.roi_pool / roi_align (CPU), torchvision.ops.boxes._batched_nms_vanilla
The reference's ROILoopPool has no CPU path (ROILoopPool.h:62); its 3-way golden comes from the
compiled reference CUDA extension on the GPU box (oracle/_ref, see oracle/build_ref.py).
"""
import math
import os
import sys
import types

import torch
from torch import nn

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")
REF = os.environ.get("WSOVOD_REFERENCE", "/root/reference")


def make_rois(R, N, img_h, img_w, g, stress=True):
    """SURVEY 8d proposal generator: LogU sizes, clipped, 5% duplicates, 1% degenerate."""
    x1 = torch.rand(R, generator=g) * 0.85 * img_w
    y1 = torch.rand(R, generator=g) * 0.85 * img_h
    w = torch.exp(torch.rand(R, generator=g) * (math.log(0.6 * img_w) - math.log(16)) + math.log(16))
    h = torch.exp(torch.rand(R, generator=g) * (math.log(0.6 * img_h) - math.log(16)) + math.log(16))
    x2 = (x1 + w).clamp(max=img_w)
    y2 = (y1 + h).clamp(max=img_h)
    boxes = torch.stack([x1, y1, x2, y2], 1)
    if stress and R >= 40:
        nd = max(R // 20, 1)
        src = torch.randint(0, R, (nd,), generator=g)
        dst = torch.randint(0, R, (nd,), generator=g)
        boxes[dst] = boxes[src]
        nz = max(R // 100, 1)
        z = torch.randint(0, R, (nz,), generator=g)
        boxes[z, 2] = boxes[z, 0]
    b = torch.randint(0, N, (R,), generator=g).sort().values.float()
    return torch.cat([b[:, None], boxes], 1)


def main():
    sys.path.insert(0, os.path.dirname(HERE))
    from oracle import d2_shim
    d2_shim.install(REF)
    import wsovod.modeling.roi_heads.fast_rcnn_open_vocabulary as fr
    import wsovod.modeling.roi_heads.roi_heads as rh
    from wsovod.modeling.class_heads import OpenVocabularyClassifier
    from oracle.d2_shim import Boxes, Instances, Matcher, ShapeSpec, Box2BoxTransform

    os.makedirs(GOLD, exist_ok=True)
    g = torch.Generator().manual_seed(20261017)

    # ---- (1) pooling: torchvision CPU ops (what POOLER_TYPE "ROIPool"/"ROIAlign" run) -----------
    N, C, H, W = 2, 6, 30, 40
    feat = torch.randn(N, C, H, W, generator=g)          # negatives included (tv semantics)
    rois = make_rois(100, N, H * 8, W * 8, g)
    rois[:6, 1:] += torch.randn(6, 4, generator=g) * 150  # out-of-image / malformed boxes
    out, arg = torch.ops.torchvision.roi_pool(feat, rois, 1 / 8, 7, 7)
    rois_ok = make_rois(100, N, H * 8, W * 8, g, stress=False)
    rois_ok[:8, 1:] += 120.0                              # partly outside the map, well formed
    al = {}
    for sr in (0, 2):
        for aligned in (False, True):
            al[(sr, aligned)] = torch.ops.torchvision.roi_align(feat, rois_ok, 1 / 8, 7, 7, sr, aligned)
    torch.save(dict(feat=feat, rois=rois, scale=1 / 8, out=out, argmax=arg.int(),
                    rois_align=rois_ok, align=al), os.path.join(GOLD, "pool.pt"))

    # ---- (2a) alignment + row softmax ----------------------------------------------------------
    cases = {}
    for name, (M, D, K) in dict(small=(70, 96, 20), wide=(40, 64, 300)).items():
        x = torch.relu(torch.randn(M, D, generator=g))
        x[::17] = 0                                      # all-zero rows (eps path of F.normalize)
        text = torch.randn(K, D, generator=g)
        m = OpenVocabularyClassifier(ShapeSpec(channels=D), num_classes=K, weight_path="rand",
                                     weight_dim=D, norm_temperature=50.0)
        m.projection = nn.Identity()
        with torch.no_grad():
            logits = m(x, text, append_background=True)
            logits_nobg = m(x, text, append_background=False)
        head = types.SimpleNamespace()
        props = [list(range(M))]                         # len() only
        probs = fr.InstanceRefinementOutputLayers.predict_probs(head, (logits, None), props)[0]
        cases[name] = dict(x=x, text=text, T=50.0, logits=logits, logits_nobg=logits_nobg, probs=probs)
    torch.save(cases, os.path.join(GOLD, "align.pt"))

    # ---- (2b) MIL two-stream ---------------------------------------------------------------------
    cases = {}
    for name, (sizes, K) in dict(multi=((50, 1, 77), 20), single=((64,), 20), k1=((33, 9), 1)).items():
        Fdim = 16
        om = fr.ObjectMiningOutputLayers(ShapeSpec(channels=Fdim), box2box_transform=Box2BoxTransform((10, 10, 5, 5)),
                                         num_classes=K, loss_weight={})
        x = torch.randn(sum(sizes), Fdim, generator=g) * 4
        props = [list(range(s)) for s in sizes]
        with torch.no_grad():
            Cl, Dl = om.cls(x), om.det(x)
            scores, _ = om(x, props)
            img = om.predict_probs_img((scores, None), props)
            pb = om.predict_probs((scores, None), props)
        cases[name] = dict(cls=Cl, det=Dl, sizes=list(sizes), scores=scores, img=img,
                           probs_bg=torch.cat(pb, 0))
    torch.save(cases, os.path.join(GOLD, "mil.pt"))

    # ---- (2) fused alignment + MIL: ObjectMiningOutputLayers with the open-vocabulary class head (roi_heads.py:588-590) ---
    cases = {}
    g_main, g = g, torch.Generator().manual_seed(20261018)       # own stream: the sections after this one keep theirs
    for name, (sizes, K, D) in dict(two=((150, 97), 20, 64), one=((300,), 80, 96), tiny=((5, 1, 40), 7, 32)).items():
        Fdim = D
        ovc = OpenVocabularyClassifier(ShapeSpec(channels=Fdim), num_classes=K, weight_path="rand", weight_dim=D,
                                       norm_temperature=50.0)
        ovc.projection = nn.Identity()
        om = fr.ObjectMiningOutputLayers(ShapeSpec(channels=Fdim), box2box_transform=Box2BoxTransform((10, 10, 5, 5)),
                                         num_classes=K, class_head=ovc, loss_weight={})
        x = torch.relu(torch.randn(sum(sizes), Fdim, generator=g))
        x[::29] = 0
        props = [list(range(s)) for s in sizes]
        with torch.no_grad():
            om.det.weight.mul_(8.0)
            scores, _ = om(x, props)
            img = om.predict_probs_img((scores, None), props)
            cases[name] = dict(x=x, class_weight=ovc.class_weight.detach().clone(), det=om.det(x), sizes=list(sizes),
                               T=50.0, logits=ovc(x), scores=scores, img=img)
    torch.save(cases, os.path.join(GOLD, "align_mil.pt"))
    g = g_main

    # ---- (4) fast_rcnn_inference (filter + clip + batched_nms + top-k) -----------------------------
    sizes, K = (300, 260), 20
    img_shapes = [(240, 320), (200, 304)]
    boxes, probs = [], []
    for s, (ih, iw) in zip(sizes, img_shapes):
        b = make_rois(s, 1, ih, iw, g)[:, 1:]
        b[:10] += torch.randn(10, 4, generator=g) * 40       # boxes leaving the image -> clip matters
        lg = torch.randn(s, K + 1, generator=g) * 2.5
        boxes.append(b)
        probs.append(torch.softmax(lg, -1))
    probs[0][5, 3] = float("nan")                            # non-finite row is dropped (:178-182)
    for p, b in zip(probs, boxes):
        valid = torch.isfinite(b).all(1) & torch.isfinite(p).all(1)
        ncand = int((p[valid][:, :-1] > 1e-5).sum())
        assert ncand * 4 > 4000, "must exercise torchvision's vanilla batched_nms branch"
    inst, kept, _, _ = fr.fast_rcnn_inference(boxes, probs, img_shapes, 1e-5, 0.3, 100)
    torch.save(dict(boxes=boxes, probs=probs, image_shapes=img_shapes, score_thresh=1e-5,
                    nms_thresh=0.3, topk=100,
                    det_boxes=[i.pred_boxes.tensor for i in inst], det_scores=[i.scores for i in inst],
                    det_classes=[i.pred_classes for i in inst], kept_indices=kept,
                    det_rows=[i.pred_inds for i in inst]),
               os.path.join(GOLD, "detections.pt"))

    # ---- (3) refinement: get_pgt_top_k + label_and_sample_proposals_wsl ----------------------------
    K = 20
    sizes = (180, 90, 40)
    img_shapes = [(240, 320), (200, 304), (120, 160)]
    self = types.SimpleNamespace()
    self.num_classes = K
    self.images = [None] * len(sizes)
    gt_inst = [types.SimpleNamespace(gt_classes=torch.tensor(c)) for c in ([7, 2, 7, 15], [0], [19, 3])]
    _, self.gt_classes_img_int, oh = rh.get_image_level_gt(gt_inst, K)
    proposals, scores_l, boxes_l = [], [], []
    for i, (s, (ih, iw)) in enumerate(zip(sizes, img_shapes)):
        b = make_rois(s, 1, ih, iw, g)[:, 1:]
        if i == 2:
            b[:, 2] = b[:, 0] + 3.0
            b[:, 3] = b[:, 1] + 3.0                          # every box has area 9 <= 20 -> fallback seed
        inst = Instances((ih, iw))
        inst.proposal_boxes = Boxes(b)
        inst.objectness_logits = torch.rand(s, generator=g)
        proposals.append(inst)
        boxes_l.append(b)
        sc = torch.rand(s, K, generator=g) * 0.01
        scores_l.append(torch.cat([sc, torch.zeros(s, 1)], 1))   # predict_probs appends a bg column
    self.pred_class_img_logits = torch.rand(len(sizes), K, generator=g).clamp(1e-6, 1 - 1e-6)
    self.proposal_matchers = [Matcher([0.5], [0, 1], allow_low_quality_matches=False)]
    self.batch_size_per_images = [4096]
    self.positive_sample_fractions = [1.0]
    self.proposal_append_gt = False
    self.cls_agnostic_bbox_known = False
    self._sample_proposals_wsl = types.MethodType(rh.WSOVODROIHeads._sample_proposals_wsl, self)
    targets = rh.WSOVODROIHeads.get_pgt_top_k(self, boxes_l, scores_l, proposals)
    labelled = rh.WSOVODROIHeads.label_and_sample_proposals_wsl(self, 0, proposals, targets)
    torch.save(dict(
        boxes=boxes_l, scores=scores_l, sizes=list(sizes), num_classes=K,
        gt_classes_img=[t.clone() for t in self.gt_classes_img_int],
        img_scores=self.pred_class_img_logits,
        seed_boxes=[t.gt_boxes.tensor for t in targets], seed_classes=[t.gt_classes for t in targets],
        seed_scores=[t.gt_scores for t in targets], seed_weights=[t.gt_weights for t in targets],
        gt_classes=[p.gt_classes for p in labelled], gt_boxes=[p.gt_boxes.tensor for p in labelled],
        gt_scores=[p.gt_scores for p in labelled], gt_weights=[p.gt_weights for p in labelled],
    ), os.path.join(GOLD, "refine.pt"))

    # ---- (3') MIST seeds: WSOVODROIHeads.get_pgt_mist verbatim (roi_heads.py:910-1040, no SAM) -------------------
    g3 = torch.Generator().manual_seed(20261020)
    mist_sizes = (400, 120, 30)
    mist_shapes = [(240, 320), (200, 304), (120, 160)]
    mp, ms_l, mb_l = [], [], []
    for i, (s_, (ih, iw)) in enumerate(zip(mist_sizes, mist_shapes)):
        b = make_rois(s_, 1, ih, iw, g3)[:, 1:]
        if i == 2:
            b[::2, 2] = b[::2, 0] + 3.0
            b[::2, 3] = b[::2, 1] + 3.0                      # every other box has area 9 <= 20: filtered (:1090-1111)
        inst = Instances((ih, iw))
        inst.proposal_boxes = Boxes(b)
        mp.append(inst)
        mb_l.append(b)
        sc = torch.rand(s_, K, generator=g3) ** 6            # a few scores above the 0.05 threshold, most below
        ms_l.append(torch.cat([sc, torch.zeros(s_, 1)], 1))
    mself = types.SimpleNamespace(num_classes=K, images=[None] * 3, gt_classes_img_int=self.gt_classes_img_int,
                                  pred_class_img_logits=self.pred_class_img_logits)
    mself.get_pgt_top_k = types.MethodType(rh.WSOVODROIHeads.get_pgt_top_k, mself)
    mt = rh.WSOVODROIHeads.get_pgt_mist(mself, mb_l, ms_l, mp)
    torch.save(dict(boxes=mb_l, scores=ms_l, sizes=list(mist_sizes), shapes=mist_shapes, num_classes=K,
                    gt_classes_img=[t.clone() for t in self.gt_classes_img_int], img_scores=self.pred_class_img_logits,
                    seed_boxes=[t.gt_boxes.tensor for t in mt], seed_classes=[t.gt_classes for t in mt],
                    seed_scores=[t.gt_scores for t in mt], seed_weights=[t.gt_weights for t in mt]),
               os.path.join(GOLD, "mist.pt"))

    # ---- (3'') get_pgt_top_k with top_k != 1 / thres > 0 (roi_heads.py:1043-1343), on the MIST inputs: an integer
    # count with a threshold, a fraction without one, and per-class boxes (num_classes * 4 columns, :1066-1069)
    g4 = torch.Generator().manual_seed(20261021)
    per_class_boxes = [b.unsqueeze(1).repeat(1, K, 1) + torch.rand(b.size(0), K, 1, generator=g4) * 2.0 for b in mb_l]
    per_class_boxes = [b.reshape(b.size(0), K * 4) for b in per_class_boxes]
    topk_cases = {}
    for name, (bx, tk, th) in dict(count3_thres=(mb_l, 3, 0.02), frac=(mb_l, 0.05, 0), count5_perclass=(per_class_boxes, 5, 0.3),
                                   huge=(mb_l, 10000, 0.5)).items():
        tt = rh.WSOVODROIHeads.get_pgt_top_k(mself, bx, ms_l, mp, top_k=tk, thres=th)
        topk_cases[name] = dict(boxes=bx, top_k=tk, thres=th, seed_boxes=[t.gt_boxes.tensor for t in tt],
                                seed_classes=[t.gt_classes for t in tt], seed_scores=[t.gt_scores for t in tt],
                                seed_weights=[t.gt_weights for t in tt])
    torch.save(dict(scores=ms_l, sizes=list(mist_sizes), shapes=mist_shapes, num_classes=K,
                    gt_classes_img=[t.clone() for t in self.gt_classes_img_int], img_scores=self.pred_class_img_logits,
                    cases=topk_cases), os.path.join(GOLD, "pgt_topk.pt"))

    # ---- (3b) weighted refinement losses: InstanceRefinementOutputLayers.losses verbatim ---
    cases = {}
    for name, (dcols_per, beta, reg) in dict(agnostic=(1, 0.0, True), specific=(K, 0.5, True), noreg=(1, 0.0, False)).items():
        head = types.SimpleNamespace(
            cross_entropy_weighted=True, box_reg_loss_type="smooth_l1_weighted", refine_k=0, refine_reg=[reg],
            loss_weight={}, box2box_transform=Box2BoxTransform((10.0, 10.0, 5.0, 5.0)), smooth_l1_beta=beta,
            num_classes=K)
        head.softmax_cross_entropy_loss = types.MethodType(fr.InstanceRefinementOutputLayers.softmax_cross_entropy_loss, head)
        head.box_reg_loss = types.MethodType(fr.InstanceRefinementOutputLayers.box_reg_loss, head)
        props = []
        for p in labelled:
            q = Instances(p.image_size)
            q.proposal_boxes = p.proposal_boxes
            q.gt_boxes = p.gt_boxes
            gc = p.gt_classes.clone()
            gc[::11] = -1                                        # ignored rows (subsampling leaves -1)
            q.gt_classes = gc
            q.gt_weights = p.gt_weights.clone()
            props.append(q)
        M = sum(len(p) for p in props)
        logits = (torch.randn(M, K + 1, generator=g) * 3).requires_grad_()
        deltas = (torch.randn(M, 4 * dcols_per, generator=g) * 0.5).requires_grad_()
        out = fr.InstanceRefinementOutputLayers.losses(head, (logits, deltas), props)
        total = sum(out.values())
        total.backward()
        cases[name] = dict(
            logits=logits.detach(), deltas=deltas.detach(), gt_classes=torch.cat([p.gt_classes for p in props]),
            gt_weights=torch.cat([p.gt_weights for p in labelled]),
            proposal_boxes=torch.cat([p.proposal_boxes.tensor for p in props]),
            gt_boxes=torch.cat([p.gt_boxes.tensor for p in props]), num_classes=K, beta=beta, reg=reg,
            loss_cls=out["loss_cls_r0"].detach(), loss_box=out.get("loss_box_reg_r0", torch.zeros(())).detach(),
            grad_logits=logits.grad.clone(), grad_deltas=(deltas.grad.clone() if deltas.grad is not None else torch.zeros_like(deltas)))
    torch.save(cases, os.path.join(GOLD, "refine_loss.pt"))

    # ---- (0) proposal ingest: unique_boxes / transform_proposals / the per-image sort --------
    import ast
    import numpy as np
    src = open(os.path.join(REF, "wsovod", "data", "detection_utils.py")).read()
    fns = [n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name in ("unique_boxes", "transform_proposals")]
    ns = dict(np=np, torch=torch, Boxes=Boxes, Instances=Instances,
              BoxMode=types.SimpleNamespace(XYXY_ABS=0, convert=lambda b, frm, to: b))
    exec(compile(ast.Module(body=fns, type_ignores=[]), "detection_utils.py", "exec"), ns)
    rng = np.random.RandomState(7)
    cases = {}
    for name, (n, ih, iw, topk, msz) in dict(plain=(600, 480, 640, 400, 0), tiny=(300, 120, 160, 4000, 2), few=(5, 50, 60, 3, 0)).items():
        bx = make_rois(n, 1, ih, iw, g)[:, 1:].numpy().astype(np.float64)
        bx[: n // 4] = np.round(bx[: n // 4])                       # integer boxes: hash collisions after clip
        bx[n // 4: n // 3] += rng.uniform(-0.4, 0.4, (n // 3 - n // 4, 4))   # near-duplicates that round together
        bx[-n // 10:, 2:] += 300.0                                   # partly outside the image
        lg = rng.rand(n)
        lg[: n // 5] = np.round(lg[: n // 5], 1)                     # tied scores
        inds = lg.argsort()[::-1]                                    # build.py:166-168
        sb, sl = bx[inds], lg[inds]
        dd = dict(proposal_boxes=sb.copy(), proposal_objectness_logits=sl.copy(), proposal_bbox_mode=0)
        tf = types.SimpleNamespace(apply_box=lambda b: b)
        ns["transform_proposals"](dd, (ih, iw), tf, proposal_topk=topk, min_box_size=msz)
        cases[name] = dict(boxes=bx, logits=lg, image_shape=(ih, iw), topk=topk, min_box_size=msz,
                           sorted_boxes=sb, sorted_logits=sl,
                           unique=ns["unique_boxes"](Boxes(torch.as_tensor(sb).float())),
                           out_boxes=dd["proposals"].proposal_boxes.tensor, out_logits=dd["proposals"].objectness_logits)
    torch.save(cases, os.path.join(GOLD, "ingest.pt"))

    # ---- (4b) RPN proposal selection: find_top_rpn_proposals verbatim ------------------------
    src = open(os.path.join(REF, "wsovod", "modeling", "proposal_generator", "proposal_utils.py")).read()
    fns = [n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name in ("find_top_rpn_proposals", "_is_tracing")]
    from typing import List as _List, Tuple as _Tuple
    ns = dict(torch=torch, List=_List, Tuple=_Tuple, Boxes=Boxes, Instances=Instances, cat=d2_shim.cat,
              batched_nms=d2_shim.batched_nms, move_device_like=lambda src_, dst_: src_)
    exec(compile(ast.Module(body=fns, type_ignores=[]), "proposal_utils.py", "exec"), ns)
    cases = {}
    for name, (Nimg, lvls, pre, post, msz) in dict(single=(2, (3000,), 2000, 1000, 0.0), fpn=(3, (2500, 700, 200, 40), 1000, 600, 4.0)).items():
        sizes = [(480, 640), (400, 600), (300, 512)][:Nimg]
        props = [torch.stack([make_rois(n, 1, 480, 640, g)[:, 1:] for _ in range(Nimg)]) for n in lvls]
        for p in props:
            p[:, ::13] += torch.randn(Nimg, p[:, ::13].shape[1], 4, generator=g) * 200       # out of the image / inverted
        logits = [torch.randn(Nimg, n, generator=g) for n in lvls]
        logits[0][0, 5] = float("nan")
        props[0][1, 9, 2] = float("inf")
        res = ns["find_top_rpn_proposals"]([p.clone() for p in props], [l.clone() for l in logits], sizes, 0.7, pre, post, msz, False)
        cases[name] = dict(proposals=props, logits=logits, image_sizes=sizes, nms_thresh=0.7, pre=pre, post=post, min_box_size=msz,
                           boxes=[r.proposal_boxes.tensor for r in res], scores=[r.objectness_logits for r in res])
    torch.save(cases, os.path.join(GOLD, "rpn_select.pt"))

    # ---- (4c) grouped RPN selection with the CSC re-weighting: find_top_rpn_proposals_group verbatim (:146-362); its
    #      `csc` (wsovod._C.csc_forward, GPU only) is the oracle's restatement here -- pinned against the compiled
    #      reference extension on the GPU box (tests/test_csc.py)
    import oracle as orc
    g2 = torch.Generator().manual_seed(20261019)
    fns = [n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name in ("find_top_rpn_proposals_group", "_is_tracing")]

    def conv(box_lists):
        return torch.cat([torch.cat((torch.full_like(b.tensor[:, :1], i), b.tensor), dim=1) for i, b in enumerate(box_lists)], dim=0)

    def csc_ref(cp, lab, pr, rois, tau, dbg, fg, mass, dens, area_sqrt, ctx):
        return orc.csc(cp, lab, pr, rois, fg, area_sqrt, ctx), lab.clone(), torch.zeros_like(lab)
    ns = dict(torch=torch, List=_List, Tuple=_Tuple, Boxes=Boxes, Instances=Instances, cat=d2_shim.cat,
              batched_nms=d2_shim.batched_nms, move_device_like=lambda src_, dst_: src_, csc=csc_ref,
              convert_boxes_to_pooler_format=conv, get_event_storage=d2_shim.get_event_storage)
    exec(compile(ast.Module(body=fns, type_ignores=[]), "proposal_utils.py", "exec"), ns)
    cases = {}
    for name, (Nimg, lvls, anchors, pre, post, with_cpg) in dict(plain=(2, (900, 300), (3, 3), 200, 300, False),
                                                                csc=(2, (1200,), (3,), 300, 250, True)).items():
        sizes = [(480, 640), (400, 600)][:Nimg]
        props = [torch.stack([make_rois(n, 1, 480, 640, g2)[:, 1:] for _ in range(Nimg)]) for n in lvls]
        logits = [torch.randn(Nimg, n, generator=g2) for n in lvls]
        cpgs = strides = None
        if with_cpg:
            cpgs = []
            for (h, w_) in sizes:
                m = torch.rand(h // 8, w_ // 8, generator=g2) * 0.09
                m[10:30, 20:50] = 0.6
                cpgs.append(m)
            strides = (8, 8)
        res = ns["find_top_rpn_proposals_group"]([p.clone() for p in props], [l.clone() for l in logits], sizes, list(anchors),
                                                 0.7, pre, post, 0.0, False, cpgs, strides)
        cases[name] = dict(proposals=props, logits=logits, image_sizes=sizes, num_anchors=list(anchors), nms_thresh=0.7, pre=pre,
                           post=post, cpgs=cpgs, cpg_strides=strides, boxes=[r.proposal_boxes.tensor for r in res],
                           scores=[r.objectness_logits for r in res], level_ids=[r.level_ids for r in res])
    torch.save(cases, os.path.join(GOLD, "rpn_group.pt"))

    for f in sorted(os.listdir(GOLD)):
        print(f, os.path.getsize(os.path.join(GOLD, f)))


if __name__ == "__main__":
    main()
