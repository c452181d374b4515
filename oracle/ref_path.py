"""The reference's CPU path of the inference slice, executed by the REFERENCE'S OWN CODE (TEST INFRASTRUCTURE: only
bench.py's `cpu_baseline` / `--impl reference` legs and tests import this).

`get()` returns the implementation to time and its kind:
  "reference"  this module: the reference's Python files (imported verbatim through oracle/d2_shim.py from
               /root/reference, or from the byte-for-byte copies oracle/build_ref.py ships under the git-ignored
               oracle/_ref/py on the GPU box) run the slice -- wsovod.modeling.poolers.ROIPooler ("ROIPool" ->
               torchvision's CPU kernel, poolers.py:183-186,221-284), the objectness scale exactly as
               roi_heads.py:733-739 writes it, InstanceRefinementOutputLayers.forward -> OpenVocabularyClassifier.forward
               (projection = Identity: the FC stack is out of scope and excluded in both arms,
               open_vocabulary_classifier.py:79-105) and InstanceRefinementOutputLayers.inference ->
               predict_probs_K / predict_boxes_K / fast_rcnn_inference (fast_rcnn_open_vocabulary.py:52-217,894-1058);
  "port"       oracle/cpu_path.py, the same library calls restated, when no reference tree is available.
All host threads (torch.set_num_threads by the caller), MODEL.DEVICE=cpu semantics.
"""
import time

import torch

from . import cpu_path, d2_shim

DESCRIPTION = ("the reference's own Python (ROIPooler -> torchvision CPU roi_pool, objectness scale, OpenVocabularyClassifier "
               "contraction + softmax, fast_rcnn_inference -> torchvision CPU batched_nms + top100), box-head FCs excluded")
cpu_path.DESCRIPTION = ("port of the reference path: torchvision CPU roi_pool + objectness scale, ATen normalize/mm/softmax, "
                        "torchvision CPU batched_nms + top100")

_state = {}


def get():
    if d2_shim.default_reference_root() is None:
        return cpu_path, "port"
    import sys
    return sys.modules[__name__], "reference"


def _objects(w):
    key = (w["K"], w["D"], w["spatial_scale"], w["score_thresh"], w["nms_thresh"], w["topk"])
    if _state.get("key") != key:
        d2_shim.install()
        import wsovod.modeling.poolers as poolers
        import wsovod.modeling.roi_heads.fast_rcnn_open_vocabulary as fr
        from wsovod.modeling.class_heads import OpenVocabularyClassifier
        K, D = w["K"], w["D"]
        ovc = OpenVocabularyClassifier(d2_shim.ShapeSpec(channels=D), num_classes=K, weight_path="rand", weight_dim=D,
                                       norm_temperature=w["temperature"])
        ovc.projection = torch.nn.Identity()
        head = fr.InstanceRefinementOutputLayers(
            d2_shim.ShapeSpec(channels=D), box2box_transform=d2_shim.Box2BoxTransform((10.0, 10.0, 5.0, 5.0)), num_classes=K,
            class_head=ovc, test_score_thresh=w["score_thresh"], test_nms_thresh=w["nms_thresh"],
            test_topk_per_image=w["topk"], refine_k=0, refine_reg=[False], loss_weight={})
        pooler = poolers.ROIPooler(output_size=7, scales=(w["spatial_scale"],), sampling_ratio=0, pooler_type="ROIPool")
        _state.update(key=key, pooler=pooler.eval(), head=head.eval())
    return _state["pooler"], _state["head"]


@torch.no_grad()
def run_slice(w, images=None, proposals=None):
    """One pass of the inference slice over a (sub)sample of workload `w` (wsovod_b200.synth.workload dict, CPU
    tensors), the statements of WSOVODROIHeads._forward_box (roi_heads.py:722-746,886-899) with the FC stack cut out.
    Returns (seconds, proposals processed, pooled, probs, instances)."""
    pooler, head = _objects(w)
    N = w["N"] if images is None else min(images, w["N"])
    R = w["R"] if proposals is None else min(proposals, w["R"])
    props, emb = [], []
    for n in range(N):
        r0 = w["offsets"][n]
        p = d2_shim.Instances((float(w["image_sizes"][n, 0]), float(w["image_sizes"][n, 1])))
        p.proposal_boxes = d2_shim.Boxes(w["rois"][r0:r0 + R, 1:])
        p.objectness_logits = w["objectness"][r0:r0 + R]
        props.append(p)
        emb.append(w["region_emb"][r0:r0 + R])
    emb = torch.cat(emb)
    features = [w["features"][:N]]
    t0 = time.perf_counter()
    box_features = pooler(features, [x.proposal_boxes for x in props], level_ids=None)            # roi_heads.py:727-731
    objectness_logits = torch.cat([x.objectness_logits + 1 for x in props], dim=0)                # :733
    box_features = box_features * objectness_logits.view(-1, 1, 1, 1)                             # :739
    # [self.box_head(box_features): out of scope; the region embeddings of the workload stand for its output]
    predictions_K = [head(emb, w["text_emb"], True)]                                              # :890-892
    inst, _, all_scores, _ = head.inference(predictions_K, props)                                 # :893-895
    dt = time.perf_counter() - t0
    return dt, N * R, box_features, torch.cat([s.squeeze(0) for s in all_scores]), inst
