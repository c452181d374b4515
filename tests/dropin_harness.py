"""Shared harness of the drop-in tests: builds the REFERENCE's own WSOVODROIHeads (roi_heads.py:430-907, imported
verbatim through oracle/d2_shim.py from /root/reference or the shipped oracle/_ref/py copies) and this package's
WSOVODROIHeads over the same weights and inputs.  Test infrastructure only."""
import types

import torch
from torch import nn

from wsovod_b200 import synth


class BoxHead(nn.Module):
    """stand-in for detectron2's FastRCNNConvFCHead (box_head.py:59-68: flatten -> FC -> ReLU -> FC -> ReLU); the FCs
    are out of scope and stay PyTorch in both arms"""

    def __init__(self, in_features, width):
        super().__init__()
        self.fc1, self.fc2 = nn.Linear(in_features, width), nn.Linear(width, width)
        self.output_shape = types.SimpleNamespace(channels=width, height=None, width=None, stride=None)

    def forward(self, x):
        x = torch.flatten(x, start_dim=1)
        return torch.relu(self.fc2(torch.relu(self.fc1(x))))


def reference_modules():
    from oracle import d2_shim
    d2_shim.install()
    import wsovod.modeling.poolers as poolers
    import wsovod.modeling.roi_heads.fast_rcnn_open_vocabulary as fr
    import wsovod.modeling.roi_heads.roi_heads as rh
    return d2_shim, rh, fr, poolers


def build_reference(device, *, C=16, K=20, D=32, width=48, pooler_type="ROIPool", refine_K=1, refine_reg=True,
                    batch_size=4096, positive_fraction=1.0, weight_path="rand", seed=0, topk=100, score_thresh=1e-5,
                    mods=None, mrrp=False):
    """the reference WSOVODROIHeads built through its own constructors (keyword path of @configurable)"""
    d2_shim, rh, fr, poolers = mods or reference_modules()
    from oracle.d2_shim import Box2BoxTransform, Matcher, ShapeSpec
    torch.manual_seed(seed)
    head = BoxHead(C * 49, width)
    shape = ShapeSpec(channels=width)
    b2b = lambda: Box2BoxTransform((10.0, 10.0, 5.0, 5.0))  # noqa: E731
    miner = fr.ObjectMiningOutputLayers(shape, box2box_transform=b2b(), num_classes=K, loss_weight={})
    refinery = []
    for k in range(refine_K):
        ovc = rh.OpenVocabularyClassifier(shape, num_classes=K, weight_path=weight_path, weight_dim=D, norm_temperature=50.0)
        refinery.append(fr.InstanceRefinementOutputLayers(
            shape, box2box_transform=b2b(), num_classes=K, class_head=ovc, test_score_thresh=score_thresh, test_nms_thresh=0.3,
            test_topk_per_image=topk, smooth_l1_beta=0.0, box_reg_loss_type="smooth_l1_weighted", loss_weight={},
            refine_k=k, refine_reg=[refine_reg] * refine_K, cross_entropy_weighted=True))
    # MRRP (roi_heads.py:567-573): three dilation branches = three pooler "levels" at one scale
    pooler = rh.ROIPooler(output_size=7, scales=(1.0 / synth.STRIDE,) * (3 if mrrp else 1), sampling_ratio=0, pooler_type=pooler_type)
    heads = rh.WSOVODROIHeads(
        mrrp_on=mrrp, mrrp_num_branch=3,
        num_classes=K, batch_size_per_image=512, positive_fraction=0.25, proposal_matcher=Matcher([0.5], [0, 1]),
        proposal_append_gt=False, pixel_mean=(103.53, 116.28, 123.675), pixel_std=(1.0, 1.0, 1.0),
        box_in_features=["res5"], box_pooler=pooler, box_head=head, object_miner=miner, sam=None, refine_K=refine_K,
        refine_mist=False, refine_reg=[refine_reg] * refine_K, box_refinery=refinery, sampling_on=True,
        proposal_matchers=[Matcher([0.5], [0, 1]) for _ in range(refine_K)], batch_size_per_images=[batch_size] * refine_K,
        positive_sample_fractions=[positive_fraction] * refine_K, pooler_type=pooler_type, rpn_on=False, metadata=None)
    return heads.to(device)


def build_ours(ref_heads, device, *, precision, pooler_type="ROIPool", mrrp=False):
    """this package's WSOVODROIHeads holding copies of the reference head's weights (same state_dict keys)"""
    from wsovod_b200 import ops  # noqa: F401
    from wsovod_b200.modeling import (InstanceRefinementOutputLayers, ObjectMiningOutputLayers, OpenVocabularyClassifier,
                                      ROIPooler, WSOVODROIHeads)
    K = ref_heads.num_classes
    width = ref_heads.box_head.fc2.out_features
    head = BoxHead(ref_heads.box_head.fc1.in_features, width)
    head.load_state_dict(ref_heads.box_head.state_dict())
    miner = ObjectMiningOutputLayers(width, K)
    miner.load_state_dict(ref_heads.object_miner.state_dict())
    refinery = []
    for r in ref_heads.box_refinery:
        D = r.cls.weight_dim
        ovc = OpenVocabularyClassifier(width, num_classes=K, weight_path="rand", weight_dim=D,
                                       norm_temperature=r.cls.norm_temperature, precision=precision)
        reg = bool(r.refine_reg[r.refine_k])
        m = InstanceRefinementOutputLayers(width, K, ovc, test_score_thresh=r.test_score_thresh,
                                           test_nms_thresh=r.test_nms_thresh, test_topk_per_image=r.test_topk_per_image,
                                           refine_reg=reg)
        m.load_state_dict(r.state_dict())
        refinery.append(m)
    pooler = ROIPooler(7, (1.0 / synth.STRIDE,) * (3 if mrrp else 1), 0, pooler_type)
    ours = WSOVODROIHeads(mrrp_on=mrrp, mrrp_num_branch=3, num_classes=K, box_in_features=["res5"], box_pooler=pooler, box_head=head, object_miner=miner,
                          box_refinery=refinery, refine_reg=[bool(r.refine_reg[r.refine_k]) for r in ref_heads.box_refinery],
                          sampling_on=True, batch_size_per_images=list(ref_heads.batch_size_per_images),
                          positive_sample_fractions=list(ref_heads.positive_sample_fractions), pooler_type=pooler_type)
    return ours.to(device)


def make_inputs(device, *, N=2, C=16, H=30, W=40, R=300, K=20, D=32, seed=1, max_labels=3, mrrp=False):
    """features dict, proposals (the reference's own Instances / Boxes carriers), image-level targets, text matrix"""
    from oracle.d2_shim import Boxes, Instances
    g = synth.gen(seed)
    feat = synth.features(3 * N if mrrp else N, C, H, W, g).to(device)     # MRRP: the three branches stacked on the batch axis
    img_h, img_w = H * synth.STRIDE, W * synth.STRIDE
    proposals, targets = [], []
    labels = synth.image_labels(N, K, g, max_labels)
    for n in range(N):
        b = synth.proposals(R, img_h, img_w, g)
        p = Instances((img_h, img_w))
        p.proposal_boxes = Boxes(b.to(device))
        p.objectness_logits = synth.objectness(R, g).to(device)
        if mrrp:   # roi_heads.py:730: branch = level_ids // 1000
            p.level_ids = (torch.randint(0, 3, (R,), generator=g) * 1000 + torch.randint(0, 1000, (R,), generator=g)).to(device)
        proposals.append(p)
        t = Instances((img_h, img_w))
        t.gt_classes = labels[n].to(device)
        targets.append(t)
    text = synth.text_embeddings(K, D, g).to(device)
    return {"res5": feat}, proposals, targets, text


def run_train(heads, rh, features, proposals, targets, classifier=None, seed=5):
    """_forward_box in training mode (roi_heads.py:761-884) -> (losses, gradient of the summed loss w.r.t. fc1.weight)"""
    heads.train()
    heads.gt_classes_img, heads.gt_classes_img_int, heads.gt_classes_img_oh = rh.get_image_level_gt(targets, heads.num_classes)
    heads.images = [None] * len(proposals)
    for p in heads.parameters():
        p.grad = None
    torch.manual_seed(seed)
    losses = heads._forward_box(features, proposals, None, classifier, True)
    sum(losses.values()).backward()
    return {k: v.detach() for k, v in losses.items()}, heads.box_head.fc1.weight.grad.clone()


@torch.no_grad()
def run_test(heads, features, proposals, classifier):
    heads.eval()
    inst, all_scores, all_boxes = heads._forward_box(features, proposals, None, classifier, True)
    return inst
