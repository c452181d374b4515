"""CPU tests: the oracle against (a) goldens produced by the reference code itself (tests/golden, see
oracle/make_golden.py), (b) torchvision's compiled CPU ops, (c) the reference's own C++ in oracle/_ref."""
import pytest
import torch

import oracle
from oracle import ref
from wsovod_b200 import synth


def _offs(sizes):
    o = [0]
    for s in sizes:
        o.append(o[-1] + int(s))
    return o


def test_pool_golden(golden):
    p = golden("pool")
    o, a = oracle.roi_pool(p["feat"], p["rois"], p["scale"], 7)
    assert torch.equal(o, p["out"]) and torch.equal(a, p["argmax"])
    for (sr, al), v in p["align"].items():
        assert torch.equal(oracle.roi_align(p["feat"], p["rois_align"], p["scale"], 7, sr, al), v)


def test_pool_vs_torchvision_cpu():
    import torchvision  # noqa: F401
    g = synth.gen(1)
    for trial in range(3):
        feat = synth.features(2, 9, 30, 40, g, relu=False)
        rois, _ = synth.rois_from([synth.proposals(200, 240, 320, g) for _ in range(2)])
        rois[:20, 1:] += torch.randn(20, 4, generator=g) * 150
        o, a = oracle.roi_pool(feat, rois, 1 / 8, 7)
        to, ta = torch.ops.torchvision.roi_pool(feat, rois, 1 / 8, 7, 7)
        assert torch.equal(o, to) and torch.equal(a, ta.int())
        assert (a == -1).any()          # empty bins are exercised


def test_pool_vs_reference_cpp():
    m = ref.cpu()
    if m is None:
        pytest.skip("oracle/_ref not built (python oracle/build_ref.py)")
    g = synth.gen(2)
    feat = synth.features(2, 6, 30, 40, g, relu=False)
    rois, _ = synth.rois_from([synth.proposals(200, 240, 320, g) for _ in range(2)])
    rois[:20, 1:] += torch.randn(20, 4, generator=g) * 150
    o, a = m.roi_pool_forward_cpu(feat, rois, 1 / 8, 7, 7)          # ROILoopPool_cpu.cpp:125
    oo, aa = oracle.roi_pool(feat, rois, 1 / 8, 7)
    assert torch.equal(o, oo) and torch.equal(a, aa)
    go = torch.randn_like(o)
    gi = m.roi_pool_backward_cpu(go, rois, a, 1 / 8, 7, 7, 2, 6, 30, 40)
    torch.testing.assert_close(gi, oracle.roi_pool_backward(go, rois, aa, feat.shape), rtol=1e-6, atol=1e-6)


def test_loop_pool_first_block_is_relu_pool():
    # ROILoopPool's first R rows equal ROIPool clamped at 0 (ROILoopPool_cuda.cu:107-113)
    g = synth.gen(4)
    feat = synth.features(1, 3, 30, 40, g, relu=False)
    rois, _ = synth.rois_from([synth.proposals(100, 240, 320, g)])
    o3, a3 = oracle.roi_loop_pool(feat, rois, 1 / 8, 7)
    o1, a1 = oracle.roi_pool(feat, rois, 1 / 8, 7)
    R = rois.size(0)
    assert torch.equal(o3[:R], o1.clamp(min=0))
    pos = o1 > 0
    assert torch.equal(a3[:R][pos], a1[pos]) and (a3[:R][~pos] == -1).all()
    # frame <= roi (subset of cells), context is pooled over a larger box
    assert (o3[R:2 * R] <= o3[:R]).all()


def test_align_golden(golden):
    for name, c in golden("align").items():
        lg, pr = oracle.align(c["x"], c["text"], c["T"], True, True)
        torch.testing.assert_close(lg, c["logits"], rtol=0, atol=5e-5)
        torch.testing.assert_close(pr, c["probs"], rtol=1e-4, atol=1e-9)
        lg2, _ = oracle.align(c["x"], c["text"], c["T"], True, False, want_probs=False)
        torch.testing.assert_close(lg2, c["logits_nobg"], rtol=0, atol=5e-5)


def test_mil_golden(golden):
    for name, c in golden("mil").items():
        s, i = oracle.mil(c["cls"], c["det"], _offs(c["sizes"]))
        torch.testing.assert_close(s, c["scores"], rtol=1e-5, atol=1e-12)
        torch.testing.assert_close(i, c["img"], rtol=1e-5, atol=1e-8)
        assert torch.equal(c["probs_bg"][:, :-1], c["scores"]) and (c["probs_bg"][:, -1] == 0).all()


def test_refine_golden(golden):
    f = golden("refine")
    off = _offs(f["sizes"])
    goff = _offs([len(x) for x in f["gt_classes_img"]])
    sd = oracle.pgt_top1(torch.cat(f["scores"]), torch.cat(f["boxes"]), off, torch.cat(f["gt_classes_img"]), goff,
                         f["img_scores"])
    ra = oracle.refine_assign(torch.cat(f["boxes"]), off, sd["seed_boxes"], sd["seed_classes"], sd["seed_scores"],
                              sd["seed_weights"], goff, sd["seed_count"], f["num_classes"], 0.5)
    for n in range(len(f["sizes"])):
        g0, c = goff[n], int(sd["seed_count"][n])
        for k in ("seed_boxes", "seed_classes", "seed_scores", "seed_weights"):
            assert torch.equal(sd[k][g0:g0 + c], f[k][n]), (n, k)
        sl = slice(off[n], off[n + 1])
        for k in ("gt_classes", "gt_boxes", "gt_scores", "gt_weights"):
            assert torch.equal(ra[k][sl], f[k][n]), (n, k)
    assert int(sd["seed_count"][2]) == 1 and sd["seed_boxes"][goff[2]].tolist() == [-10000, -10000, 10000, 10000]


def test_detections_golden(golden):
    d = golden("detections")
    off = _offs([len(b) for b in d["boxes"]])
    r = oracle.detections(torch.cat(d["probs"]), torch.cat(d["boxes"]), off, d["image_shapes"], d["score_thresh"],
                          d["nms_thresh"], d["topk"], oracle.IOU_TV_CPU)
    for n in range(len(d["boxes"])):
        c = int(r["det_count"][n])
        assert c == len(d["det_scores"][n])
        assert torch.equal(r["det_rows"][n, :c], d["det_rows"][n])          # == Instances.pred_inds
        assert torch.equal(r["det_classes"][n, :c], d["det_classes"][n])
        assert torch.equal(r["det_scores"][n, :c], d["det_scores"][n])
        assert torch.equal(r["det_boxes"][n, :c], d["det_boxes"][n])


@pytest.mark.parametrize("grid", [None, 4, 16])
def test_nms_vs_torchvision_cpu(grid):
    from torchvision.ops.boxes import _batched_nms_vanilla
    g = synth.gen(7)
    for trial in range(3):
        b = synth.proposals(2500, 480, 640, g)
        if grid:
            b = (b / grid).round() * grid      # exact IoU == threshold pairs, duplicates, zero-area boxes
        s = torch.rand(2500, generator=g)
        idx = torch.randint(0, 20, (2500,), generator=g)
        assert torch.equal(oracle.batched_nms(b, s, idx, 0.3, oracle.IOU_TV_CPU), _batched_nms_vanilla(b, s, idx, 0.3))
    # the double-vs-float threshold probe of SURVEY C-9
    b = torch.tensor([[0., 0., 13., 1.], [7., 0., 20., 1.]])
    s = torch.tensor([0.9, 0.8])
    z = torch.zeros(2, dtype=torch.int64)
    assert oracle.batched_nms(b, s, z, 0.3, oracle.IOU_TV_CPU).tolist() == [0]
    assert oracle.batched_nms(b, s, z, 0.3, oracle.IOU_TV_CUDA).tolist() == [0, 1]


def test_cpu_path_matches_oracle():
    """bench's CPU baseline (oracle/cpu_path.py: the reference's library calls) agrees with the C oracle."""
    from oracle import cpu_path
    w = synth.workload("c1")
    w = dict(w)
    dt, n, pooled, probs, dets = cpu_path.run_slice(w, images=1, proposals=300)
    rois = w["rois"][:300]
    o, _ = oracle.roi_pool(w["features"], rois, w["spatial_scale"], 7)
    assert torch.equal(pooled, o * (w["objectness"][:300] + 1).view(-1, 1, 1, 1))
    lg, pr = oracle.align(w["region_emb"][:300], w["text_emb"], 50.0, True, True)
    torch.testing.assert_close(probs, pr, rtol=1e-4, atol=1e-9)
    r = oracle.detections(probs, rois[:, 1:], [0, 300], [(480, 640)], 1e-5, 0.3, 100, oracle.IOU_TV_CPU)
    c = int(r["det_count"][0])
    assert torch.equal(r["det_rows"][0, :c], dets[0][3]) and torch.equal(r["det_classes"][0, :c], dets[0][2])


def test_refine_losses_oracle_matches_reference_golden(golden):
    """the torch restatement of InstanceRefinementOutputLayers.losses against the reference's own function
    (values and autograd gradients), tests/golden/refine_loss.pt"""
    for name, c in golden("refine_loss").items():
        logits = c["logits"].clone().requires_grad_()
        deltas = c["deltas"].clone().requires_grad_() if c["reg"] else None
        lc, lb = oracle.refine_losses(logits, deltas, c["gt_classes"], c["gt_weights"], c["proposal_boxes"], c["gt_boxes"],
                                      c["num_classes"], beta=c["beta"])
        torch.testing.assert_close(lc.detach(), c["loss_cls"], rtol=1e-6, atol=1e-7)
        torch.testing.assert_close(lb.detach(), c["loss_box"], rtol=1e-6, atol=1e-8)
        (lc + lb).backward()
        torch.testing.assert_close(logits.grad, c["grad_logits"], rtol=1e-6, atol=1e-9)
        if deltas is not None:
            torch.testing.assert_close(deltas.grad, c["grad_deltas"], rtol=1e-6, atol=1e-9)


def test_align_mil_composition_vs_reference_class_head_miner(golden):
    """oracle.align_mil (align without background -> mil) against the reference's own
    ObjectMiningOutputLayers(class_head=OpenVocabularyClassifier) (fast_rcnn_open_vocabulary.py:280-285,318-367)"""
    for name, c in golden("align_mil").items():
        off = [0]
        for s in c["sizes"]:
            off.append(off[-1] + s)
        s, img, lg = oracle.align_mil(c["x"], c["class_weight"].t().contiguous(), c["det"], off, c["T"], 2)
        assert (lg - c["logits"]).abs().max().item() <= 5e-5, name
        torch.testing.assert_close(s, c["scores"], rtol=2e-4, atol=1e-7)
        torch.testing.assert_close(img, c["img"], rtol=2e-4, atol=1e-7)


def _mist_check(golden, device, **kw):
    from wsovod_b200.modeling import get_pgt_mist
    from wsovod_b200.structures import Boxes, Instances
    f = golden("mist")
    props = [Instances(tuple(s), proposal_boxes=Boxes(b.to(device))) for b, s in zip(f["boxes"], f["shapes"])]
    targets, seeds = get_pgt_mist([b.to(device) for b in f["boxes"]], [s.to(device) for s in f["scores"]], props,
                                  [g.to(device) for g in f["gt_classes_img"]], f["img_scores"].to(device), f["num_classes"], **kw)
    for n, t in enumerate(targets):
        assert torch.equal(t.gt_boxes.tensor.cpu(), f["seed_boxes"][n])
        assert torch.equal(t.gt_classes.cpu(), f["seed_classes"][n])
        assert torch.equal(t.gt_scores.cpu(), f["seed_scores"][n])
        assert torch.equal(t.gt_weights.cpu(), f["seed_weights"][n])
    assert seeds["seed_offsets"].tolist()[-1] == sum(len(t.gt_classes) for t in targets)


def test_get_pgt_top_k_general_selection_vs_reference(golden):
    """WSOVODROIHeads.get_pgt_top_k with top_k != 1 / thres > 0 (roi_heads.py:1043-1343): an integer count with a
    threshold, a fraction, per-class boxes, a count larger than the image -- bit-exact against the reference's method
    (golden generated by oracle/make_golden.py from the verbatim reference), fallback seed of the empty image included"""
    from wsovod_b200.modeling.roi_heads import get_pgt_top_k
    from wsovod_b200.structures import Boxes, Instances
    f = golden("pgt_topk")
    for name, c in f["cases"].items():
        props = [Instances(tuple(s), proposal_boxes=Boxes(b.reshape(b.size(0), -1)[:, :4])) for b, s in zip(c["boxes"], f["shapes"])]
        targets, seeds = get_pgt_top_k(c["boxes"], f["scores"], props, f["gt_classes_img"], f["img_scores"], f["num_classes"],
                                       top_k=c["top_k"], thres=c["thres"])
        for n, t in enumerate(targets):
            assert torch.equal(t.gt_boxes.tensor, c["seed_boxes"][n]), (name, n)
            assert torch.equal(t.gt_classes, c["seed_classes"][n]), (name, n)
            assert torch.equal(t.gt_scores, c["seed_scores"][n]), (name, n)
            assert torch.equal(t.gt_weights, c["seed_weights"][n]), (name, n)
        assert seeds["seed_offsets"].tolist()[-1] == sum(len(t.gt_classes) for t in targets)
        assert seeds["seed_boxes"].size(0) == seeds["seed_offsets"].tolist()[-1]


def test_get_pgt_mist_vs_reference_with_oracle_nms(golden):
    """WSOVODROIHeads.get_pgt_mist (roi_heads.py:910-1040) mirror: top 15 % per class, 0.05 threshold, class-agnostic
    NMS at 0.2 -- with the oracle's NMS in place of the kernel (CPU)"""
    _mist_check(golden, "cpu", nms_fn=lambda b, s, g, t: oracle.batched_nms(b, s, g, t, oracle.IOU_TV_CPU))


@pytest.mark.gpu
def test_get_pgt_mist_vs_reference_on_gpu(golden):
    from wsovod_b200 import ops
    _mist_check(golden, "cuda:0", iou_mode=ops.IOU_TV_CPU)
