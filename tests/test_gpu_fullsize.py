"""Full-size (BASELINE.json configs) GPU checks: where the CPU oracle would take minutes, compare against
the reference's own GPU library path (torchvision CUDA ops) and size-independent properties."""
import pytest
import torch

pytestmark = pytest.mark.gpu

import oracle  # noqa: E402
from wsovod_b200 import ops, synth  # noqa: E402

DEV = "cuda:0"


@pytest.fixture(scope="module")
def c2():
    w = synth.workload("c2")
    w["d"] = {k: w[k].to(DEV) for k in ("features", "rois", "objectness", "region_emb", "text_emb", "image_sizes")}
    return w


def test_c2_pool_equals_torchvision_cuda(c2):
    import torchvision  # noqa: F401
    d = c2["d"]
    out, arg = ops.roi_pool(d["features"], d["rois"], 1 / 8, 7, with_argmax=True)       # 3.2 GB + 3.2 GB
    tv_out, tv_arg = torch.ops.torchvision.roi_pool(d["features"], d["rois"], 1 / 8, 7, 7)
    assert torch.equal(out, tv_out)
    assert torch.equal(arg, tv_arg.int())
    del tv_out, tv_arg, arg
    out2, _ = ops.roi_pool(d["features"], d["rois"], 1 / 8, 7, with_argmax=False)
    assert torch.equal(out2, out)
    # linearity of the epilogue: scale by (objectness + 1) == separate multiply
    out3, _ = ops.roi_pool(d["features"], d["rois"], 1 / 8, 7, d["objectness"], 1.0, False)
    assert torch.equal(out3, out * (d["objectness"] + 1).view(-1, 1, 1, 1))
    # a checksum of checksums against a recomputation on a permuted proposal order
    perm = torch.randperm(d["rois"].size(0), device=DEV)
    out4, _ = ops.roi_pool(d["features"], d["rois"][perm], 1 / 8, 7, with_argmax=False)
    assert torch.equal(out4, out[perm])


def test_c2_roi_align_separable_kernel_full_size(c2):
    """ROIAlignV2 (aligned, adaptive grid) at c2 through the separable tap-table kernel: within 1e-5 of the per-sample
    kernel (torchvision's CPU operation order, the order the goldens pin) on all 0.8 G outputs; against torchvision's
    CUDA op within 1e-5 relative + 5e-5 absolute -- that kernel contracts `start + ph * bin` into an FMA, which moves a
    sample coordinate near 100 by one ulp (7.6e-6) and with it the bilinear weights, so torchvision's own two
    implementations are that far apart; the objectness epilogue is linear."""
    import torchvision  # noqa: F401
    from wsovod_b200 import _lib
    d = c2["d"]
    out = ops.roi_align(d["features"], d["rois"], 1 / 8, 7, 0, True)
    old = _lib.tune(_lib.TUNE_POOL_PATH, _lib.POOL_SCAN)
    try:
        ref = ops.roi_align(d["features"], d["rois"], 1 / 8, 7, 0, True)
    finally:
        _lib.tune(_lib.TUNE_POOL_PATH, old)
    err = (out - ref).abs_()
    assert bool((err <= ref.abs_().mul_(1e-5).add_(1e-5)).all())
    del ref, err
    tv = torch.ops.torchvision.roi_align(d["features"], d["rois"], 1 / 8, 7, 7, 0, True)
    err = (out - tv).abs_()
    assert bool((err <= tv.abs_().mul_(1e-5).add_(5e-5)).all())
    del tv, err
    out2 = ops.roi_align(d["features"], d["rois"], 1 / 8, 7, 0, True, d["objectness"], 1.0)
    assert torch.equal(out2, out * (d["objectness"] + 1).view(-1, 1, 1, 1))


def test_c2_roi_loop_pool_blockmax_equals_scan_kernel(c2):
    """3-way ROILoopPool at c2 (9.6 GB of output): the scan kernel the library picks against the block-max path (three
    floor-0 pooling passes + fix-up, forced through the tune switch), bit for bit, with the objectness scale folded in"""
    from wsovod_b200 import _lib
    d = c2["d"]
    out, _ = ops.roi_loop_pool(d["features"], d["rois"], 1 / 8, 7, d["objectness"], 1.0, False)
    old = _lib.tune(_lib.TUNE_POOL_PATH, _lib.POOL_BLOCKMAX)
    try:
        ref, _ = ops.roi_loop_pool(d["features"], d["rois"], 1 / 8, 7, d["objectness"], 1.0, False)
    finally:
        _lib.tune(_lib.TUNE_POOL_PATH, old)
    assert torch.equal(out, ref)
    R = d["rois"].size(0)
    plain, _ = ops.roi_pool(d["features"], d["rois"], 1 / 8, 7, d["objectness"], 1.0, False)
    assert torch.equal(out[:R], plain.clamp_(min=0))         # stream 0 is the max-pool with maxima starting at 0


def test_c2_alignment_and_detections_properties(c2):
    import torchvision
    d = c2["d"]
    lg_tc, pr_tc = ops.align(d["region_emb"], d["text_emb"], 50.0, 1, True, None, ops.ALIGN_TF32, True, True)
    lg_32, pr_32 = ops.align(d["region_emb"], d["text_emb"], 50.0, 1, True, None, ops.ALIGN_FP32, True, True)
    assert (lg_tc - lg_32).abs().max().item() <= 5e-2                 # stated TF32 tolerance
    torch.testing.assert_close(pr_tc.sum(1), torch.ones_like(pr_tc[:, 0]), rtol=0, atol=1e-5)
    torch.testing.assert_close(pr_tc, torch.softmax(lg_tc, -1), rtol=1e-5, atol=1e-9)
    assert torch.equal(lg_tc[:, -1], torch.zeros_like(lg_tc[:, -1]))   # background logit (zero weight column)
    boxes = d["rois"][:, 1:].contiguous()
    off = torch.tensor(c2["offsets"], device=DEV)
    r = ops.detections(pr_32, boxes, off, d["image_sizes"], c2["R"], 1e-5, 0.3, 100, ops.IOU_TV_CUDA)
    N, K = c2["N"], c2["K"]
    assert (r["det_count"] == 100).all()
    s = r["det_scores"]
    assert (s[:, 1:] <= s[:, :-1]).all() and (s > 1e-5).all()           # sorted, above the threshold
    for n in (0, N - 1):                                               # the reference's GPU path, per image
        a, b = c2["offsets"][n], c2["offsets"][n + 1]
        p = pr_32[a:b, :-1]
        m = p > 1e-5
        idx = m.nonzero()
        bx = boxes[a:b].clone()
        h, w = c2["image_sizes"][n].tolist()
        bx[:, 0::2].clamp_(0, w)
        bx[:, 1::2].clamp_(0, h)
        keep = torchvision.ops.boxes._batched_nms_vanilla(bx[idx[:, 0]], p[m], idx[:, 1], 0.3)[:100]
        assert torch.equal(r["det_rows"][n], idx[keep, 0])
        assert torch.equal(r["det_classes"][n], idx[keep, 1])
        assert torch.equal(r["det_scores"][n], p[m][keep])
        # idempotence: NMS over the survivors keeps all of them
        k2 = ops.batched_nms(r["det_boxes"][n], r["det_scores"][n], r["det_classes"][n], 0.3, ops.IOU_TV_CUDA)
        assert k2.numel() == 100


def test_c5_stress_refinement_and_nms_vs_oracle():
    """5000 proposals / image (config 5): refinement assignment and NMS against the CPU oracle."""
    g = synth.gen(55)
    K, R = 20, 5000
    boxes = synth.proposals(R, 800, 1216, g)
    probs = torch.softmax(torch.randn(R, K + 1, generator=g) * 2.0, -1)
    off = [0, R]
    r = ops.detections(probs.to(DEV), boxes.to(DEV), torch.tensor(off, device=DEV),
                       torch.tensor([[800., 1216.]], device=DEV), R, 1e-5, 0.3, 100, ops.IOU_TV_CPU)
    o = oracle.detections(probs, boxes, off, [(800, 1216)], 1e-5, 0.3, 100, oracle.IOU_TV_CPU)
    for k in o:
        assert torch.equal(r[k].cpu(), o[k]), k
    scores = torch.rand(R, K + 1, generator=g)
    gts = [torch.tensor([1, 4, 9, 17])]
    img = torch.rand(1, K, generator=g)
    d = lambda t: t.to(DEV)  # noqa: E731
    sd = ops.pgt_top1(d(scores), d(boxes), d(torch.tensor(off)), d(gts[0]), d(torch.tensor([0, 4])), d(img))
    a = ops.refine_assign(d(boxes), d(torch.tensor(off)), sd["seed_boxes"], sd["seed_classes"], sd["seed_scores"],
                          sd["seed_weights"], d(torch.tensor([0, 4])), sd["seed_count"], K, 0.5)
    so = oracle.pgt_top1(scores, boxes, off, gts[0], [0, 4], img)
    ao = oracle.refine_assign(boxes, off, so["seed_boxes"], so["seed_classes"], so["seed_scores"], so["seed_weights"],
                              [0, 4], so["seed_count"], K, 0.5)
    assert all(torch.equal(a[k].cpu(), ao[k]) for k in ao)


def test_c4_large_vocabulary_alignment():
    g = synth.gen(44)
    M, D, K = 4096, 768, 1203
    x, t = synth.region_embeddings(M, D, g).to(DEV), synth.text_embeddings(K, D, g).to(DEV)
    lg_tc, pr_tc = ops.align(x, t, 50.0, 1, True, None, ops.ALIGN_TF32, True, True)
    lg_32, _ = ops.align(x, t, 50.0, 1, True, None, ops.ALIGN_FP32, True, False)
    assert (lg_tc - lg_32).abs().max().item() <= 5e-2
    torch.testing.assert_close(pr_tc, torch.softmax(lg_tc, -1), rtol=1e-5, atol=1e-10)
    # top-1 agreement between TF32 and fp32 logits (rows whose two best fp32 logits differ by > the tolerance)
    top2 = lg_32.topk(2, dim=1).values
    clear = (top2[:, 0] - top2[:, 1]) > 0.1
    assert torch.equal(lg_tc.argmax(1)[clear], lg_32.argmax(1)[clear])
