"""GPU parity tests: every CUDA kernel (called through the C ABI via wsovod_b200.ops) against the CPU
oracle on the same seeded inputs, against the reference-generated goldens in tests/golden, and --
where torchvision's CUDA ops are the reference's own GPU path -- against those.

Bars (north star): bit-exact for ROIPool values+argmax, refinement assignments and NMS keep-lists;
1e-5 relative for fp32 pooling (ROIAlign) and softmax; stated TF32 tolerance |dlogit| <= 5e-2 for the
tensor-core similarity logits.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

import oracle  # noqa: E402  (CPU checker)
from wsovod_b200 import ops, synth  # noqa: E402

DEV = "cuda:0"


def _offs(sizes):
    o = [0]
    for s in sizes:
        o.append(o[-1] + int(s))
    return o


# ------------------------------------------------------------------------------------------------ (1)
def _pool_case(N, C, H, W, R, seed, relu=False, wild=True):
    g = synth.gen(seed)
    feat = synth.features(N, C, H, W, g, relu=relu)
    boxes = [synth.proposals(R, H * 8, W * 8, g) for _ in range(N)]
    rois, _ = synth.rois_from(boxes)
    if wild:   # out-of-image, inverted and huge boxes
        k = max(R // 10, 1)
        rois[:k, 1:] += torch.randn(k, 4, generator=g) * 200
    perm = torch.randperm(rois.size(0), generator=g)     # arbitrary batch order
    return feat, rois[perm].contiguous()


@pytest.mark.parametrize("N,C,H,W,R", [(1, 4, 30, 40, 300), (3, 5, 23, 37, 200), (2, 1, 16, 16, 64),
                                         (2, 2, 60, 80, 500), (1, 7, 9, 11, 50), (2, 64, 60, 80, 700), (1, 6, 100, 152, 400), (2, 3, 100, 167, 300)])
def test_roi_pool_bit_exact(N, C, H, W, R):
    feat, rois = _pool_case(N, C, H, W, R, seed=N * 1000 + C)
    out, arg = ops.roi_pool(feat.to(DEV), rois.to(DEV), 1 / 8, 7, with_argmax=True)
    o_ref, a_ref = oracle.roi_pool(feat, rois, 1 / 8, 7)
    assert torch.equal(out.cpu(), o_ref)
    assert torch.equal(arg.cpu(), a_ref)
    out2, arg2 = ops.roi_pool(feat.to(DEV), rois.to(DEV), 1 / 8, 7, with_argmax=False)
    assert torch.equal(out2.cpu(), o_ref) and arg2.numel() == 0


def test_roi_pool_tiny_and_coinciding_bins():
    """proposals narrower/shorter than 7 cells (coinciding bins), 1-cell and sub-cell boxes, boxes on the
    map border: exercises the neighbour-column exchange of the fast kernel"""
    g = synth.gen(314)
    N, C, H, W = 2, 8, 40, 48
    feat = synth.features(N, C, H, W, g, relu=False)
    R = 600
    cx = torch.rand(R, generator=g) * W * 8
    cy = torch.rand(R, generator=g) * H * 8
    w = torch.rand(R, generator=g) ** 3 * 120 + 0.5           # many boxes of 0..4 cells
    h = torch.rand(R, generator=g) ** 3 * 120 + 0.5
    boxes = torch.stack([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2], 1)
    rois = torch.cat([torch.randint(0, N, (R, 1), generator=g).float(), boxes], 1)
    for with_arg in (True, False):
        out, arg = ops.roi_pool(feat.to(DEV), rois.to(DEV), 1 / 8, 7, with_argmax=with_arg)
        o_ref, a_ref = oracle.roi_pool(feat, rois, 1 / 8, 7)
        assert torch.equal(out.cpu(), o_ref)
        if with_arg:
            assert torch.equal(arg.cpu(), a_ref)
    # equal values everywhere: the first index in h-major order must win in every bin
    flat = torch.ones(1, 4, 30, 40)
    rois2, _ = synth.rois_from([synth.proposals(300, 240, 320, g)])
    out, arg = ops.roi_pool(flat.to(DEV), rois2.to(DEV), 1 / 8, 7, with_argmax=True)
    o_ref, a_ref = oracle.roi_pool(flat, rois2, 1 / 8, 7)
    assert torch.equal(out.cpu(), o_ref) and torch.equal(arg.cpu(), a_ref)


def test_roi_pool_vs_torchvision_cuda_and_golden(golden):
    import torchvision  # noqa: F401
    g = golden("pool")
    out, arg = ops.roi_pool(g["feat"].to(DEV), g["rois"].to(DEV), g["scale"], 7, with_argmax=True)
    assert torch.equal(out.cpu(), g["out"]) and torch.equal(arg.cpu(), g["argmax"])
    feat, rois = _pool_case(2, 32, 60, 80, 1000, seed=5)
    tv_out, tv_arg = torch.ops.torchvision.roi_pool(feat.to(DEV), rois.to(DEV), 1 / 8, 7, 7)
    out, arg = ops.roi_pool(feat.to(DEV), rois.to(DEV), 1 / 8, 7, with_argmax=True)
    assert torch.equal(out, tv_out) and torch.equal(arg, tv_arg.int())


def test_roi_pool_scale_epilogue_and_sizes():
    feat, rois = _pool_case(2, 6, 30, 40, 200, seed=9, relu=True, wild=False)
    obj = torch.rand(rois.size(0))
    out, _ = ops.roi_pool(feat.to(DEV), rois.to(DEV), 1 / 8, 7, row_scale=obj.to(DEV), row_scale_bias=1.0)
    ref, _ = oracle.roi_pool(feat, rois, 1 / 8, 7)
    ref = ref * (obj + 1).view(-1, 1, 1, 1)               # roi_heads.py:733-739
    assert torch.equal(out.cpu(), ref)
    for ps in [(3, 5), (14, 14), (1, 1)]:
        o, a = ops.roi_pool(feat.to(DEV), rois.to(DEV), 1 / 8, ps, with_argmax=True)
        r, ra = oracle.roi_pool(feat, rois, 1 / 8, ps)
        assert torch.equal(o.cpu(), r) and torch.equal(a.cpu(), ra)
    # empty inputs
    o, a = ops.roi_pool(feat.to(DEV), rois[:0].to(DEV), 1 / 8, 7, with_argmax=True)
    assert o.shape == (0, 6, 7, 7)


def test_roi_pool_large_map_fallback():
    # a plane that does not fit shared memory (H*W*4 > 227 KB) takes the global-memory variant
    g = synth.gen(3)
    feat = torch.randn(1, 2, 250, 260, generator=g)
    boxes = synth.proposals(150, 250 * 8, 260 * 8, g)
    rois, _ = synth.rois_from([boxes])
    o, a = ops.roi_pool(feat.to(DEV), rois.to(DEV), 1 / 8, 7, with_argmax=True)
    r, ra = oracle.roi_pool(feat, rois, 1 / 8, 7)
    assert torch.equal(o.cpu(), r) and torch.equal(a.cpu(), ra)


def test_roi_pool_backward():
    feat, rois = _pool_case(2, 3, 20, 24, 120, seed=11, wild=False)
    x = feat.to(DEV).requires_grad_(True)
    out, arg = ops.roi_pool(x, rois.to(DEV), 1 / 8, 7)
    assert arg.numel() == out.numel()
    go = torch.randn_like(out)
    out.backward(go)
    ref = oracle.roi_pool_backward(go.cpu(), rois, arg.cpu(), feat.shape)
    torch.testing.assert_close(x.grad.cpu(), ref, rtol=1e-5, atol=1e-5)   # atomics: order differs


@pytest.mark.parametrize("N,C,H,W,R", [(1, 4, 30, 40, 300), (2, 3, 23, 37, 200), (2, 1, 16, 16, 64), (1, 6, 100, 152, 300),
                                         (2, 16, 60, 80, 500)])
def test_roi_loop_pool_vs_oracle(N, C, H, W, R):
    feat, rois = _pool_case(N, C, H, W, R, seed=77 + C, relu=True, wild=False)
    out, arg = ops.roi_loop_pool(feat.to(DEV), rois.to(DEV), 1 / 8, 7)
    o_ref, a_ref = oracle.roi_loop_pool(feat, rois, 1 / 8, 7)
    assert out.shape == (3 * rois.size(0), C, 7, 7)
    assert torch.equal(out.cpu(), o_ref)
    assert torch.equal(arg.cpu(), a_ref)
    # value-only fast path (what a frozen backbone uses), with the objectness epilogue
    obj = torch.rand(rois.size(0))
    out2, arg2 = ops.roi_loop_pool(feat.to(DEV), rois.to(DEV), 1 / 8, 7, obj.to(DEV), 1.0, with_argmax=False)
    s3 = torch.cat([obj, obj, obj]) + 1
    assert arg2.numel() == 0 and torch.equal(out2.cpu(), o_ref * s3.view(-1, 1, 1, 1))


@pytest.mark.parametrize("sr,aligned", [(0, False), (0, True), (2, False), (2, True)])
def test_roi_align(sr, aligned, golden):
    g = golden("pool")
    out = ops.roi_align(g["feat"].to(DEV), g["rois_align"].to(DEV), g["scale"], 7, sr, aligned)
    torch.testing.assert_close(out.cpu(), g["align"][(sr, aligned)], rtol=1e-5, atol=1e-5)
    feat, rois = _pool_case(2, 5, 30, 40, 300, seed=21, wild=False)
    out = ops.roi_align(feat.to(DEV), rois.to(DEV), 1 / 8, 7, sr, aligned)
    ref = oracle.roi_align(feat, rois, 1 / 8, 7, sr, aligned)
    torch.testing.assert_close(out.cpu(), ref, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("N,C,H,W,R,aligned", [(2, 5, 30, 40, 300, True), (1, 3, 86, 128, 500, False),
                                               (1, 6, 100, 152, 400, True), (3, 2, 17, 9, 120, True), (1, 4, 1, 33, 50, True)])
def test_roi_align_separable_kernel(N, C, H, W, R, aligned):
    """7x7, adaptive grid: the separable tap-table kernel (roi_align_sep.cu; four channels per CTA, two on the 100x152
    map) against the oracle and against the per-sample kernel, wild boxes and the objectness scale included"""
    from wsovod_b200 import _lib
    feat, rois = _pool_case(N, C, H, W, R, seed=33 + H, wild=True)
    obj = synth.objectness(rois.size(0), synth.gen(5))
    fd, rd, od = feat.to(DEV), rois.to(DEV), obj.to(DEV)
    n0 = _lib.launch_count()
    out = ops.roi_align(fd, rd, 1 / 8, 7, 0, aligned)
    n_sep = _lib.launch_count() - n0
    ref = oracle.roi_align(feat, rois, 1 / 8, 7, 0, aligned)
    torch.testing.assert_close(out.cpu(), ref, rtol=1e-5, atol=1e-5)
    scaled = ops.roi_align(fd, rd, 1 / 8, 7, 0, aligned, od, 1.0)
    torch.testing.assert_close(scaled.cpu(), ref * (obj + 1).view(-1, 1, 1, 1), rtol=1e-5, atol=1e-5)
    old = _lib.tune(_lib.TUNE_POOL_PATH, _lib.POOL_SCAN)
    try:
        n0 = _lib.launch_count()
        scan = ops.roi_align(fd, rd, 1 / 8, 7, 0, aligned)
        n_scan = _lib.launch_count() - n0
    finally:
        _lib.tune(_lib.TUNE_POOL_PATH, old)
    torch.testing.assert_close(out, scan, rtol=1e-5, atol=1e-5)
    assert n_sep == n_scan + 1 or C == 1      # the tap-table prologue: proof that the separable kernel is what ran


# ------------------------------------------------------------------------------------------------ (2)
TF32_LOGIT_TOL = 5e-2      # stated tolerance: T*2*2^-11 worst case at T=50


@pytest.mark.parametrize("precision,tol", [(ops.ALIGN_FP32, 5e-5), (ops.ALIGN_TF32, TF32_LOGIT_TOL)])
def test_align_golden(precision, tol, golden):
    for name, c in golden("align").items():
        lg, pr = ops.align(c["x"].to(DEV), c["text"].to(DEV), c["T"], True, True, None, precision, True, True)
        assert (lg.cpu() - c["logits"]).abs().max().item() <= tol, name
        # softmax itself is fp32: compare against softmax of OUR logits at 1e-5 relative
        torch.testing.assert_close(pr, torch.softmax(lg, -1), rtol=1e-5, atol=1e-9)
        if precision == ops.ALIGN_FP32:
            torch.testing.assert_close(pr.cpu(), c["probs"], rtol=2e-4, atol=1e-7)
        lg2, _ = ops.align(c["x"].to(DEV), c["text"].to(DEV), c["T"], True, False, None, precision, True, False)
        assert (lg2.cpu() - c["logits_nobg"]).abs().max().item() <= tol, name


@pytest.mark.parametrize("M,D,K", [(1000, 768, 20), (300, 512, 80), (257, 768, 1203), (64, 64, 5), (129, 100, 33)])
@pytest.mark.parametrize("precision", [ops.ALIGN_FP32, ops.ALIGN_TF32])
def test_align_vs_oracle(M, D, K, precision):
    g = synth.gen(M + K)
    x = synth.region_embeddings(M, D, g)
    x[3] = 0
    t = synth.text_embeddings(K, D, g)
    bias = torch.tensor([0.25])
    lo, po = oracle.align(x, t, 50.0, True, True, bias=0.25)
    lg, pr = ops.align(x.to(DEV), t.to(DEV), 50.0, True, True, bias.to(DEV), precision, True, True)
    tol = 5e-5 if precision == ops.ALIGN_FP32 else TF32_LOGIT_TOL
    assert (lg.cpu() - lo).abs().max().item() <= tol
    torch.testing.assert_close(pr, torch.softmax(lg, -1), rtol=1e-5, atol=1e-9)
    # probs-only call (logits staged in the probs buffer)
    _, pr2 = ops.align(x.to(DEV), t.to(DEV), 50.0, True, True, bias.to(DEV), precision, False, True)
    torch.testing.assert_close(pr2, pr, rtol=1e-6, atol=1e-9)
    # no normalisation branch (:94 skipped when norm_weight is False)
    lo3, _ = oracle.align(x, t, 50.0, False, False, want_probs=False)
    lg3, _ = ops.align(x.to(DEV), t.to(DEV), 50.0, False, False, None, precision, True, False)
    scale = lo3.abs().max().item()
    assert (lg3.cpu() - lo3).abs().max().item() <= (1e-5 if precision == ops.ALIGN_FP32 else 4e-3) * scale


@pytest.mark.parametrize("M,D,K", [(257, 768, 1203), (1, 64, 300), (700, 100, 513), (2049, 768, 1203), (130, 32, 256),
                                   (5000, 512, 1000), (3000, 64, 1023), (777, 128, 2047), (300, 64, 2051), (515, 96, 259),
                                   (40000, 32, 1203)])
def test_align_cta_pair_kernel(M, D, K):
    """K + 1 > 256: the CTA-pair kernel (cta_group::2, units interleaved over the pairs, TMA-stored logits, in-kernel
    softmax finish behind tickets when the row pitch is a multiple of 16 bytes -- K + 1 = 1204, 1024, 2048, 260 --, plain
    stores and a separate softmax pass otherwise, no fusion above 2048 columns) against the oracle at the stated TF32
    tolerance and against the one-CTA-per-tile kernel (same TF32 products in the same order: identical logits;
    probabilities at 1e-5: the chunk statistics are merged in another order, the finish uses ex2.approx)"""
    from wsovod_b200 import _lib
    g = synth.gen(M + K + 1)
    x = synth.region_embeddings(M, D, g)
    x[0] = 0
    t = synth.text_embeddings(K, D, g)
    lo, po = oracle.align(x, t, 50.0, True, True)
    xd, td = x.to(DEV), t.to(DEV)
    assert _lib.tune(_lib.TUNE_ALIGN_PAIR, 1) == 1
    lg, pr = ops.align(xd, td, 50.0, True, True, None, ops.ALIGN_TF32, True, True)
    _, pr_only = ops.align(xd, td, 50.0, True, True, None, ops.ALIGN_TF32, False, True)
    lg_only, _ = ops.align(xd, td, 50.0, True, True, None, ops.ALIGN_TF32, True, False)
    _lib.tune(_lib.TUNE_ALIGN_PAIR, 0)
    try:
        lg1, pr1 = ops.align(xd, td, 50.0, True, True, None, ops.ALIGN_TF32, True, True)
    finally:
        _lib.tune(_lib.TUNE_ALIGN_PAIR, 1)
    assert (lg.cpu() - lo).abs().max().item() <= TF32_LOGIT_TOL
    assert torch.equal(lg, lg1) and torch.equal(lg_only, lg)
    torch.testing.assert_close(pr, torch.softmax(lg, -1), rtol=1e-5, atol=1e-9)
    torch.testing.assert_close(pr, pr1, rtol=2e-5, atol=1e-9)          # two paths, each within 1e-5 of softmax(logits)
    torch.testing.assert_close(pr_only, pr, rtol=1e-6, atol=1e-9)


def test_align_cta_pair_kernel_two_streams():
    """the pair kernel's finishing warps wait for tickets of other CTAs of the grid, so it is launched cooperatively
    (gang-scheduled): two instances issued on two streams at once must both complete, with identical results, instead of
    each holding half of the SMs and waiting for the other half forever"""
    g = synth.gen(4242)
    x = synth.region_embeddings(32000, 256, g).to(DEV)
    t = synth.text_embeddings(1203, 256, g).to(DEV)
    x2 = x.clone()
    ref = ops.align(x, t, 50.0, True, True, None, ops.ALIGN_TF32, False, True)[1].clone()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    outs = []
    for _ in range(10):
        with torch.cuda.stream(s1):
            a = ops.align(x, t, 50.0, True, True, None, ops.ALIGN_TF32, False, True)[1]
        with torch.cuda.stream(s2):
            b = ops.align(x2, t, 50.0, True, True, None, ops.ALIGN_TF32, False, True)[1]
        outs.append((a, b))
    torch.cuda.synchronize()
    assert all(torch.equal(a, ref) and torch.equal(b, ref) for a, b in outs)


def test_align_backward():
    g = synth.gen(5)
    x = synth.region_embeddings(200, 96, g).add_(0.01)
    t = synth.text_embeddings(20, 96, g)
    xg = x.to(DEV).requires_grad_(True)
    lg, _ = ops.align(xg, t.to(DEV), 50.0, True, True, None, ops.ALIGN_FP32, True, False)
    go = torch.randn_like(lg)
    lg.backward(go)
    xr = x.clone().requires_grad_(True)
    w = torch.nn.functional.normalize(t.t().contiguous(), p=2, dim=0)
    w = torch.cat([w, w.new_zeros(96, 1)], 1)
    ref = torch.mm(50.0 * torch.nn.functional.normalize(xr, p=2, dim=1), w)   # open_vocabulary_classifier.py:94-102
    ref.backward(go.cpu())
    torch.testing.assert_close(xg.grad.cpu(), xr.grad, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("M,D,K,mode,bg", [(200, 96, 20, 1, True), (5024, 768, 80, 1, True), (333, 100, 130, 2, False),
                                           (64, 64, 5, 0, True), (1000, 512, 1203, 1, True)])
def test_align_backward_classifier(M, D, K, mode, bg):
    """gradient w.r.t. the classifier (the "rand" Parameter of open_vocabulary_classifier.py:62-65) and w.r.t. x
    against autograd through the reference's own expression (:85-104), all three norm modes"""
    g = synth.gen(M + K)
    x = synth.region_embeddings(M, D, g).add_(0.01)
    x[3] = 0
    t = synth.text_embeddings(K, D, g)
    if mode == 2:
        t = torch.nn.functional.normalize(t, p=2, dim=1)
    xg, tg = x.to(DEV).requires_grad_(True), t.to(DEV).requires_grad_(True)
    lg, _ = ops.align(xg, tg, 50.0, mode, bg, None, ops.ALIGN_FP32, True, False)
    go = torch.randn(lg.shape, generator=g)
    lg.backward(go.to(DEV))
    xr, tr = x.double().requires_grad_(True), t.double().requires_grad_(True)
    w = tr.permute(1, 0)
    if mode == 1:
        w = torch.nn.functional.normalize(w, p=2, dim=0)
    xn = 50.0 * torch.nn.functional.normalize(xr, p=2, dim=1) if mode else xr
    if bg:
        w = torch.cat([w, w.new_zeros(D, 1)], 1)
    torch.mm(xn, w).backward(go.double())
    sw, sx = tr.grad.abs().max().item(), xr.grad.abs().max().item()
    assert (tg.grad.cpu().double() - tr.grad).abs().max().item() <= 2e-5 * sw
    assert (xg.grad.cpu().double() - xr.grad).abs().max().item() <= 1e-4 * sx
    # classifier-only request (x detached): the x half is skipped
    tg2 = t.to(DEV).requires_grad_(True)
    lg2, _ = ops.align(x.to(DEV), tg2, 50.0, mode, bg, None, ops.ALIGN_FP32, True, False)
    lg2.backward(go.to(DEV))
    assert torch.equal(tg2.grad, tg.grad)        # deterministic: fixed-order split-M sums


def test_align_mil_fused_golden_and_oracle(golden):
    """north star kernel 2, fused: against the reference's ObjectMiningOutputLayers(class_head=OpenVocabularyClassifier)
    golden and the oracle composition align -> mil, at the stated TF32 tolerance on the logits"""
    for name, c in golden("align_mil").items():
        off = torch.tensor(_offs(c["sizes"]), dtype=torch.int64, device=DEV)
        w = c["class_weight"].t().contiguous().to(DEV)                 # stored (D, K) -> (K, D)
        s, img, lg = ops.align_mil(c["x"].to(DEV), w, c["det"].to(DEV), off, c["T"], 2, None, True)
        assert (lg.cpu() - c["logits"]).abs().max().item() <= TF32_LOGIT_TOL, name
        # the score is compared against the two-stream formula on OUR logits at 1e-5 (fp32 softmaxes), and against
        # the reference's fp32 result at the error the logit tolerance allows (|d softmax| <= |d logit| / 2)
        so, io = oracle.mil(lg.cpu(), c["det"], _offs(c["sizes"]))
        torch.testing.assert_close(s.cpu(), so, rtol=1e-5, atol=1e-10)
        torch.testing.assert_close(img.cpu(), io, rtol=1e-5, atol=1e-8)
        assert (s.cpu() - c["scores"]).abs().max().item() <= 0.5 * TF32_LOGIT_TOL * c["scores"].max().item() + 1e-6, name
        assert (img.cpu() - c["img"]).abs().max().item() <= 0.5 * TF32_LOGIT_TOL, name
    g = synth.gen(31)
    for sizes, K, D in (([4000, 1, 2500, 0, 333], 80, 768), ([5024], 80, 768), ([130, 126, 1], 255, 64), ([7, 0, 0, 9], 3, 32)):
        M = sum(sizes)
        x = synth.region_embeddings(M, D, g)
        t = synth.text_embeddings(K, D, g)
        det = synth.mil_logits(M, K, g)[1]
        off = _offs(sizes)
        s, img, lg = ops.align_mil(x.to(DEV), t.to(DEV), det.to(DEV), torch.tensor(off, device=DEV), 50.0, 1, None, True)
        so, io, lo = oracle.align_mil(x, t, det, off, 50.0, 1)
        assert (lg.cpu() - lo).abs().max().item() <= TF32_LOGIT_TOL
        s2, i2 = oracle.mil(lg.cpu(), det, off)
        torch.testing.assert_close(s.cpu(), s2, rtol=1e-5, atol=1e-12)
        keep = [i for i, n in enumerate(sizes) if n > 0]
        torch.testing.assert_close(img.cpu()[keep], i2[keep], rtol=1e-5, atol=1e-8)
        empty = [i for i, n in enumerate(sizes) if n == 0]
        assert (img.cpu()[empty] == 1e-6).all()
        # equals the unfused pair of ops on the same logits, and is deterministic
        s3, i3 = ops.mil(lg, det.to(DEV), torch.tensor(off, device=DEV))
        torch.testing.assert_close(s, s3, rtol=1e-5, atol=1e-12)
        s4, i4, _ = ops.align_mil(x.to(DEV), t.to(DEV), det.to(DEV), torch.tensor(off, device=DEV), 50.0, 1, None, False)
        assert torch.equal(s, s4) and torch.equal(img, i4)


def test_align_mil_fused_backward():
    g = synth.gen(41)
    sizes, K, D = [200, 77], 20, 96
    M = sum(sizes)
    x = synth.region_embeddings(M, D, g).add_(0.01)
    t = synth.text_embeddings(K, D, g)
    det = synth.mil_logits(M, K, g)[1]
    off = torch.tensor(_offs(sizes), device=DEV)
    xg, tg, dg = (v.to(DEV).requires_grad_(True) for v in (x, t, det))
    s, img, _ = ops.align_mil(xg, tg, dg, off, 50.0, 1)
    gs, gi = torch.randn(s.shape, generator=g).to(DEV), torch.randn(img.shape, generator=g).to(DEV)
    (s * gs).sum().add((img * gi).sum()).backward()
    # reference expression in fp64 on the SAME (TF32) logits is not available to autograd: compare with the fp32
    # composition of the two ops (align fp32 -> mil), whose backward kernels are the ones this op reuses
    xr, tr, dr = (v.to(DEV).requires_grad_(True) for v in (x, t, det))
    lg, _ = ops.align(xr, tr, 50.0, 1, False, None, ops.ALIGN_FP32, True, False)
    s2, i2 = ops.mil(lg, dr, off)
    (s2 * gs).sum().add((i2 * gi).sum()).backward()
    for a, b in ((xg.grad, xr.grad), (tg.grad, tr.grad), (dg.grad, dr.grad)):
        scale = b.abs().max().item()
        assert (a - b).abs().max().item() <= 2e-2 * scale          # TF32 logits vs fp32 logits in the saved activations


def test_mil_golden_and_oracle(golden):
    for name, c in golden("mil").items():
        off = torch.tensor(_offs(c["sizes"]), dtype=torch.int64, device=DEV)
        s, img = ops.mil(c["cls"].to(DEV), c["det"].to(DEV), off)
        torch.testing.assert_close(s.cpu(), c["scores"], rtol=1e-5, atol=1e-10)
        torch.testing.assert_close(img.cpu(), c["img"], rtol=1e-5, atol=1e-8)
    g = synth.gen(8)
    sizes = [4000, 1, 2500, 0, 333]
    for K in (20, 80, 300):
        C, D = synth.mil_logits(sum(sizes), K, g)
        off = _offs(sizes)
        s, img = ops.mil(C.to(DEV), D.to(DEV), torch.tensor(off, device=DEV))
        so, io = oracle.mil(C, D, off)
        torch.testing.assert_close(s.cpu(), so, rtol=1e-5, atol=1e-12)
        keep = [i for i, n in enumerate(sizes) if n > 0]
        torch.testing.assert_close(img.cpu()[keep], io[keep], rtol=1e-5, atol=1e-8)


def test_mil_backward():
    g = synth.gen(12)
    sizes = [70, 33]
    C, D = synth.mil_logits(sum(sizes), 20, g)
    off = _offs(sizes)
    Cg, Dg = C.to(DEV).requires_grad_(True), D.to(DEV).requires_grad_(True)
    s, img = ops.mil(Cg, Dg, torch.tensor(off, device=DEV))
    gs, gi = torch.randn_like(s), torch.randn_like(img)
    (s * gs).sum().add((img * gi).sum()).backward()
    Cr, Dr = C.clone().requires_grad_(True), D.clone().requires_grad_(True)
    parts = [torch.softmax(c, 1) * torch.softmax(d, 0) for c, d in zip(Cr.split(sizes), Dr.split(sizes))]
    sr = torch.cat(parts)
    ir = torch.stack([p.sum(0) for p in parts]).clamp(1e-6, 1 - 1e-6)
    (sr * gs.cpu()).sum().add((ir * gi.cpu()).sum()).backward()
    torch.testing.assert_close(Cg.grad.cpu(), Cr.grad, rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(Dg.grad.cpu(), Dr.grad, rtol=1e-4, atol=1e-6)


# ------------------------------------------------------------------------------------------------ (3)
def _refine_run(boxes_l, scores_l, gt_l, img_scores, K):
    off = _offs([len(b) for b in boxes_l])
    goff = _offs([len(gc) for gc in gt_l])
    d = lambda t, dt=None: (t if dt is None else t.to(dt)).to(DEV)
    offd, goffd = d(torch.tensor(off)), d(torch.tensor(goff))
    seeds = ops.pgt_top1(d(torch.cat(scores_l)), d(torch.cat(boxes_l)), offd, d(torch.cat(gt_l)), goffd, d(img_scores))
    asg = ops.refine_assign(d(torch.cat(boxes_l)), offd, seeds["seed_boxes"], seeds["seed_classes"],
                            seeds["seed_scores"], seeds["seed_weights"], goffd, seeds["seed_count"], K, 0.5)
    so = oracle.pgt_top1(torch.cat(scores_l), torch.cat(boxes_l), off, torch.cat(gt_l), goff, img_scores)
    ao = oracle.refine_assign(torch.cat(boxes_l), off, so["seed_boxes"], so["seed_classes"], so["seed_scores"],
                              so["seed_weights"], goff, so["seed_count"], K, 0.5)
    return seeds, asg, so, ao, off, goff


def test_refine_golden(golden):
    f = golden("refine")
    seeds, asg, so, ao, off, goff = _refine_run(f["boxes"], f["scores"], f["gt_classes_img"], f["img_scores"],
                                                f["num_classes"])
    for n in range(len(f["sizes"])):
        g0, c = goff[n], int(seeds["seed_count"][n])
        assert c == len(f["seed_classes"][n])
        assert torch.equal(seeds["seed_boxes"][g0:g0 + c].cpu(), f["seed_boxes"][n])
        assert torch.equal(seeds["seed_classes"][g0:g0 + c].cpu(), f["seed_classes"][n])
        assert torch.equal(seeds["seed_scores"][g0:g0 + c].cpu(), f["seed_scores"][n])
        assert torch.equal(seeds["seed_weights"][g0:g0 + c].cpu(), f["seed_weights"][n])
        sl = slice(off[n], off[n + 1])
        assert torch.equal(asg["gt_classes"][sl].cpu(), f["gt_classes"][n])
        assert torch.equal(asg["gt_boxes"][sl].cpu(), f["gt_boxes"][n])
        assert torch.equal(asg["gt_scores"][sl].cpu(), f["gt_scores"][n])
        assert torch.equal(asg["gt_weights"][sl].cpu(), f["gt_weights"][n])


def test_refine_vs_oracle_full_size():
    g = synth.gen(31)
    K, sizes = 80, [5024, 4000, 1, 3000]
    boxes = [synth.proposals(s, 800, 1216, g) if s > 1 else torch.tensor([[10., 10., 200., 300.]]) for s in sizes]
    boxes[1] = (boxes[1] / 16).round() * 16            # integer grid: exact IoU ties and IoU == 0.5 cases
    scores = [torch.rand(s, K + 1, generator=g) for s in sizes]
    scores[0][:, 5] = 0.25                              # all-equal column: first index must win
    gts = synth.image_labels(len(sizes), K, g, max_labels=8)
    gts[0] = torch.unique(torch.cat([gts[0], torch.tensor([5])]))
    img = torch.rand(len(sizes), K, generator=g)
    seeds, asg, so, ao, _, _ = _refine_run(boxes, scores, gts, img, K)
    for k in so:
        assert torch.equal(seeds[k].cpu(), so[k]), k
    for k in ao:
        assert torch.equal(asg[k].cpu(), ao[k]), k
    assert (ao["gt_classes"] < K).sum() > 20           # the test actually has foreground


# ------------------------------------------------------------------------------------------ (3b) losses
def _loss_case(M, K, dcols, seed, fg_frac=0.3):
    g = synth.gen(seed)
    logits = torch.randn(M, K + 1, generator=g) * 3
    deltas = torch.randn(M, dcols, generator=g) * 0.5
    gt = torch.randint(0, K, (M,), generator=g)
    gt[torch.rand(M, generator=g) > fg_frac] = K                 # background
    gt[torch.rand(M, generator=g) < 0.1] = -1                    # ignored by the subsampling
    w = torch.rand(M, generator=g)
    w[torch.rand(M, generator=g) < 0.05] = 0.0                   # zero-weight rows do not count as valid
    pb = synth.proposals(M, 480, 640, g, stress=False)
    gb = synth.proposals(M, 480, 640, g, stress=False)
    return logits, deltas, gt, w, pb, gb


def _loss_run(logits, deltas, gt, w, pb, gb, K, beta, dev):
    lg = logits.clone().to(dev).requires_grad_()
    dl = None if deltas is None else deltas.clone().to(dev).requires_grad_()
    args = (lg, dl, gt.to(dev), w.to(dev), pb.to(dev), gb.to(dev), K)
    lc, lb = (ops.refine_losses(*args, smooth_l1_beta=beta) if dev != "cpu" else oracle.refine_losses(*args, beta=beta))
    (lc * 1.7 + lb * 0.6).backward()
    gd = None if dl is None else (torch.zeros_like(deltas) if dl.grad is None else dl.grad.cpu())
    return lc.detach().cpu(), lb.detach().cpu(), lg.grad.cpu(), gd


@pytest.mark.parametrize("M,K,specific,beta", [(5000, 80, False, 0.0), (3000, 20, True, 0.5), (777, 1203, False, 0.11),
                                                (64, 5, True, 0.0), (1, 3, False, 0.0)])
def test_refine_losses_vs_oracle(M, K, specific, beta):
    """fused weighted CE + weighted smooth-L1 (forward and backward) against the torch restatement of the
    reference; sums are accumulated in double on the GPU and pairwise in fp32 by torch: 1e-5 relative"""
    case = _loss_case(M, K, 4 * K if specific else 4, 1000 + M)
    a = _loss_run(*case, K, beta, DEV)
    b = _loss_run(*case, K, beta, "cpu")
    torch.testing.assert_close(a[0], b[0], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(a[1], b[1], rtol=1e-5, atol=1e-7)
    # d/d logits = g (softmax - onehot): the subtraction cancels where the target class dominates, so the error
    # is a few 1e-7 of the row's gradient scale (expf, fp32), not of the element
    torch.testing.assert_close(a[2], b[2], rtol=2e-5, atol=3e-7 * float(b[2].abs().max()) + 1e-12)
    torch.testing.assert_close(a[3], b[3], rtol=2e-5, atol=3e-7 * float(b[3].abs().max()) + 1e-12)


def test_refine_losses_golden(golden):
    """against the reference's own InstanceRefinementOutputLayers.losses (values and autograd gradients)"""
    for name, c in golden("refine_loss").items():
        lg = c["logits"].to(DEV).requires_grad_()
        dl = c["deltas"].to(DEV).requires_grad_() if c["reg"] else None
        lc, lb = ops.refine_losses(lg, dl, c["gt_classes"].to(DEV), c["gt_weights"].to(DEV), c["proposal_boxes"].to(DEV),
                                   c["gt_boxes"].to(DEV), c["num_classes"], smooth_l1_beta=c["beta"])
        torch.testing.assert_close(lc.detach().cpu(), c["loss_cls"], rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(lb.detach().cpu(), c["loss_box"], rtol=1e-5, atol=1e-8)
        (lc + lb).backward()
        torch.testing.assert_close(lg.grad.cpu(), c["grad_logits"], rtol=2e-5, atol=3e-7 * float(c["grad_logits"].abs().max()))
        if dl is not None:
            torch.testing.assert_close(dl.grad.cpu(), c["grad_deltas"], rtol=2e-5, atol=3e-7 * float(c["grad_deltas"].abs().max()))


def test_refine_losses_edge_cases():
    logits, deltas, gt, w, pb, gb = _loss_case(200, 7, 4, 5)
    # a NaN regression target (inverted gt box on a foreground row): the reference returns zeros(1) (:869-872)
    gt[3] = 2
    gb[3] = torch.tensor([50.0, 10.0, 20.0, 40.0])
    a = _loss_run(logits, deltas, gt, w, pb, gb, 7, 0.0, DEV)
    b = _loss_run(logits, deltas, gt, w, pb, gb, 7, 0.0, "cpu")
    assert float(a[1]) == 0.0 and float(b[1]) == 0.0 and float(a[3].abs().sum()) == 0.0
    torch.testing.assert_close(a[0], b[0], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(a[2], b[2], rtol=2e-5, atol=3e-7 * float(b[2].abs().max()))
    # every row ignored: 0 / 0 like the reference
    lc, lb = ops.refine_losses(logits.to(DEV), deltas.to(DEV), torch.full((200,), -1, device=DEV), w.to(DEV), pb.to(DEV),
                               gb.to(DEV), 7)
    assert torch.isnan(lc) and float(lb) == 0.0
    # no box regression (refine_reg off) and an empty batch
    lc, lb = ops.refine_losses(logits.to(DEV), None, gt.to(DEV), w.to(DEV))
    torch.testing.assert_close(lc.cpu(), b[0], rtol=1e-5, atol=1e-6)
    assert float(lb) == 0.0
    lc, lb = ops.refine_losses(torch.zeros(0, 8, device=DEV), torch.zeros(0, 4, device=DEV), torch.zeros(0, dtype=torch.int64, device=DEV),
                               torch.zeros(0, device=DEV), torch.zeros(0, 4, device=DEV), torch.zeros(0, 4, device=DEV), 7)
    assert torch.isnan(lc) and float(lb) == 0.0


# ------------------------------------------------------------------------------------------------ (4)
def _nms_case(M, ngroups, seed, grid=None):
    g = synth.gen(seed)
    b = synth.proposals(M, 480, 640, g)
    if grid:
        b = (b / grid).round() * grid
    s = torch.rand(M, generator=g)
    idx = torch.randint(0, ngroups, (M,), generator=g)
    return b, s, idx


@pytest.mark.parametrize("M,G,grid", [(3000, 20, None), (3000, 20, 4), (5000, 3, 16), (17, 4, None), (9000, 1, 8),
                                        (20000, 80, None)])
@pytest.mark.parametrize("mode", [ops.IOU_TV_CPU, ops.IOU_TV_CUDA])
def test_batched_nms_vs_oracle(M, G, grid, mode):
    b, s, idx = _nms_case(M, G, M + G, grid)
    keep = ops.batched_nms(b.to(DEV), s.to(DEV), (idx * 7 + 3).to(DEV), 0.3, mode)   # non-dense ids
    ref = oracle.batched_nms(b, s, idx, 0.3, mode)
    assert torch.equal(keep.cpu(), ref)


def test_batched_nms_vs_torchvision():
    from torchvision.ops.boxes import _batched_nms_vanilla
    for grid in (None, 4, 16):
        b, s, idx = _nms_case(4000, 20, 99, grid)
        # CPU arithmetic mode == torchvision CPU kernel
        keep = ops.batched_nms(b.to(DEV), s.to(DEV), idx.to(DEV), 0.3, ops.IOU_TV_CPU)
        assert torch.equal(keep.cpu(), _batched_nms_vanilla(b, s, idx, 0.3))
        # CUDA arithmetic mode == torchvision CUDA kernel (the reference's own GPU path)
        keep = ops.batched_nms(b.to(DEV), s.to(DEV), idx.to(DEV), 0.3, ops.IOU_TV_CUDA)
        tv = _batched_nms_vanilla(b.to(DEV), s.to(DEV), idx.to(DEV), 0.3)
        assert torch.equal(keep, tv)


def test_detections_golden(golden):
    d = golden("detections")
    off = _offs([len(b) for b in d["boxes"]])
    r = ops.detections(torch.cat(d["probs"]).to(DEV), torch.cat(d["boxes"]).to(DEV), torch.tensor(off, device=DEV),
                       torch.tensor(d["image_shapes"], dtype=torch.float32, device=DEV), max(len(b) for b in d["boxes"]),
                       d["score_thresh"], d["nms_thresh"], d["topk"], ops.IOU_TV_CPU)
    for n in range(len(d["boxes"])):
        c = int(r["det_count"][n])
        assert c == len(d["det_scores"][n])
        assert torch.equal(r["det_rows"][n, :c].cpu(), d["det_rows"][n])
        assert torch.equal(r["det_classes"][n, :c].cpu(), d["det_classes"][n])
        assert torch.equal(r["det_scores"][n, :c].cpu(), d["det_scores"][n])
        assert torch.equal(r["det_boxes"][n, :c].cpu(), d["det_boxes"][n])


@pytest.mark.parametrize("mode", [ops.IOU_TV_CPU, ops.IOU_TV_CUDA])
def test_detections_vs_oracle(mode):
    g = synth.gen(41)
    K, sizes = 20, [2000, 1500, 3]
    shapes = [(480, 640), (400, 600), (100, 100)]
    boxes = [synth.proposals(s, h, w, g) for s, (h, w) in zip(sizes, shapes)]
    boxes[1] = (boxes[1] / 8).round() * 8
    boxes[0][:20] += torch.randn(20, 4, generator=g) * 60
    probs = [torch.softmax(torch.randn(s, K + 1, generator=g) * 2.0, -1) for s in sizes]
    probs[0][7, 2] = float("inf")
    off = _offs(sizes)
    r = ops.detections(torch.cat(probs).to(DEV), torch.cat(boxes).to(DEV), torch.tensor(off, device=DEV),
                       torch.tensor(shapes, dtype=torch.float32, device=DEV), max(sizes), 1e-5, 0.3, 100, mode)
    o = oracle.detections(torch.cat(probs), torch.cat(boxes), off, shapes, 1e-5, 0.3, 100, mode)
    for k in o:
        assert torch.equal(r[k].cpu(), o[k]), k


@pytest.mark.parametrize("thr", [0.5, 0.25, 1.0 / 3.0, 0.2, 0.0, 1.0, 1e-7])
@pytest.mark.parametrize("mode", [ops.IOU_TV_CPU, ops.IOU_TV_CUDA])
def test_detections_threshold_ties(mode, thr):
    """Boxes snapped to a coarse grid: many pairs overlap by exactly 1/2, 1/3, 1/4, 1/5 ... -- quotients that
    sit on (or one rounding away from) the threshold, where the kernel's approximate-quotient screening must
    hand over to the exact IEEE division."""
    g = synth.gen(43)
    K, sizes = 6, [1800, 900]
    shapes = [(480, 640), (320, 480)]
    boxes = []
    for s, (h, w) in zip(sizes, shapes):
        x1 = torch.randint(0, w // 32 - 2, (s,), generator=g).float() * 32
        y1 = torch.randint(0, h // 32 - 2, (s,), generator=g).float() * 32
        bw = torch.randint(1, 5, (s,), generator=g).float() * 32
        bh = torch.randint(1, 5, (s,), generator=g).float() * 32
        boxes.append(torch.stack([x1, y1, x1 + bw, y1 + bh], 1))
    probs = [torch.softmax(torch.randn(s, K + 1, generator=g) * 2.0, -1) for s in sizes]
    off = _offs(sizes)
    r = ops.detections(torch.cat(probs).to(DEV), torch.cat(boxes).to(DEV), torch.tensor(off, device=DEV),
                       torch.tensor(shapes, dtype=torch.float32, device=DEV), max(sizes), 1e-5, thr, 100, mode)
    o = oracle.detections(torch.cat(probs), torch.cat(boxes), off, shapes, 1e-5, thr, 100, mode)
    for k in o:
        assert torch.equal(r[k].cpu(), o[k]), k


@pytest.mark.parametrize("K,topk,sizes", [(300, 7, [500, 0, 40]), (9, 100, [700, 300]), (1, 5, [64]), (33, 250, [900]),
                                          (257, 100, [300, 10]), (20, 1000, [1500]), (6, 3000, [3500]), (5, 3000, [3500, 200])])
def test_detections_topk_merge_shapes(K, topk, sizes):
    """The top-k stage splits an image's class runs over G = ceil(K / 8) <= 32 CTAs and the last one merges
    their lists: class counts that leave the last CTAs without runs (K = 257, 300), one class, fewer
    candidates than topk, an image without proposals, topk above the per-class run length, and topk so large
    that the run merge does not fit shared memory (packed list + selection / one sort instead)."""
    g = synth.gen(100 + K)
    shapes = [(480, 640)] * len(sizes)
    boxes = [synth.proposals(s, 480, 640, g) if s else torch.zeros(0, 4) for s in sizes]
    probs = [torch.softmax(torch.randn(s, K + 1, generator=g) * 3.0, -1) for s in sizes]
    off = _offs(sizes)
    r = ops.detections(torch.cat(probs).to(DEV), torch.cat(boxes).to(DEV), torch.tensor(off, device=DEV),
                       torch.tensor(shapes, dtype=torch.float32, device=DEV), max(sizes), 1e-4, 0.4, topk, ops.IOU_TV_CUDA)
    o = oracle.detections(torch.cat(probs), torch.cat(boxes), off, shapes, 1e-4, 0.4, topk, ops.IOU_TV_CUDA)
    for k in o:
        assert torch.equal(r[k].cpu(), o[k]), k


@pytest.mark.parametrize("K,topk,thr,quant", [(1203, 100, 1e-5, 0.0), (1203, 100, 1e-5, 2e-3), (400, 10, 1e-4, 1e-3),
                                             (150, 100, 0.02, 0.0), (1203, 100, 0.5, 0.0)])
def test_detections_image_pruning_threshold(K, topk, thr, quant):
    """K >= topk: candidates below tau = the topk-th largest per-class maximum of their image are dropped before the per-class
    NMS (det_scan / det_tau / det_compact kernels; exact: the best box of a class is never suppressed, so topk kept boxes at
    or above tau exist).
    LVIS-sized class counts, scores quantised so that many tie with tau, a score threshold that leaves fewer than topk
    non-empty classes (tau = 0), images of 5 and 0 proposals, row tiles that straddle two images -- bit-exact vs the oracle."""
    g = synth.gen(700 + K + topk)
    sizes = [700, 333, 0, 5]
    shapes = [(480, 640)] * len(sizes)
    boxes = [synth.proposals(s, 480, 640, g) if s else torch.zeros(0, 4) for s in sizes]
    probs = [torch.softmax(torch.randn(s, K + 1, generator=g) * 3.0, -1) for s in sizes]
    if quant:
        probs = [(p / quant).round() * quant for p in probs]
    else:
        # rows the finite filter drops (:178-182), among them the image's best score: the front end of the pruned path
        # accumulates class maxima while it checks a row and must take a dropped row's contribution back
        best = int(probs[0][:, :K].max(1).values.argmax())
        probs[0][best, K // 2] = float("nan")
        probs[0][13, 3] = float("inf")
        boxes[1][40, 2] = float("inf")
        probs[1][int(probs[1][:, :K].max(1).values.argmax()), 0] = float("-inf")
    off = _offs(sizes)
    r = ops.detections(torch.cat(probs).to(DEV), torch.cat(boxes).to(DEV), torch.tensor(off, device=DEV),
                       torch.tensor(shapes, dtype=torch.float32, device=DEV), max(sizes), thr, 0.4, topk, ops.IOU_TV_CUDA)
    o = oracle.detections(torch.cat(probs), torch.cat(boxes), off, shapes, thr, 0.4, topk, ops.IOU_TV_CUDA)
    for k in o:
        assert torch.equal(r[k].cpu(), o[k]), k


@pytest.mark.parametrize("mode", [ops.IOU_TV_CPU, ops.IOU_TV_CUDA])
def test_detections_prefix_fallback_and_long_columns(mode):
    """Columns longer than 2048 candidates take the histogram pre-selection; when the selected prefix
    runs dry before `topk` boxes are kept (here: the 2000 best-scoring boxes are near-duplicates) the
    kernel must fall back to the full column and still return the reference's detections."""
    g = synth.gen(77)
    K, R = 3, 6000
    shapes = [(800, 1200)]
    base = synth.proposals(R, 800, 1200, g, stress=False)
    boxes = base.clone()
    boxes[:2000] = torch.tensor([100., 100., 400., 500.]) + torch.rand(2000, 4, generator=g)   # one cluster
    probs = torch.rand(R, K + 1, generator=g) * 0.2
    probs[:2000, 0] = 0.7 + 0.3 * torch.rand(2000, generator=g)     # class 0: the cluster scores highest,
    probs[2000:, 0] = 0.5 + 0.2 * torch.rand(R - 2000, generator=g)  # ... the distinct boxes come after it
    probs[:, 1] = 0.4 * torch.rand(R, generator=g)                   # class 1: plain long column, lower scores
    probs[:, 2] = 1e-6                                               # class 2: nothing passes the threshold
    off = [0, R]
    r = ops.detections(probs.to(DEV), boxes.to(DEV), torch.tensor(off, device=DEV),
                       torch.tensor(shapes, dtype=torch.float32, device=DEV), R, 1e-5, 0.3, 100, mode)
    o = oracle.detections(probs, boxes, off, shapes, 1e-5, 0.3, 100, mode)
    for k in o:
        assert torch.equal(r[k].cpu(), o[k]), k
    # survivors of class 0 beyond the cluster only exist if the full column was processed
    assert int(o["det_count"][0]) == 100 and (o["det_classes"][0] == 0).sum() > 50


# ------------------------------------------------------------------------------ (1) block-max fast path
import contextlib  # noqa: E402
import os  # noqa: E402


@contextlib.contextmanager
def _pool_variant(scan=None):
    """select the values-only pooling kernel for the calls inside: scan=True the plain scan kernels,
    scan=False the block-max path wherever it applies (the library's own choice needs >= 1200 proposals
    per image)"""
    from wsovod_b200 import _lib
    old = _lib.tune(_lib.TUNE_POOL_PATH, _lib.POOL_AUTO if scan is None else _lib.POOL_SCAN if scan else _lib.POOL_BLOCKMAX)
    try:
        yield
    finally:
        _lib.tune(_lib.TUNE_POOL_PATH, old)


@pytest.mark.parametrize("N,C,H,W,R,seed", [(2, 9, 60, 80, 900, 1), (1, 3, 86, 128, 1500, 2), (3, 6, 100, 152, 700, 3),
                                             (1, 2, 150, 180, 400, 4), (2, 1, 7, 5, 300, 5), (1, 4, 1, 1, 50, 6),
                                             (1, 5, 117, 120, 300, 7), (2, 8, 60, 80, 1200, 8), (1, 12, 86, 128, 2500, 9),
                                             (1, 8, 40, 56, 3300, 10)])   # last: few channel groups -> proposals split over CTAs
def test_roi_pool_blockmax_path(N, C, H, W, R, seed):
    """values-only 7x7 pooling through the block-max planes (roi_pool_pyr.cu): every (kh, kw) phase, the
    direct-scan fallback phase (whole-map proposals on big maps), border-clipped bins, NaN / -inf cells,
    proposals in arbitrary batch order -- bit-exact against the oracle and against the scan kernels."""
    g = synth.gen(900 + seed)
    feat = synth.features(N, C, H, W, g, relu=False)
    flat = feat.view(-1)
    idx = torch.randint(0, flat.numel(), (max(flat.numel() // 50, 1),), generator=g)
    flat[idx[::2]] = float("nan")
    flat[idx[1::2]] = float("-inf")
    boxes = []
    for _ in range(N):
        b = synth.proposals(R, H * 8, W * 8, g)
        k = R // 8
        b[:k, :2] = torch.rand(k, 2, generator=g) * 40 - 20        # whole-map and beyond
        b[:k, 2] = W * 8 - torch.rand(k, generator=g) * 40 + 20
        b[:k, 3] = H * 8 - torch.rand(k, generator=g) * 40 + 20
        b[k:2 * k] += torch.randn(k, 4, generator=g) * 150          # out-of-image, inverted
        b[2 * k:3 * k] = (b[2 * k:3 * k] / 8).round() * 8           # integer cell grid
        boxes.append(b)
    rois, _ = synth.rois_from(boxes)
    rois = rois[torch.randperm(rois.size(0), generator=g)].contiguous()
    obj = torch.rand(rois.size(0), generator=g)
    ref, _ = oracle.roi_pool(feat, rois, 1 / 8, 7)
    fin = torch.where(torch.isfinite(ref), ref, torch.zeros_like(ref))
    sc = fin * (obj + 1).view(-1, 1, 1, 1)
    with _pool_variant(scan=True):
        scan = ops.roi_pool(feat.to(DEV), rois.to(DEV), 1 / 8, 7, with_argmax=False)[0]
    with _pool_variant(scan=False):
        out, arg = ops.roi_pool(feat.to(DEV), rois.to(DEV), 1 / 8, 7, with_argmax=False)
        out_s, _ = ops.roi_pool(feat.to(DEV), rois.to(DEV), 1 / 8, 7, row_scale=obj.to(DEV), row_scale_bias=1.0,
                                with_argmax=False)
    assert arg.numel() == 0
    assert torch.equal(out.cpu(), ref)
    assert torch.equal(scan, out)
    got = out_s.cpu()
    assert torch.equal(torch.where(torch.isfinite(ref), got, torch.zeros_like(got)), sc)
    # + argmax (planes of (value, index) pairs): the FIRST maximal cell of the row-major scan, -1 where nothing beats
    # -FLT_MAX; also on a map with few distinct values, where almost every bin has ties
    for f in (feat, torch.where(torch.isfinite(feat), (feat * 2).round() / 2, feat)):
        ref_v, ref_a = oracle.roi_pool(f, rois, 1 / 8, 7)
        with _pool_variant(scan=True):
            sv, sa = ops.roi_pool(f.to(DEV), rois.to(DEV), 1 / 8, 7, with_argmax=True)
        with _pool_variant(scan=False):
            bv, ba = ops.roi_pool(f.to(DEV), rois.to(DEV), 1 / 8, 7, with_argmax=True)
            bvs, bas = ops.roi_pool(f.to(DEV), rois.to(DEV), 1 / 8, 7, row_scale=obj.to(DEV), row_scale_bias=1.0,
                                    with_argmax=True)
        assert torch.equal(bv.cpu(), ref_v) and torch.equal(ba.cpu().to(ref_a.dtype), ref_a)
        assert torch.equal(sv, bv) and torch.equal(sa, ba) and torch.equal(bas, ba)
        fin_v = torch.where(torch.isfinite(ref_v), ref_v, torch.zeros_like(ref_v)) * (obj + 1).view(-1, 1, 1, 1)
        assert torch.equal(torch.where(torch.isfinite(ref_v), bvs.cpu(), torch.zeros_like(ref_v)), fin_v)


def test_roi_pool_blockmax_empty_images_and_single_class():
    """images without proposals, one proposal only, all proposals in one (kh, kw) class"""
    g = synth.gen(77)
    feat = synth.features(4, 8, 40, 56, g, relu=False)
    b = synth.proposals(200, 320, 448, g)
    rois = torch.cat([torch.full((200, 1), 2.0), b], 1)             # images 0, 1, 3 stay empty
    same = torch.tensor([[1.0, 64.0, 64.0, 64.0 + 8 * 27, 64.0 + 8 * 20]]).repeat(300, 1)   # 28 x 21 cells
    with _pool_variant(scan=False):
        for r in (rois, rois[:1].contiguous(), same):
            ref, _ = oracle.roi_pool(feat, r, 1 / 8, 7)
            out, _ = ops.roi_pool(feat.to(DEV), r.to(DEV), 1 / 8, 7, with_argmax=False)
            assert torch.equal(out.cpu(), ref)


@pytest.mark.parametrize("sr,aligned", [(0, False), (0, True), (2, True)])
def test_roi_align_backward_vs_torchvision(sr, aligned):
    """ROIAlign backward (round 2) against torchvision's compiled op on the same GPU, incl. the folded objectness scale"""
    import torchvision  # noqa: F401
    g = synth.gen(61)
    feat = synth.features(2, 5, 30, 40, g, relu=False)
    boxes = [synth.proposals(150, 240, 320, g, stress=False) for _ in range(2)]
    boxes[0][:10] += 90.0                                     # partly outside the map
    rois, _ = synth.rois_from(boxes)
    obj = synth.objectness(300, g)
    x1 = feat.to(DEV).requires_grad_(True)
    out = ops.roi_align(x1, rois.to(DEV), 1 / 8, 7, sr, aligned, obj.to(DEV), 1.0)
    go = torch.randn(out.shape, generator=g).to(DEV)
    out.backward(go)
    x2 = feat.to(DEV).requires_grad_(True)
    ref = torch.ops.torchvision.roi_align(x2, rois.to(DEV), 1 / 8, 7, 7, sr, aligned) * (obj.to(DEV) + 1).view(-1, 1, 1, 1)
    ref.backward(go)
    # torchvision's CUDA kernel contracts its bilinear sums into FMAs (ours rounds every operation, like torchvision's
    # CPU kernel, which the forward goldens pin at 1e-5): a few 1e-5 absolute on sums that cancel (features with negatives)
    torch.testing.assert_close(out, ref, rtol=1e-4, atol=1e-4)
    scale = x2.grad.abs().max().item()
    err = (x1.grad - x2.grad).abs().max().item()
    print(f"roi_align backward: max |d grad| {err:.3e} at gradient scale {scale:.3e}")
    assert err <= 5e-5 * scale


@pytest.mark.parametrize("N,C,H,W,R,seed", [(2, 9, 60, 80, 900, 1), (1, 3, 86, 128, 1500, 2), (3, 6, 100, 152, 700, 3),
                                             (2, 2, 7, 5, 300, 5), (1, 4, 1, 1, 50, 6), (1, 12, 86, 128, 2500, 9),
                                             (1, 8, 40, 56, 3300, 10)])
def test_roi_loop_pool_blockmax_path(N, C, H, W, R, seed):
    """values-only 3-way ROILoopPool through the block-max planes: three floor-0 pooling passes on the integer boxes of
    the ROI / outer grid + the fix-up kernel for the bins the excluded interior touches -- bit-exact against the oracle
    (ROILoopPool_cuda.cu restated on the CPU), against the scan kernel and, where it is built, the reference's own
    extension; negative, NaN and -inf cells (maxima start at 0), whole-map / out-of-image / inverted / integer-grid
    boxes, proposals in arbitrary batch order, the objectness scale."""
    from oracle import ref as oref
    g = synth.gen(1900 + seed)
    feat = synth.features(N, C, H, W, g, relu=False)
    flat = feat.view(-1)
    idx = torch.randint(0, flat.numel(), (max(flat.numel() // 50, 1),), generator=g)
    flat[idx[::2]] = float("nan")
    flat[idx[1::2]] = float("-inf")
    boxes = []
    for _ in range(N):
        b = synth.proposals(R, H * 8, W * 8, g)
        k = R // 8
        b[:k, :2] = torch.rand(k, 2, generator=g) * 40 - 20
        b[:k, 2] = W * 8 - torch.rand(k, generator=g) * 40 + 20
        b[:k, 3] = H * 8 - torch.rand(k, generator=g) * 40 + 20
        b[k:2 * k] += torch.randn(k, 4, generator=g) * 150
        b[2 * k:3 * k] = (b[2 * k:3 * k] / 8).round() * 8
        boxes.append(b)
    rois, _ = synth.rois_from(boxes)
    rois = rois[torch.randperm(rois.size(0), generator=g)].contiguous()
    obj = torch.rand(rois.size(0), generator=g)
    ref, _ = oracle.roi_loop_pool(feat, rois, 1 / 8, 7)
    fd, rd, od = feat.to(DEV), rois.to(DEV), obj.to(DEV)
    with _pool_variant(scan=True):
        scan = ops.roi_loop_pool(fd, rd, 1 / 8, 7, with_argmax=False)[0]
    from wsovod_b200 import _lib
    with _pool_variant(scan=False):
        n0 = _lib.launch_count()
        out, arg = ops.roi_loop_pool(fd, rd, 1 / 8, 7, with_argmax=False)
        launches = _lib.launch_count() - n0
        out_s, _ = ops.roi_loop_pool(fd, rd, 1 / 8, 7, od, 1.0, False)
    assert arg.numel() == 0
    assert torch.equal(out.cpu(), ref)
    assert torch.equal(scan, out)
    s3 = torch.cat([obj, obj, obj]) + 1
    assert torch.equal(out_s.cpu(), ref * s3.view(-1, 1, 1, 1))
    if C >= 2 and H * W * 8 <= 227 * 1024:
        assert launches >= 12            # 3 x (classify, order, bins, pool) + prologue + order + fix-up: the planes ran
    m = oref.cuda()
    if m is not None:
        assert torch.equal(out, m.roi_loop_pool_forward(fd, rd, 1 / 8, 7, 7)[0])
