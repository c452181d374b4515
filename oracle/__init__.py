"""CPU oracle for the WSOVOD region-scoring path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this package.  ``wsovod_b200`` (the product) never does: it has no CPU path at all.

The functions here wrap ``oracle/oracle.c`` (a plain-C restatement; each C function cites the
reference lines it follows) with CPU ``torch`` tensors in / out.  ``oracle.pins`` documents what the
restatement was checked against (torchvision's compiled CPU ops, ``oracle/_ref`` built from the
reference's own C++ sources, and the goldens in ``tests/golden`` produced by importing the reference
Python verbatim through ``oracle/d2_shim.py``).
"""
import ctypes
import os
import subprocess

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None


def build(force=False):
    """Compile oracle.c with gcc (seconds)."""
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "all"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.orc_batched_nms.restype = ctypes.c_int64
        _lib.orc_detections_image.restype = ctypes.c_int64
    return _lib


def _f(t):
    t = t.detach().to("cpu", torch.float32).contiguous()
    return t, ctypes.c_void_p(t.data_ptr())


def _i64(t):
    t = t.detach().to("cpu", torch.int64).contiguous()
    return t, ctypes.c_void_p(t.data_ptr())


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


_c64 = ctypes.c_int64
_cf = ctypes.c_float
_ci = ctypes.c_int


def roi_pool(input, rois, spatial_scale, output_size, with_argmax=True):
    ph, pw = (output_size, output_size) if isinstance(output_size, int) else output_size
    x, xp = _f(input)
    r, rp = _f(rois)
    N, C, H, W = x.shape
    R = r.shape[0]
    out = torch.empty(R, C, ph, pw, dtype=torch.float32)
    arg = torch.empty(R, C, ph, pw, dtype=torch.int32) if with_argmax else None
    lib().orc_roi_pool_fwd(xp, _c64(N), _c64(C), _c64(H), _c64(W), rp, _c64(R), _cf(spatial_scale),
                           _ci(ph), _ci(pw), _p(out), _p(arg))
    return out, arg


def roi_pool_backward(grad_out, rois, argmax, input_shape, num_rois=None):
    g, gp = _f(grad_out)
    r, rp = _f(rois)
    a = argmax.detach().to("cpu", torch.int32).contiguous()
    N, C, H, W = input_shape
    rows, _, ph, pw = g.shape
    R = r.shape[0] if num_rois is None else num_rois
    gi = torch.empty(N, C, H, W, dtype=torch.float32)
    lib().orc_roi_pool_bwd(gp, rp, _p(a), _c64(rows), _c64(R), _c64(N), _c64(C), _c64(H), _c64(W),
                           _ci(ph), _ci(pw), _p(gi))
    return gi


def roi_loop_pool(input, rois, spatial_scale, output_size):
    ph, pw = (output_size, output_size) if isinstance(output_size, int) else output_size
    x, xp = _f(input)
    r, rp = _f(rois)
    N, C, H, W = x.shape
    R = r.shape[0]
    out = torch.empty(3 * R, C, ph, pw, dtype=torch.float32)
    arg = torch.empty(3 * R, C, ph, pw, dtype=torch.int32)
    lib().orc_roi_loop_pool_fwd(xp, _c64(N), _c64(C), _c64(H), _c64(W), rp, _c64(R),
                                _cf(spatial_scale), _ci(ph), _ci(pw), _p(out), _p(arg))
    return out, arg


def roi_align(input, rois, spatial_scale, output_size, sampling_ratio=0, aligned=False):
    ph, pw = (output_size, output_size) if isinstance(output_size, int) else output_size
    x, xp = _f(input)
    r, rp = _f(rois)
    N, C, H, W = x.shape
    R = r.shape[0]
    out = torch.empty(R, C, ph, pw, dtype=torch.float32)
    lib().orc_roi_align_fwd(xp, _c64(N), _c64(C), _c64(H), _c64(W), rp, _c64(R), _cf(spatial_scale),
                            _ci(ph), _ci(pw), _ci(int(sampling_ratio)), _ci(int(bool(aligned))), _p(out))
    return out


def align(x, classifier, temperature=50.0, norm_weight=True, append_background=True, bias=None,
          want_probs=True):
    xx, xp = _f(x)
    cc, cp = _f(classifier)
    M, D = xx.shape
    K = cc.shape[0]
    KO = K + (1 if append_background else 0)
    logits = torch.empty(M, KO, dtype=torch.float32)
    probs = torch.empty(M, KO, dtype=torch.float32) if want_probs else None
    b = None if bias is None else torch.as_tensor([float(bias)], dtype=torch.float32)
    lib().orc_align_fwd(xp, cp, _c64(M), _c64(D), _c64(K), _cf(temperature), _ci(int(norm_weight)),
                        _ci(int(append_background)), _p(b), _p(logits), _p(probs))
    return logits, probs


def mil(cls, det, offsets):
    c, cp = _f(cls)
    d, dp = _f(det)
    o, op = _i64(torch.as_tensor(offsets))
    M, K = c.shape
    N = o.numel() - 1
    scores = torch.empty(M, K, dtype=torch.float32)
    img = torch.empty(N, K, dtype=torch.float32)
    lib().orc_mil_fwd(cp, dp, op, _c64(M), _c64(N), _c64(K), _p(scores), _p(img))
    return scores, img


def align_mil(x, classifier, det, offsets, temperature=50.0, norm_weight=True, bias=None):
    """ObjectMiningOutputLayers.forward with `cls` = the open-vocabulary class head
    (fast_rcnn_open_vocabulary.py:280-285,318-367; roi_heads.py:588-590): the alignment logits WITHOUT background
    column (open_vocabulary_classifier.py:79-105, append_background default False) feed the MIL two-stream score.
    Returns (scores, img, logits)."""
    logits, _ = align(x, classifier, temperature, norm_weight, False, bias, want_probs=False)
    scores, img = mil(logits, det, offsets)
    return scores, img, logits


def pgt_top1(scores, boxes, offsets, gt_classes, gt_offsets, img_scores):
    s, sp = _f(scores)
    b, bp = _f(boxes)
    o, op = _i64(torch.as_tensor(offsets))
    gc, gcp = _i64(torch.as_tensor(gt_classes))
    go, gop = _i64(torch.as_tensor(gt_offsets))
    im, imp = _f(img_scores)
    N = o.numel() - 1
    K = im.shape[1]
    G = gc.numel()
    sb = torch.zeros(G, 4)
    sc = torch.zeros(G, dtype=torch.int64)
    ss = torch.zeros(G)
    sw = torch.zeros(G)
    sr = torch.zeros(G, dtype=torch.int64)
    cnt = torch.zeros(N, dtype=torch.int64)
    lib().orc_pgt_top1(sp, _c64(s.shape[1]), bp, op, gcp, gop, imp, _c64(N), _c64(K), _p(sb), _p(sc),
                       _p(ss), _p(sw), _p(sr), _p(cnt))
    return dict(seed_boxes=sb, seed_classes=sc, seed_scores=ss, seed_weights=sw, seed_rows=sr,
                seed_count=cnt)


def refine_assign(boxes, offsets, seed_boxes, seed_classes, seed_scores, seed_weights, seed_offsets,
                  seed_count, num_classes, iou_thresh=0.5):
    b, bp = _f(boxes)
    o, op = _i64(torch.as_tensor(offsets))
    sb, sbp = _f(seed_boxes)
    sc, scp = _i64(seed_classes)
    ss, ssp = _f(seed_scores)
    sw, swp = _f(seed_weights)
    so, sop = _i64(torch.as_tensor(seed_offsets))
    cnt = None if seed_count is None else _i64(torch.as_tensor(seed_count))[0]
    M = b.shape[0]
    N = o.numel() - 1
    midx = torch.zeros(M, dtype=torch.int64)
    mlab = torch.zeros(M, dtype=torch.int8)
    miou = torch.zeros(M)
    gcls = torch.zeros(M, dtype=torch.int64)
    gbox = torch.zeros(M, 4)
    gsc = torch.zeros(M)
    gw = torch.zeros(M)
    lib().orc_refine_assign(bp, op, sbp, scp, ssp, swp, sop, _p(cnt), _c64(N), _c64(num_classes),
                            _cf(iou_thresh), _p(midx), _p(mlab), _p(miou), _p(gcls), _p(gbox),
                            _p(gsc), _p(gw))
    return dict(matched_idx=midx, matched_label=mlab, matched_iou=miou, gt_classes=gcls,
                gt_boxes=gbox, gt_scores=gsc, gt_weights=gw)


def refine_losses(logits, deltas, gt_classes, gt_weights, proposal_boxes, gt_boxes, num_classes,
                  box_weights=(10.0, 10.0, 5.0, 5.0), beta=0.0):
    """InstanceRefinementOutputLayers.losses, weighted flavours (fast_rcnn_open_vocabulary.py:754-892),
    restated with differentiable torch CPU ops (a floating-point path: torch fp32 is the reference
    arithmetic; pinned against the reference's own function in tests/golden/refine_loss.pt).
    Returns (loss_cls, loss_box_reg); deltas=None is refine_reg off."""
    import torch.nn.functional as F
    w = gt_weights.clone()
    w[gt_classes == -1] = 0.0                                                    # :791-792
    valid = (w > 1e-12).to(w.dtype).sum()                                        # :794-795
    ce = F.cross_entropy(logits, gt_classes, reduction="none", ignore_index=-1)  # :817
    loss_cls = (ce * w).sum() / valid                                            # :818-820
    if deltas is None:
        return loss_cls, torch.zeros(())
    fg = torch.nonzero((gt_classes >= 0) & (gt_classes < num_classes))[:, 0]     # :833
    if deltas.shape[1] == 4:
        fd = deltas[fg]                                                          # :834-835
    else:
        fd = deltas.view(-1, num_classes, 4)[fg, gt_classes[fg]]                 # :836-839
    src, tgt = proposal_boxes[fg], gt_boxes[fg]
    sw, sh = src[:, 2] - src[:, 0], src[:, 3] - src[:, 1]                         # d2 Box2BoxTransform.get_deltas
    scx, scy = src[:, 0] + 0.5 * sw, src[:, 1] + 0.5 * sh
    tw, th = tgt[:, 2] - tgt[:, 0], tgt[:, 3] - tgt[:, 1]
    tcx, tcy = tgt[:, 0] + 0.5 * tw, tgt[:, 1] + 0.5 * th
    wx, wy, ww, wh = box_weights
    target = torch.stack((wx * (tcx - scx) / sw, wy * (tcy - scy) / sh, ww * torch.log(tw / sw),
                          wh * torch.log(th / sh)), dim=1)
    if torch.isnan(target).any():                                                # :869-872
        return loss_cls, torch.zeros(())
    n = torch.abs(fd - target)                                                   # fvcore smooth_l1_loss
    l = n if beta < 1e-5 else torch.where(n < beta, 0.5 * n ** 2 / beta, n - 0.5 * beta)
    loss_box = (l * w[fg, None]).sum()                                           # :874-878
    return loss_cls, loss_box / max(gt_classes.numel(), 1.0)                     # :892


IOU_TV_CPU = 0
IOU_TV_CUDA = 1


def batched_nms(boxes, scores, groups, iou_thresh, iou_mode=IOU_TV_CPU):
    b, bp = _f(boxes)
    s, sp = _f(scores)
    g, gp = _i64(groups)
    M = b.shape[0]
    keep = torch.empty(max(M, 1), dtype=torch.int64)
    n = lib().orc_batched_nms(bp, sp, gp, _c64(M), ctypes.c_double(iou_thresh), _ci(iou_mode), _p(keep))
    return keep[:n].clone()


def detections(probs, boxes, offsets, image_sizes, score_thresh, nms_thresh, topk,
               iou_mode=IOU_TV_CPU):
    """fast_rcnn_inference for class-agnostic boxes; returns padded [N,topk,...] tensors + counts."""
    p, _ = _f(probs)
    b, _ = _f(boxes)
    offsets = [int(v) for v in offsets]
    N = len(offsets) - 1
    K = p.shape[1] - 1
    db = torch.zeros(N, topk, 4)
    ds = torch.zeros(N, topk)
    dc = torch.full((N, topk), -1, dtype=torch.int64)
    dr = torch.full((N, topk), -1, dtype=torch.int64)
    cnt = torch.zeros(N, dtype=torch.int64)
    for n in range(N):
        r0, r1 = offsets[n], offsets[n + 1]
        pn = p[r0:r1].contiguous()
        bn = b[r0:r1].contiguous()
        h, w = float(image_sizes[n][0]), float(image_sizes[n][1])
        k = lib().orc_detections_image(_p(pn), _p(bn), _c64(r1 - r0), _c64(K), _cf(h), _cf(w),
                                       _cf(score_thresh), ctypes.c_double(nms_thresh), _c64(topk), _ci(iou_mode),
                                       _p(db[n]), _p(ds[n]), _p(dc[n]), _p(dr[n]))
        cnt[n] = k
    return dict(det_boxes=db, det_scores=ds, det_classes=dc, det_rows=dr, det_count=cnt)


def csc(cpgs, labels, preds, rois, fg_threshold=0.1, area_sqrt=True, context_scale=1.8):
    """csc_forward restated in numpy, statement for statement (wsovod/layers/csc/csc_cuda.cu:183-531): the host loop
    over (image, class) with a positive label, `binary_and_integral_cpu` (:117-147), `CSCPool` with its int / float /
    double mix (:200-306), the max / min normalisation (:466-505) and the blend with `preds` (:506-509).  float32
    scalars are numpy float32 so every operation rounds where the reference's `T = float` rounds."""
    import numpy as np
    f32 = np.float32
    M = cpgs.detach().cpu().numpy().astype(np.float32)
    X = labels.detach().cpu().numpy().astype(np.float32)
    Y = preds.detach().cpu().numpy().astype(np.float32)
    Rr = rois.detach().cpu().numpy().astype(np.float32)
    B, K, H, Wd = M.shape
    R = Rr.shape[0]
    Wout = np.ones((R, K), np.float32)
    cs = f32(context_scale)

    def rnd(v):                       # C round(): half away from zero
        return np.where(v >= 0, np.floor(v + 0.5), np.ceil(v - 0.5))

    def box(I, ws, hs, we, he):
        a1 = I[he, we]
        a2 = np.where(ws - 1 >= 0, I[he, np.maximum(ws - 1, 0)], f32(0))
        a3 = np.where(hs - 1 >= 0, I[np.maximum(hs - 1, 0), we], f32(0))
        a4 = np.where((hs - 1 >= 0) & (ws - 1 >= 0), I[np.maximum(hs - 1, 0), np.maximum(ws - 1, 0)], f32(0))
        return ((a1 - a2).astype(f32) - a3).astype(f32) + a4

    for b in range(B):
        for c in range(K):
            if X[b, c] < 0.5:
                continue
            thr = f32(1.0) * f32(fg_threshold)
            binm = (M[b, c] >= thr).astype(np.float32)
            I = np.cumsum(np.cumsum(binm, axis=1, dtype=np.float32), axis=0, dtype=np.float32)     # exact integer counts
            ws = np.clip(rnd(Rr[:, 1]).astype(np.int64), 0, Wd - 1)
            hs = np.clip(rnd(Rr[:, 2]).astype(np.int64), 0, H - 1)
            we = np.clip(rnd(Rr[:, 3]).astype(np.int64), 0, Wd - 1)
            he = np.clip(rnd(Rr[:, 4]).astype(np.int64), 0, H - 1)
            wr, hr = (we - ws).astype(f32), (he - hs).astype(f32)
            wi, hi = (wr.astype(np.float64) / np.float64(cs)).astype(f32), (hr.astype(np.float64) / np.float64(cs)).astype(f32)
            wo, ho = (wr.astype(np.float64) * np.float64(cs)).astype(f32), (hr.astype(np.float64) * np.float64(cs)).astype(f32)
            wc, hc = ((we + ws) / 2.0).astype(f32), ((he + hs) / 2.0).astype(f32)
            d = np.float64
            ws_i = rnd(wc.astype(d) - wi.astype(d) / 2.0).astype(np.int64)
            hs_i = rnd(hc.astype(d) - hi.astype(d) / 2.0).astype(np.int64)
            we_i = np.minimum(rnd(wc.astype(d) + wi.astype(d) / 2.0).astype(np.int64), Wd - 1)
            he_i = np.minimum(rnd(hc.astype(d) + hi.astype(d) / 2.0).astype(np.int64), H - 1)
            ws_o = rnd(np.maximum(wc.astype(d) - wo.astype(d) / 2.0, 0.0)).astype(np.int64)
            hs_o = rnd(np.maximum(hc.astype(d) - ho.astype(d) / 2.0, 0.0)).astype(np.int64)
            we_o = rnd(np.minimum(wc.astype(d) + wo.astype(d) / 2.0, Wd - 1.0)).astype(np.int64)
            he_o = rnd(np.minimum(hc.astype(d) + ho.astype(d) / 2.0, H - 1.0)).astype(np.int64)
            a_roi = ((he - hs + 1).astype(f32) * (we - ws + 1).astype(f32)).astype(f32)
            a_in = ((he_i - hs_i + 1).astype(f32) * (we_i - ws_i + 1).astype(f32)).astype(f32)
            a_out = ((he_o - hs_o + 1).astype(f32) * (we_o - ws_o + 1).astype(f32)).astype(f32)
            s_roi, s_in, s_out = box(I, ws, hs, we, he), box(I, ws_i, hs_i, we_i, he_i), box(I, ws_o, hs_o, we_o, he_o)
            a_frame, a_ctx = np.maximum(a_roi - a_in, f32(1)), np.maximum(a_out - a_roi, f32(1))
            s_frame, s_ctx = (s_roi - s_in).astype(f32), (s_out - s_roi).astype(f32)
            if area_sqrt:
                score = (s_frame / np.sqrt(a_frame)).astype(f32) - (s_ctx / np.sqrt(a_ctx)).astype(f32)
            else:
                score = (s_frame / a_frame).astype(f32) - (s_ctx / a_ctx).astype(f32)
            score = score.astype(f32)
            mx = max(f32(0), score.max()) if R else f32(0)
            mn = min(f32(0), score.min()) if R else f32(0)
            if mx > 0 and mn < 0:
                v = np.where(score > 0, score / mx, score / (-mn)).astype(f32)
            elif mx > 0 and mn == 0:
                v = (score / mx).astype(f32)
            else:
                v = np.ones(R, f32)
            p = Y[b, c]
            Wout[:, c] = ((p * v).astype(f32) + ((f32(1) - p) * f32(1))).astype(f32)
    return torch.from_numpy(Wout)
