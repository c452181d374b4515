"""Per-kernel timing on one GPU (CUDA events, L2 flushed between iterations) with the torchvision /
ATen library path of the reference timed beside it.  Usage: python tools/kbench.py [c1|c2|...] [--tv]"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wsovod_b200 import ops, synth  # noqa: E402

DEV = "cuda:0"


def timeit(fn, iters=10, warm=3, flush=None):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    name = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else "c2"
    with_tv = "--tv" in sys.argv
    w = synth.workload(name)
    N, C, H, W, R, K, D = (w[k] for k in "NCHWRKD")
    M = N * R
    feat = w["features"].to(DEV)
    rois = w["rois"].to(DEV)
    obj = w["objectness"].to(DEV)
    x = w["region_emb"].to(DEV)
    t = w["text_emb"].to(DEV)
    off = torch.tensor(w["offsets"], device=DEV)
    sizes = w["image_sizes"].to(DEV)
    boxes = rois[:, 1:].contiguous()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    res = {"config": name, "M": M}
    out_bytes = M * C * 49 * 4
    ms = timeit(lambda: ops.roi_pool(feat, rois, 1 / 8, 7, obj, 1.0, True), flush=flush)
    res["roi_pool+argmax_ms"] = ms
    res["roi_pool+argmax_GBs"] = (2 * out_bytes + feat.numel() * 4) / ms / 1e6
    ms = timeit(lambda: ops.roi_pool(feat, rois, 1 / 8, 7, obj, 1.0, False), flush=flush)
    res["roi_pool_ms"] = ms
    res["roi_pool_GBs"] = (out_bytes + feat.numel() * 4) / ms / 1e6
    for prec, nm in ((ops.ALIGN_FP32, "fp32"), (ops.ALIGN_TF32, "tf32")):
        try:
            ms = timeit(lambda: ops.align(x, t, 50.0, True, True, None, prec, False, True), flush=flush)
            res[f"align_{nm}_ms"] = ms
            res[f"align_{nm}_TFLOPs"] = 2 * M * D * (K + 1) / ms / 1e9
            res[f"align_{nm}_GBs"] = (M * D * 4 + M * (K + 1) * 4) / ms / 1e6
        except Exception as e:  # noqa: BLE001
            res[f"align_{nm}_error"] = repr(e)[:200]
    _, probs = ops.align(x, t, 50.0, True, True, None, ops.ALIGN_FP32, False, True)
    ms = timeit(lambda: ops.detections(probs, boxes, off, sizes, R, 1e-5, 0.3, 100, ops.IOU_TV_CUDA), flush=flush)
    res["detections_ms"] = ms
    res["candidates"] = int((probs[:, :-1] > 1e-5).sum())
    g = synth.gen(7)
    Cl, Dl = synth.mil_logits(M, min(K, 80), g)
    Cl, Dl = Cl.to(DEV), Dl.to(DEV)
    ms = timeit(lambda: ops.mil(Cl, Dl, off), flush=flush)
    res["mil_ms"] = ms
    res["mil_GBs"] = 3 * Cl.numel() * 4 / ms / 1e6
    s_mil, img = ops.mil(Cl, Dl, off)
    gts = synth.image_labels(N, min(K, 80), g, 8)
    goff = [0]
    for gt in gts:
        goff.append(goff[-1] + len(gt))
    gtc, goffd = torch.cat(gts).to(DEV), torch.tensor(goff, device=DEV)

    def refine():
        sd = ops.pgt_top1(s_mil, boxes, off, gtc, goffd, img)
        return ops.refine_assign(boxes, off, sd["seed_boxes"], sd["seed_classes"], sd["seed_scores"],
                                 sd["seed_weights"], goffd, sd["seed_count"], min(K, 80), 0.5)
    res["refine_ms"] = timeit(refine, flush=flush)
    if with_tv:
        import torchvision
        ms = timeit(lambda: torch.ops.torchvision.roi_pool(feat, rois, 1 / 8, 7, 7), iters=5, flush=flush)
        res["tv_roi_pool_ms"] = ms

        def tv_align():
            xn = 50.0 * torch.nn.functional.normalize(x, p=2, dim=1)
            wn = torch.nn.functional.normalize(t.t().contiguous(), p=2, dim=0)
            wn = torch.cat([wn, wn.new_zeros(D, 1)], 1)
            return torch.softmax(torch.mm(xn, wn), -1)
        res["aten_align_ms"] = timeit(tv_align, iters=5, flush=flush)

        def tv_det():
            outs = []
            for n in range(N):
                p = probs[w["offsets"][n]:w["offsets"][n + 1], :-1]
                b = boxes[w["offsets"][n]:w["offsets"][n + 1]]
                m = p > 1e-5
                idx = m.nonzero()
                keep = torchvision.ops.boxes.batched_nms(b[idx[:, 0]], p[m], idx[:, 1], 0.3)[:100]
                outs.append(keep)
            return outs
        res["tv_detections_ms"] = timeit(tv_det, iters=3, warm=1, flush=flush)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
