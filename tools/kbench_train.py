"""Training-slice timing (BASELINE configs[2] "c3" / configs[4] "c5"): the in-scope kernels of one
WSOVOD training step on one GPU -- values-only ROI pool (+objectness scale; the backbone is frozen, SURVEY
fact 6), alignment forward + backward (fp32 path, as autograd runs it), MIL two-stream forward + backward,
seed selection + pseudo-label assignment.  Usage: python tools/kbench_train.py [c3 c5]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wsovod_b200 import ops, synth  # noqa: E402
from tools.kbench import timeit  # noqa: E402

DEV = "cuda:0"


def main():
    names = [a for a in sys.argv[1:] if not a.startswith("-")] or ["c3"]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    for name in names:
        w = synth.workload(name)
        N, C, H, W, R, K, D = (w[k] for k in "NCHWRKD")
        M = N * R
        feat, rois, obj = w["features"].to(DEV), w["rois"].to(DEV), w["objectness"].to(DEV)
        x, t = w["region_emb"].to(DEV), w["text_emb"].to(DEV)
        off = torch.tensor(w["offsets"], device=DEV)
        boxes = rois[:, 1:].contiguous()
        res = {"config": name, "M": M, "C": C, "map": [H, W]}
        ms = timeit(lambda: ops.roi_pool(feat, rois, 1 / 8, 7, obj, 1.0, False), flush=flush)
        res["roi_pool_ms"] = round(ms, 4)
        res["roi_pool_GBs"] = round((M * C * 49 * 4 + feat.numel() * 4) / ms / 1e6, 1)

        xg = x.clone().requires_grad_(True)
        tg = t.clone().requires_grad_(True)

        def align_fb():
            lg, _ = ops.align(xg, tg, 50.0, True, True, None, ops.ALIGN_FP32, True, False)
            lg.backward(torch.ones_like(lg))
            xg.grad = None
            tg.grad = None
        res["align_fp32_fwd_bwd_ms"] = round(timeit(align_fb, flush=flush), 4)
        res["align_tf32_fwd_ms"] = round(timeit(lambda: ops.align(x, t, 50.0, True, True, None, ops.ALIGN_TF32, True, True), flush=flush), 4)

        g = synth.gen(7)
        Cl, Dl = synth.mil_logits(M, K, g)
        Cl, Dl = Cl.to(DEV).requires_grad_(True), Dl.to(DEV).requires_grad_(True)

        def mil_fb():
            s, img = ops.mil(Cl, Dl, off)
            (s.sum() + img.sum()).backward()
            Cl.grad = None
            Dl.grad = None
        res["mil_fwd_bwd_ms"] = round(timeit(mil_fb, flush=flush), 4)
        with torch.no_grad():
            s_mil, img = ops.mil(Cl, Dl, off)
        gts = synth.image_labels(N, K, g, 8)
        goff = [0]
        for gt in gts:
            goff.append(goff[-1] + len(gt))
        gtc, goffd = torch.cat(gts).to(DEV), torch.tensor(goff, device=DEV)

        def refine():
            sd = ops.pgt_top1(s_mil, boxes, off, gtc, goffd, img)
            return ops.refine_assign(boxes, off, sd["seed_boxes"], sd["seed_classes"], sd["seed_scores"],
                                     sd["seed_weights"], goffd, sd["seed_count"], K, 0.5)
        res["seeds_assign_ms"] = round(timeit(refine, flush=flush), 4)

        # weighted refinement losses on the labels just made (fused op) and the same in plain torch ops
        lab = refine()
        logits = (torch.randn(M, K + 1, generator=g) * 3).to(DEV).requires_grad_(True)
        deltas = (torch.randn(M, 4, generator=g) * 0.5).to(DEV).requires_grad_(True)

        def losses_fb():
            lc, lb = ops.refine_losses(logits, deltas, lab["gt_classes"], lab["gt_weights"], boxes, lab["gt_boxes"], K)
            (lc + lb).backward()
            logits.grad = None
            deltas.grad = None
        res["refine_losses_fwd_bwd_ms"] = round(timeit(losses_fb, flush=flush), 4)

        def losses_torch():
            import torch.nn.functional as F
            gc, w = lab["gt_classes"], lab["gt_weights"].clone()
            w[gc == -1] = 0.0
            lc = (F.cross_entropy(logits, gc, reduction="none", ignore_index=-1) * w).sum() / (w > 1e-12).float().sum()
            fg = torch.nonzero((gc >= 0) & (gc < K))[:, 0]
            src, tgt = boxes[fg], lab["gt_boxes"][fg]
            sw, sh = src[:, 2] - src[:, 0], src[:, 3] - src[:, 1]
            tw, th = tgt[:, 2] - tgt[:, 0], tgt[:, 3] - tgt[:, 1]
            tdel = torch.stack((10 * (tgt[:, 0] + 0.5 * tw - src[:, 0] - 0.5 * sw) / sw,
                                10 * (tgt[:, 1] + 0.5 * th - src[:, 1] - 0.5 * sh) / sh,
                                5 * torch.log(tw / sw), 5 * torch.log(th / sh)), 1)
            lb = ((deltas[fg] - tdel).abs() * w[fg, None]).sum() / max(M, 1)
            (lc + lb).backward()
            logits.grad = None
            deltas.grad = None
        res["refine_losses_torch_fwd_bwd_ms"] = round(timeit(losses_torch, flush=flush), 4)
        print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
