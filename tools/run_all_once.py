"""One launch of EVERY kernel family of libwsovod_b200.so at its benchmark shape, for ncu
(`ncu --set full -k regex:... python tools/run_all_once.py`): pooling flavours at c2, the contraction at c2 and c4
(K = 1203), the training kernels at c3's shapes (5024 proposals, K = 80), NMS tail at c2, generic batched NMS
(one group of 6000 boxes; RPN-like 5 x 2048), CSC.  Usage: python tools/run_all_once.py [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wsovod_b200 import _lib, ops, synth  # noqa: E402

DEV = "cuda:0"
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
w = synth.workload("c2")
feat, rois, obj = w["features"].to(DEV), w["rois"].to(DEV), w["objectness"].to(DEV)
x, t = w["region_emb"].to(DEV), w["text_emb"].to(DEV)
off = torch.tensor(w["offsets"], device=DEV)
sizes = w["image_sizes"].to(DEV)
boxes = rois[:, 1:].contiguous()
g = synth.gen(99)
t4 = synth.text_embeddings(1203, w["D"], g).to(DEV)
M3, K, D = 5024, 80, 768
Cl, Dl = (v.to(DEV) for v in synth.mil_logits(M3, K, g))
off3 = torch.tensor([0, M3], dtype=torch.int64, device=DEV)
b3 = synth.proposals(M3, 800, 1216, g).to(DEV)
gt = synth.image_labels(1, K, g, 8)[0].to(DEV)
goff = torch.tensor([0, gt.numel()], dtype=torch.int64, device=DEV)
x3, t3 = synth.region_embeddings(M3, D, g).to(DEV), synth.text_embeddings(K, D, g).to(DEV)
nb = synth.proposals(6000, 800, 1216, g).to(DEV)
ns = torch.rand(6000, generator=g).to(DEV)
rb = torch.cat([synth.proposals(2048, 800, 1216, g) for _ in range(5)]).to(DEV)
rs = torch.rand(5 * 2048, generator=g).to(DEV)
rg = torch.arange(5).repeat_interleave(2048).to(DEV)
cp = torch.rand(2, 20, 100, 152, generator=g).to(DEV)
cl = (torch.rand(2, 20, generator=g) < 0.5).float().to(DEV)
cr = torch.cat([torch.zeros(4000, 1), synth.proposals(4000, 100, 152, g)], 1).to(DEV)
for _ in range(reps):
    ops.roi_pool(feat, rois, 1 / 8, 7, obj, 1.0, False)                       # block-max kernel
    _lib.tune(_lib.TUNE_POOL_PATH, _lib.POOL_SCAN)
    ops.roi_pool(feat, rois, 1 / 8, 7, obj, 1.0, False)                       # the scan kernel it replaced
    _lib.tune(_lib.TUNE_POOL_PATH, _lib.POOL_AUTO)
    out, arg = ops.roi_pool(feat, rois, 1 / 8, 7, obj, 1.0, True)             # + argmax
    torch.ops.wsovod_b200.roi_pool_backward(out[:4000], rois[:4000], arg[:4000], 8, 512, 86, 128, False)
    del out, arg
    ops.roi_loop_pool(feat, rois, 1 / 8, 7, obj, 1.0, True)
    ra = ops.roi_align(feat, rois, 1 / 8, 7, 0, True, obj, 1.0)             # roi_align7_sep_kernel + roi_align_tables_kernel
    _lib.tune(_lib.TUNE_POOL_PATH, _lib.POOL_SCAN)
    ops.roi_align(feat, rois, 1 / 8, 7, 0, True, obj, 1.0)                  # the per-sample kernel it replaced (roi_plane_kernel)
    _lib.tune(_lib.TUNE_POOL_PATH, _lib.POOL_AUTO)
    f16, r16 = feat.half(), rois.half()
    ops.roi_loop_pool(f16, r16, 1 / 8, 7, None, 0.0, True)                  # loop_prepare_kernel<half> + loop_plane_kernel<half, 2>
    del f16, r16
    torch.ops.wsovod_b200.roi_align_backward(ra[:4000].contiguous(), rois[:4000].contiguous(), 1 / 8, 0, True, 8, 512, 86, 128)
    del ra
    _, probs = ops.align(x, t, 50.0, 1, True, None, ops.ALIGN_TF32, False, True)
    _, probs4 = ops.align(x, t4, 50.0, 1, True, None, ops.ALIGN_TF32, False, True)   # c4: K = 1203 (CTA-pair kernel)
    ops.detections(probs4, boxes, off, sizes, w["R"], 1e-5, 0.3, 100, ops.IOU_TV_CUDA)  # c4 tail: det_tau pruning threshold
    del probs4
    ops.detections(probs, boxes, off, sizes, w["R"], 1e-5, 0.3, 100, ops.IOU_TV_CUDA)
    ops.align(x3, t3, 50.0, 1, True, None, ops.ALIGN_FP32, True, False)
    gl = torch.randn(M3, K + 1, device=DEV)
    torch.ops.wsovod_b200.align_backward(gl, x3, t3, 50.0, 1, True, True, True)
    s, img = ops.mil(Cl, Dl, off3)
    torch.ops.wsovod_b200.mil_backward(torch.randn_like(s), torch.randn_like(img), Cl, Dl, off3)
    ops.align_mil(x3, t3, Dl, off3, 50.0, 1, None, True)
    sd = ops.pgt_top1(s, b3, off3, gt, goff, img)
    a = ops.refine_assign(b3, off3, sd["seed_boxes"], sd["seed_classes"], sd["seed_scores"], sd["seed_weights"], goff,
                          sd["seed_count"], K, 0.5)
    lg = torch.randn(M3, K + 1, device=DEV) * 3
    dl = torch.randn(M3, 4, device=DEV) * 0.1
    o, lse = torch.ops.wsovod_b200.refine_losses(lg, dl, a["gt_classes"], a["gt_weights"], b3, a["gt_boxes"], K, 10.0, 10.0, 5.0, 5.0, 0.0)
    torch.ops.wsovod_b200.refine_losses_backward(torch.ones(2, device=DEV), o, lse, lg, dl, a["gt_classes"], a["gt_weights"], b3,
                                                 a["gt_boxes"], K, 10.0, 10.0, 5.0, 5.0, 0.0)
    ops.batched_nms(nb, ns, torch.zeros(6000, dtype=torch.int64, device=DEV), 0.5)
    ops.batched_nms(rb, rs, rg, 0.7)
    ops.csc(cp, cl, torch.rand(2, 20, device=DEV), cr, 0.1, True, 1.8)
torch.cuda.synchronize()
print("ok")
