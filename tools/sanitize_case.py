"""Small invocation of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool racecheck python tools/sanitize_case.py
The block-max pooling path is forced (TUNE_POOL_PATH = block-max) and checked against the scan kernels."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wsovod_b200 import _lib, ops, synth  # noqa: E402

DEV = "cuda:0"
g = synth.gen(3)
N, C, H, W, R, K, D = 2, 8, 24, 30, 160, 12, 64
feat = synth.features(N, C, H, W, g, relu=False).to(DEV)
boxes = [synth.proposals(R, H * 8, W * 8, g) for _ in range(N)]
boxes[0][:20, 2:] = torch.tensor([W * 8.0, H * 8.0])      # whole-map proposals: big blocks, clipped bins
rois, off = synth.rois_from(boxes)
rois = rois.to(DEV)
obj = synth.objectness(N * R, g).to(DEV)
_lib.tune(_lib.TUNE_POOL_PATH, _lib.POOL_BLOCKMAX)
a = ops.roi_pool(feat, rois, 1 / 8, 7, obj, 1.0, False)[0]
_lib.tune(_lib.TUNE_POOL_PATH, _lib.POOL_SCAN)
b = ops.roi_pool(feat, rois, 1 / 8, 7, obj, 1.0, False)[0]
_lib.tune(_lib.TUNE_POOL_PATH, _lib.POOL_AUTO)
assert torch.equal(a, b)
va, ia = ops.roi_pool(feat, rois, 1 / 8, 7, obj, 1.0, True)                  # scan kernel with argmax
_lib.tune(_lib.TUNE_POOL_PATH, _lib.POOL_BLOCKMAX)
vb, ib = ops.roi_pool(feat, rois, 1 / 8, 7, obj, 1.0, True)                  # (value, index) planes
_lib.tune(_lib.TUNE_POOL_PATH, _lib.POOL_AUTO)
assert torch.equal(va, vb) and torch.equal(ia, ib)
ra = ops.roi_align(feat, rois, 1 / 8, 7, 0, True, obj, 1.0)                  # separable tap tables (+ their prologue)
_lib.tune(_lib.TUNE_POOL_PATH, _lib.POOL_SCAN)
rb = ops.roi_align(feat, rois, 1 / 8, 7, 0, True, obj, 1.0)                  # per-sample kernel
_lib.tune(_lib.TUNE_POOL_PATH, _lib.POOL_AUTO)
assert torch.allclose(ra, rb, rtol=1e-5, atol=1e-5)
la = ops.roi_loop_pool(feat, rois, 1 / 8, 7, obj, 1.0, False)[0]              # converged scan kernel
_lib.tune(_lib.TUNE_POOL_PATH, _lib.POOL_BLOCKMAX)
lb = ops.roi_loop_pool(feat, rois, 1 / 8, 7, obj, 1.0, False)[0]              # floor-0 pooling passes + fix-up kernel (queues)
_lib.tune(_lib.TUNE_POOL_PATH, _lib.POOL_AUTO)
assert torch.equal(la, lb)
# half / double ROILoopPool (roi_loop_dtype.cu): staged planes (two channels, one channel) and the backward atomics
for dt in (torch.float16, torch.float64):
    o3, a3 = ops.roi_loop_pool(feat.to(dt), rois.to(dt), 1 / 8, 7, with_argmax=True)
    torch.ops.wsovod_b200.roi_pool_backward(torch.ones_like(o3), rois.to(dt), a3, N, C, H, W, True)
ops.roi_loop_pool(feat[:, :1].double().contiguous(), rois.double(), 1 / 8, 7, with_argmax=True)
x, t = synth.region_embeddings(N * R, D, g).to(DEV), synth.text_embeddings(K, D, g).to(DEV)
_, probs = ops.align(x, t, 50.0, True, True, None, ops.ALIGN_TF32, False, True)
ops.align(x, t, 50.0, True, True, None, ops.ALIGN_FP32, True, True)
# K + 1 > 256: the CTA-pair kernel (cluster barriers, multicast commits, TMA stores, tickets, in-kernel finish), K + 1 a
# multiple of 4 (TMA-stored logits) and not (plain stores + softmax pass); and the one-CTA multi-chunk kernel
for K3 in (299, 300):
    t3 = synth.text_embeddings(K3, D, g).to(DEV)
    l2, p2 = ops.align(x, t3, 50.0, True, True, None, ops.ALIGN_TF32, True, True)
    _lib.tune(_lib.TUNE_ALIGN_PAIR, 0)
    l1, p1 = ops.align(x, t3, 50.0, True, True, None, ops.ALIGN_TF32, True, True)
    _lib.tune(_lib.TUNE_ALIGN_PAIR, 1)
    assert torch.equal(l1, l2) and torch.allclose(p1, p2, rtol=2e-5, atol=1e-9)
offd = torch.tensor(off, device=DEV)
sizes = torch.tensor([[H * 8.0, W * 8.0]] * N, device=DEV)
bx = rois[:, 1:].contiguous()
det = ops.detections(probs, bx, offd, sizes, R, 1e-5, 0.3, 20, ops.IOU_TV_CUDA)
# a long column (histogram pre-selection, register sort, head stage + chunked continuation, two-level top-k merge)
R2, K2 = 900, 19
bx2 = synth.proposals(R2, 480, 640, g).to(DEV)
pr2 = torch.softmax(torch.randn(R2, K2 + 1, generator=g) * 2.0, -1).to(DEV)
det2 = ops.detections(pr2, bx2, torch.tensor([0, R2], device=DEV), torch.tensor([[480.0, 640.0]], device=DEV), R2, 1e-5,
                      0.3, 100, ops.IOU_TV_CPU)
# K >= topk: the pruned front end (scan / tau / compact, gather in det_class), rows dropped by the finite filter included
R3, K3 = 333, 150
bx3 = synth.proposals(R3, 480, 640, g).to(DEV)
pr3 = torch.softmax(torch.randn(R3, K3 + 1, generator=g) * 3.0, -1)
pr3[5, 7] = float("nan")
pr3[int(pr3[:, :K3].max(1).values.argmax()), 2] = float("inf")
det3 = ops.detections(pr3.to(DEV), bx3, torch.tensor([0, 200, R3], device=DEV), torch.tensor([[480.0, 640.0]] * 2, device=DEV), 200,
                      1e-5, 0.3, 10, ops.IOU_TV_CUDA)
Cl, Dl = synth.mil_logits(N * R, K, g)
s, img = ops.mil(Cl.to(DEV), Dl.to(DEV), offd)
gts = synth.image_labels(N, K, g)
goff = [0]
for gt in gts:
    goff.append(goff[-1] + len(gt))
sd = ops.pgt_top1(s, bx, offd, torch.cat(gts).to(DEV), torch.tensor(goff, device=DEV), img)
ops.refine_assign(bx, offd, sd["seed_boxes"], sd["seed_classes"], sd["seed_scores"], sd["seed_weights"],
                  torch.tensor(goff, device=DEV), sd["seed_count"], K, 0.5)
torch.cuda.synchronize()
print("sanitize case ok", int(det["det_count"].sum()), int(det2["det_count"].sum()), int(det3["det_count"].sum()))
