"""One invocation of each hot kernel at a named config (for ncu). Usage: python tools/run_once.py c2"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wsovod_b200 import _lib, ops, synth  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
w = synth.workload(name)
DEV = "cuda:0"
feat, rois, obj = w["features"].to(DEV), w["rois"].to(DEV), w["objectness"].to(DEV)
x, t = w["region_emb"].to(DEV), w["text_emb"].to(DEV)
off = torch.tensor(w["offsets"], device=DEV)
sizes = w["image_sizes"].to(DEV)
boxes = rois[:, 1:].contiguous()
for _ in range(reps):
    out, arg = ops.roi_pool(feat, rois, 1 / 8, 7, obj, 1.0, True)
    out2, _ = ops.roi_pool(feat, rois, 1 / 8, 7, obj, 1.0, False)          # block-max kernel (>= 3000 proposals/image)
    _lib.tune(_lib.TUNE_POOL_PATH, _lib.POOL_SCAN)
    out3, _ = ops.roi_pool(feat, rois, 1 / 8, 7, obj, 1.0, False)          # the scan kernel it replaced
    _lib.tune(_lib.TUNE_POOL_PATH, _lib.POOL_AUTO)
    _, probs = ops.align(x, t, 50.0, True, True, None, ops.ALIGN_TF32, False, True)
    det = ops.detections(probs, boxes, off, sizes, w["R"], 1e-5, 0.3, 100, ops.IOU_TV_CUDA)
torch.cuda.synchronize()
print("ok", out.shape, int(det["det_count"].sum()))
