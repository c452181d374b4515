"""One values-only ROILoopPool call at a named config, for ncu.  Usage: python tools/run_loop_pool_once.py c2 [scan]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wsovod_b200 import _lib, ops, synth  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
w = synth.workload(name)
DEV = "cuda:0"
feat, rois, obj = w["features"].to(DEV), w["rois"].to(DEV), w["objectness"].to(DEV)
if len(sys.argv) > 2 and sys.argv[2] == "scan":
    _lib.tune(_lib.TUNE_POOL_PATH, _lib.POOL_SCAN)
for _ in range(2):
    out = ops.roi_loop_pool(feat, rois, 1 / 8, 7, obj, 1.0, False)[0]
torch.cuda.synchronize()
print("ok", out.shape, float(out[::997].sum()))
