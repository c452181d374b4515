"""ROILoopPool (vs the reference's own CUDA extension, oracle/_ref) and ROIAlign (vs torchvision CUDA) at a
named config.  Usage: python tools/kbench_pool_variants.py c2"""
import json
import os
import sys

import torch
import torchvision  # noqa: F401  (registers torch.ops.torchvision.*)

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.kbench import timeit  # noqa: E402
from wsovod_b200 import ops, synth  # noqa: E402

DEV = "cuda:0"
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
w = synth.workload(name)
feat, rois = w["features"].to(DEV), w["rois"].to(DEV)
M, C = rois.size(0), feat.size(1)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
res = {"config": name, "M": M}
out_bytes = M * C * 49 * 4
res["roi_loop_pool_ms"] = timeit(lambda: ops.roi_loop_pool(feat, rois, 1 / 8, 7, with_argmax=True), iters=5, flush=flush)
res["roi_loop_pool_GBs"] = 6 * out_bytes / res["roi_loop_pool_ms"] / 1e6
res["roi_loop_pool_noarg_ms"] = timeit(lambda: ops.roi_loop_pool(feat, rois, 1 / 8, 7, with_argmax=False), iters=5, flush=flush)
try:
    from oracle import ref
    m = ref.cuda()
    if m is not None:
        res["reference_ext_roi_loop_pool_ms"] = timeit(lambda: m.roi_loop_pool_forward(feat, rois, 1 / 8, 7, 7), iters=3, warm=1, flush=flush)
except Exception as e:  # noqa: BLE001
    res["reference_ext_error"] = repr(e)[:200]
for aligned in (False, True):
    k = "roi_align_v2" if aligned else "roi_align"
    res[k + "_ms"] = timeit(lambda: ops.roi_align(feat, rois, 1 / 8, 7, 0, aligned), iters=5, flush=flush)
    res[k + "_GBs"] = out_bytes / res[k + "_ms"] / 1e6
    res["tv_" + k + "_ms"] = timeit(lambda: torch.ops.torchvision.roi_align(feat, rois, 1 / 8, 7, 7, 0, aligned), iters=3, warm=1, flush=flush)
print(json.dumps(res))
