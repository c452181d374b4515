"""ROIAlign-only timing (7x7, adaptive grid, aligned): separable tap-table kernel vs the per-sample kernel, and
torchvision's CUDA op next to them.  Usage: python tools/kbench_roi_align.py [c1 c2 c5 ...]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wsovod_b200 import _lib, ops, synth  # noqa: E402
from tools.kbench import timeit  # noqa: E402

DEV = "cuda:0"


def main():
    names = [a for a in sys.argv[1:] if not a.startswith("-")] or ["c2"]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    for name in names:
        w = synth.workload(name)
        N, C, H, W, R = (w[k] for k in "NCHWR")
        feat, rois, obj = w["features"].to(DEV), w["rois"].to(DEV), w["objectness"].to(DEV)
        nbytes = N * R * C * 49 * 4 + feat.numel() * 4
        res = {"config": name, "bytes": nbytes}
        outs = {}
        for tag, env in (("separable", _lib.POOL_AUTO), ("per_sample", _lib.POOL_SCAN)):
            _lib.tune(_lib.TUNE_POOL_PATH, env)
            ms = timeit(lambda: ops.roi_align(feat, rois, 1 / 8, 7, 0, True, obj, 1.0), iters=8, flush=flush)
            outs[tag] = ops.roi_align(feat, rois, 1 / 8, 7, 0, True)
            res[f"{tag}_ms"] = round(ms, 4)
            res[f"{tag}_GBs"] = round(nbytes / ms / 1e6, 1)
        _lib.tune(_lib.TUNE_POOL_PATH, _lib.POOL_AUTO)
        a, b = outs["separable"], outs["per_sample"]
        res["max_abs_diff"] = float((a - b).abs().max())
        res["within_1e-5"] = bool(((a - b).abs() <= 1e-5 + 1e-5 * b.abs()).all())
        del outs, a, b
        try:
            import torchvision  # noqa: F401
            ms = timeit(lambda: torch.ops.torchvision.roi_align(feat, rois, 1 / 8, 7, 7, 0, True), iters=3, flush=flush)
            res["torchvision_cuda_ms"] = round(ms, 4)
        except Exception as e:   # noqa: BLE001
            res["torchvision_cuda_ms"] = repr(e)
        print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
