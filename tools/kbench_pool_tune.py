"""A/B timing of the block-max pooling kernel's tuning switches (wsovod_b200_tune) at a named config; every
variant is compared bit for bit with the scan kernels.  Usage: python tools/kbench_pool_tune.py [c2 c1 c5]"""
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
    iters = 15
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    for name in names:
        w = synth.workload(name)
        feat, rois, obj = w["features"].to(DEV), w["rois"].to(DEV), w["objectness"].to(DEV)
        nbytes = rois.size(0) * feat.size(1) * 49 * 4 + feat.numel() * 4 + rois.size(0) * 20
        _lib.tune(_lib.TUNE_POOL_PATH, _lib.POOL_SCAN)
        ref = ops.roi_pool(feat, rois, 1 / 8, 7, obj, 1.0, False)[0]
        _lib.tune(_lib.TUNE_POOL_PATH, _lib.POOL_BLOCKMAX)
        res = {"config": name, "algorithmic_bytes": nbytes}
        for group in (0, 1):
            _lib.tune(_lib.TUNE_POOL_GROUP, group)
            out = ops.roi_pool(feat, rois, 1 / 8, 7, obj, 1.0, False)[0]
            ms = timeit(lambda: ops.roi_pool(feat, rois, 1 / 8, 7, obj, 1.0, False), iters=iters, flush=flush)
            res[f"group{group}"] = {"ms": round(ms, 4), "GBs": round(nbytes / ms / 1e6, 1), "equal_scan": bool(torch.equal(out, ref))}
            del out
        _lib.tune(_lib.TUNE_POOL_GROUP, 1)
        _lib.tune(_lib.TUNE_POOL_PATH, _lib.POOL_AUTO)
        print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
