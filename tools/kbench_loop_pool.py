"""ROILoopPool timing (values only): block-max path (three floor-0 pooling passes + fix-up) vs the scan kernel.
Usage: python tools/kbench_loop_pool.py [c1 c2 c5 ...]"""
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
        nbytes = 3 * N * R * C * 49 * 4 + feat.numel() * 4
        res = {"config": name, "bytes": nbytes}
        outs = {}
        for tag, env in (("blockmax", _lib.POOL_BLOCKMAX), ("scan", _lib.POOL_SCAN)):
            _lib.tune(_lib.TUNE_POOL_PATH, env)
            ms = timeit(lambda: ops.roi_loop_pool(feat, rois, 1 / 8, 7, obj, 1.0, False), iters=8, flush=flush)
            outs[tag] = ops.roi_loop_pool(feat, rois, 1 / 8, 7, obj, 1.0, False)[0]
            res[f"{tag}_ms"] = round(ms, 4)
            res[f"{tag}_GBs"] = round(nbytes / ms / 1e6, 1)
        _lib.tune(_lib.TUNE_POOL_PATH, _lib.POOL_AUTO)
        res["equal"] = bool(torch.equal(outs["blockmax"], outs["scan"]))
        del outs
        ms = timeit(lambda: ops.roi_loop_pool(feat, rois, 1 / 8, 7, obj, 1.0, False), iters=8, flush=flush)
        res["default_ms"] = round(ms, 4)
        print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
