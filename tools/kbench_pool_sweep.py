"""Crossover of the block-max path against the scan kernel over proposals/image (values-only 7x7 max-pool)."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wsovod_b200 import _lib, ops, synth  # noqa: E402
from tools.kbench import timeit  # noqa: E402
DEV = "cuda:0"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
for (N, C, H, W) in [(1, 512, 60, 80), (8, 512, 86, 128), (1, 512, 86, 128), (2, 2048, 50, 76)]:
    g = synth.gen(5)
    feat = synth.features(N, C, H, W, g).to(DEV)
    for R in (250, 500, 1000, 1500, 2000, 3000):
        rois, _ = synth.rois_from([synth.proposals(R, H * 8, W * 8, g) for _ in range(N)])
        rois = rois.to(DEV)
        res = {"N": N, "C": C, "HW": [H, W], "R": R}
        for tag, env in (("blockmax", _lib.POOL_BLOCKMAX), ("scan", _lib.POOL_SCAN)):
            _lib.tune(_lib.TUNE_POOL_PATH, env)
            res[tag + "_ms"] = round(timeit(lambda: ops.roi_pool(feat, rois, 1 / 8, 7, None, 0.0, False), iters=10, flush=flush), 4)
        print(json.dumps(res), flush=True)
