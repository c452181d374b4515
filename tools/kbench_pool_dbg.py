"""Experiments: time the block-max pooling kernel with parts switched off (results are wrong by design)."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wsovod_b200 import ops, synth  # noqa: E402
from tools.kbench import timeit  # noqa: E402
DEV = "cuda:0"
w = synth.workload(sys.argv[1] if len(sys.argv) > 1 else "c2")
feat, rois, obj = w["features"].to(DEV), w["rois"].to(DEV), w["objectness"].to(DEV)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
for dbg in [0, 1, 2, 4, 3, 5, 6, 7]:
    os.environ["WSOVOD_B200_POOL_DEBUG"] = str(dbg)
    ms = timeit(lambda: ops.roi_pool(feat, rois, 1 / 8, 7, obj, 1.0, False), iters=10, flush=flush)
    print(json.dumps({"debug": dbg, "ms": round(ms, 4)}), flush=True)
