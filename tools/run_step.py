"""The timed step of bench.py at a named BASELINE config, issued eagerly N times (no CUDA graph), for the ncu launch
list: python tools/run_step.py c2 5   (inference: pool -> alignment + softmax -> detections; c3 / c5: the training step
through WSOVODROIHeads)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wsovod_b200 import steps, synth  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = torch.device("cuda:0")
w = synth.workload(name, seed=1234, rank=0)
st = steps.make(name, w, dev, 1, graph=False)
for _ in range(reps):
    out = st()
torch.cuda.synchronize()
print("ok", name, reps)
