"""Generic batched_nms (all survivors returned) against torchvision's CUDA kernel on the same GPU.
    python tools/kbench_nms.py
Cases: an RPN-like selection (few groups, threshold 0.7, most boxes survive) and the inference shape of c1
(20 classes, threshold 0.3)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wsovod_b200 import ops, synth  # noqa: E402
from torchvision.ops.boxes import _batched_nms_vanilla  # noqa: E402

DEV = "cuda:0"


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    g = synth.gen(11)
    for name, (M, G, thr) in dict(rpn_like=(2048 * 5, 5, 0.7), c1_inference=(40000, 20, 0.3), one_group=(6000, 1, 0.5)).items():
        b = synth.proposals(M, 480, 640, g).to(DEV)
        s = torch.rand(M, generator=g).to(DEV)
        idx = torch.randint(0, G, (M,), generator=g).to(DEV)
        ours = ops.batched_nms(b, s, idx, thr, ops.IOU_TV_CUDA)
        tv = _batched_nms_vanilla(b, s, idx, thr)
        res = dict(case=name, M=M, groups=G, thr=thr, kept=int(ours.numel()), equal=bool(torch.equal(ours, tv)),
                   ours_ms=round(timeit(lambda: ops.batched_nms(b, s, idx, thr, ops.IOU_TV_CUDA)), 4),
                   torchvision_ms=round(timeit(lambda: _batched_nms_vanilla(b, s, idx, thr)), 4))
        print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
