"""Soak of the CTA-pair alignment kernel: many launches on two streams at once and back to back, several shapes, every
result compared with the first one (tickets, cooperative launch, in-kernel finish).  python tools/soak_align_pair.py [iters]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wsovod_b200 import ops, synth  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 300
g = synth.gen(7)
bad = 0
for M, D, K in ((32000, 768, 1203), (5000, 256, 1023), (257, 64, 259), (40000, 32, 1203)):
    x = synth.region_embeddings(M, D, g).cuda()
    t = synth.text_embeddings(K, D, g).cuda()
    x2 = x.clone()
    ref = ops.align(x, t, 50.0, 1, True, None, ops.ALIGN_TF32, False, True)[1].clone()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    for it in range(iters):
        with torch.cuda.stream(s1):
            a = ops.align(x, t, 50.0, 1, True, None, ops.ALIGN_TF32, False, True)[1]
        with torch.cuda.stream(s2):
            b = ops.align(x2, t, 50.0, 1, True, None, ops.ALIGN_TF32, False, True)[1]
        if it % 10 == 9:
            torch.cuda.synchronize()
            if not (torch.equal(a, ref) and torch.equal(b, ref)):
                bad += 1
    torch.cuda.synchronize()
    print("shape", (M, D, K), "iterations", iters, "mismatching checks", bad, flush=True)
print("soak", "ok" if bad == 0 else "FAILED")
