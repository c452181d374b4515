"""c4 contraction (32000 x 768 x 1204): CTA-pair kernel vs one CTA per tile.  python tools/kbench_align_pair.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wsovod_b200 import _lib, ops, synth  # noqa: E402

g = synth.gen(1)
M, D, K = 32000, 768, 1203
x = synth.region_embeddings(M, D, g).cuda()
t = synth.text_embeddings(K, D, g).cuda()


def timeit(f, n=20):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True)
    b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        f()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


f1 = lambda: ops.align(x, t, 50.0, 1, True, None, ops.ALIGN_TF32, False, True)
f2 = lambda: ops.align(x, t, 50.0, 1, True, None, ops.ALIGN_TF32, True, False)
for pair in (1, 0):
    _lib.tune(_lib.TUNE_ALIGN_PAIR, pair)
    print("pair", pair, "probs-only op ms", round(timeit(f1), 4), "logits-only ms", round(timeit(f2), 4))
_lib.tune(_lib.TUNE_ALIGN_PAIR, 1)
for K2 in (200, 255):
    t2 = synth.text_embeddings(K2, D, g).cuda()
    f3 = lambda: ops.align(x, t2, 50.0, 1, True, None, ops.ALIGN_TF32, False, True)
    print("K =", K2, "(one chunk, three-sweep epilogue) probs-only op ms", round(timeit(f3), 4))
