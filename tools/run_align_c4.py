"""One launch of the c4 contraction (32000 x 768 x 1204, TF32) logits-only and one probs-only, for ncu.
python tools/run_align_c4.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wsovod_b200 import ops, synth  # noqa: E402

g = synth.gen(1)
x = synth.region_embeddings(32000, 768, g).cuda()
t = synth.text_embeddings(1203, 768, g).cuda()
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 1):
    ops.align(x, t, 50.0, 1, True, None, ops.ALIGN_TF32, True, False)
    ops.align(x, t, 50.0, 1, True, None, ops.ALIGN_TF32, False, True)
torch.cuda.synchronize()
