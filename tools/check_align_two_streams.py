import sys, torch
sys.path.insert(0, '/root/repo')
from wsovod_b200 import ops, synth
g = synth.gen(1)
x = synth.region_embeddings(32000, 768, g).cuda(); t = synth.text_embeddings(1203, 768, g).cuda()
x2 = x.clone()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
ref = ops.align(x, t, 50.0, 1, True, None, ops.ALIGN_TF32, False, True)[1].clone()
torch.cuda.synchronize()
outs = []
for it in range(30):
    with torch.cuda.stream(s1):
        a = ops.align(x, t, 50.0, 1, True, None, ops.ALIGN_TF32, False, True)[1]
    with torch.cuda.stream(s2):
        b = ops.align(x2, t, 50.0, 1, True, None, ops.ALIGN_TF32, False, True)[1]
    outs.append((a, b))
torch.cuda.synchronize()
ok = all(torch.equal(a, ref) and torch.equal(b, ref) for a, b in outs)
print("two streams x 30 concurrent pair-kernel launches:", "ok" if ok else "MISMATCH")
