import ctypes, os, sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from wsovod_b200 import _lib, ops, synth
g = synth.gen(1)
x = synth.region_embeddings(32000, 768, g).cuda(); t = synth.text_embeddings(1203, 768, g).cuda()
f = lambda: ops.align(x, t, 50.0, 1, True, None, ops.ALIGN_TF32, False, True)
for _ in range(3): f()
_lib.tune(15, 16 << 1)
f(); torch.cuda.synchronize()
_lib.tune(15, 0)
buf = np.zeros(148 * 512 * 4, dtype=np.uint64)
L = _lib.lib()
L.wsovod_b200_debug_trace.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
print("rc", L.wsovod_b200_debug_trace(buf.ctypes.data, buf.nbytes))
tr = buf[:148 * 512].reshape(148, 512)
ep = []; jobs = []
for cta in range(148):
    for i in range(8):
        r = tr[cta, i * 8:i * 8 + 8]
        if r[0] == 1: ep.append((cta, i, int(r[1]), int(r[2]), int(r[3]), int(r[4]), int(r[5]), int(r[6]), int(r[7])))
    for j in range(112):
        r = tr[cta, 64 + j * 4: 64 + j * 4 + 4]
        if (int(r[0]) & 255) == 2: jobs.append((cta, j, int(r[0]) >> 8, int(r[1]), int(r[2]), int(r[3])))
t0 = min(e[2] for e in ep)
print("epilogue units:", len(ep), "jobs:", len(jobs))
import collections
# per unit index: mean start / end (us)
for i in range(10):
    es = [e for e in ep if e[1] == i]
    if es: print(f"unit {i}: tfull at {np.mean([e[2]-t0 for e in es])/1e3:7.1f} us (min {min(e[2]-t0 for e in es)/1e3:6.1f} max {max(e[2]-t0 for e in es)/1e3:6.1f}), epilogue {np.mean([e[3]-e[2] for e in es])/1e3:5.2f} us")
print("epilogue cycles per unit: tmem loads %.0f, wait_group.read %.0f, slab loop %.0f, whole unit %.0f" % tuple(np.mean([e[k] for e in ep]) for k in (5, 6, 7, 8)))
print("last epilogue end", max(e[3] - t0 for e in ep) / 1e3)
w = np.array([(j[4] - j[3]) / 1e3 for j in jobs]); pr = np.array([(j[5] - j[4]) / 1e3 for j in jobs])
print("jobs: ticket wait mean %.1f us max %.1f; processing mean %.1f us min %.1f max %.1f" % (w.mean(), w.max(), pr.mean(), pr.min(), pr.max()))
print("last job end", max(j[5] - t0 for j in jobs) / 1e3)
for lo in range(0, 200, 20):
    sel = [j for j in jobs if lo <= (j[3] - t0) / 1e3 < lo + 20]
    if sel: print(f"jobs grabbed in [{lo},{lo+20}) us: {len(sel)}, wait {np.mean([(j[4]-j[3])/1e3 for j in sel]):.1f}, proc {np.mean([(j[5]-j[4])/1e3 for j in sel]):.1f}, by finisher warps {sum(1 for j in sel if j[2] >= 8)}")
c0 = [j for j in jobs if j[0] == 0]
for j in sorted(c0, key=lambda j: j[3]): print("cta0 job", j[1], "warp", j[2], "grab %.1f ticket %.1f done %.1f" % ((j[3]-t0)/1e3, (j[4]-t0)/1e3, (j[5]-t0)/1e3))
