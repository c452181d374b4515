"""Bank-conflict simulation of the block-max pooling hot loop on the c2 descriptor stream (CPU only).

    python tools/bank_sim.py [proposals]

Replays the geometry of csrc/pool_pyr.cuh in numpy (classes, first-block cells, distances to the last block),
walks the lane slots of a proposal the way pyr_run does (two passes: bins 0..31 and 32..48; quarter-warps of 8
lanes, one LDS.128 per block position) and counts shared-memory wavefronts per quarter-warp load (cells in the
same 16-byte bank group, cell index mod 8, serialise; equal cells broadcast):

  * row-major lanes (round 1):                         1.70 wavefronts per quarter-warp load (ncu: 1.66)
  * the first-fit lane order of pyr_bins_kernel<GROUP>: 1.44  (ncu, c2: 72.8 M -> 40.3 M conflict wavefronts)
  * column rotations / XOR swizzles / other row pitches of the plane: within 2 % of the plain layout -- the
    lower bound over all lane orders is 1.07, so the layout is not what limits, the lane order is."""
import collections
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wsovod_b200 import synth  # noqa: E402

NR = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
w = synth.workload("c2")
rois = w["rois"].numpy()[:NR]
H, W = w["features"].shape[2:]
scale = np.float32(1 / 8)
RS = W + 3


def roundf(v):
    return np.sign(v) * np.floor(np.abs(v) + np.float32(0.5))


def axis(lo, hi, L):
    rs = np.clip(roundf((lo * scale).astype(np.float32)), -1e6, 1e6).astype(np.int64)
    re = np.clip(roundf((hi * scale).astype(np.float32)), -1e6, 1e6).astype(np.int64)
    n = np.maximum(re - rs + 1, 1)
    b = (n.astype(np.float32) / np.float32(7)).astype(np.float32)
    S, E = [], []
    for p in range(7):
        s = np.floor((np.float32(p) * b).astype(np.float32)).astype(np.int64) + rs
        e = np.ceil((np.float32(p + 1) * b).astype(np.float32)).astype(np.int64) + rs
        S.append(np.clip(s, 0, L))
        E.append(np.clip(e, 0, L))
    S, E = np.stack(S, 1), np.stack(E, 1)
    nn = E - S
    ne = nn > 0
    border = (S == 0) | (E == L)
    ok4 = np.all(~ne | (nn >= 4) | border, 1)
    ok2 = np.all(~ne | (nn >= 2) | border, 1)
    k = np.where(ok4, 4, np.where(ok2, 2, 1))
    nmax = np.max(np.where(ne, nn, 0), 1)
    c = np.maximum((nmax + k - 1) // k, 1)
    kk = k[:, None]
    p0 = np.where(nn >= kk, S, np.where(S == 0, E - kk, S))
    last = np.where(nn >= kk, nn - kk, 0)
    return k, c, p0, last, ne


kh, ch, p0h, lasth, neh = axis(rois[:, 2], rois[:, 4], H)
kw, cw, p0w, lastw, new = axis(rois[:, 1], rois[:, 3], W)
fb = (ch > 4) | (cw > 4)
CH, CW = np.where(ch <= 2, 2, 4), np.where(cw <= 2, 2, 4)


def loads_of(r):
    """(rows[49], cols[49], active[49]) of every block load of proposal r, in pyr_run's order"""
    bins = np.arange(49)
    ph, pw = bins // 7, bins % 7
    empty = ~(neh[r][ph] & new[r][pw])
    row0, col0, lh, lw = p0h[r][ph] + 3, p0w[r][pw] + 3, lasth[r][ph], lastw[r][pw]
    out = []
    for i in range(CH[r]):
        ni = (i == 0) | ((i - 1) * kh[r] < lh)
        ro = np.minimum(i * kh[r], lh)
        for j in range(CW[r]):
            nj = (j == 0) | ((j - 1) * kw[r] < lw)
            co = np.minimum(j * kw[r], lw)
            act = ni & nj if (i or j) else np.ones(49, bool)
            out.append((np.where(empty, H + 5, row0 + ro), np.where(empty, 0, col0 + co), act))
    return out


def cost(quarters, ld):
    tot = n = 0
    for rr, cc, act in ld:
        cell = rr * RS + cc
        for q in quarters:
            a = act[q]
            if not a.any():
                continue
            n += 1
            u = np.unique(cell[q][a])
            tot += np.bincount(u % 8, minlength=8).max()
    return tot, n


def rowmajor():
    return [np.arange(q, min(q + 8, 49)) for q in range(0, 49, 8)]


def firstfit(ld, bins, nq, r):
    """pyr_bins_kernel<GROUP>: bank-group masks of the loads (0,0), (0,1), (1,0), (1,1)"""
    consider = [0, 1, CW[r], CW[r] + 1]
    occ, fill, groups = [0] * nq, [0] * nq, [[] for _ in range(nq)]
    for b in bins:
        m = 0
        for k, l in enumerate(consider):
            rr, cc, act = ld[l]
            if act[b]:
                m |= 1 << (8 * k + int((rr[b] * RS + cc[b]) % 8))
        best = min((q for q in range(nq) if fill[q] < 8), key=lambda q: (bin(occ[q] & m).count("1"), -fill[q]))
        groups[best].append(b)
        fill[best] += 1
        occ[best] |= m
    return [np.array(g, int) for g in groups if g]


def lower_bound(ld):
    tot = 0
    for rr, cc, act in ld:
        cell = rr * RS + cc
        for bins in (np.arange(0, 32), np.arange(32, 49)):
            a = act[bins]
            if a.any():
                u = np.unique(cell[bins][a])
                tot += max(int(np.ceil(a.sum() / 8)), np.bincount(u % 8, minlength=8).max())
    return tot


res = collections.Counter()
for r in range(len(rois)):
    if fb[r]:
        continue
    ld = loads_of(r)
    a, n = cost(rowmajor(), ld)
    b, _ = cost(firstfit(ld, range(0, 32), 4, r) + firstfit(ld, range(32, 49), 3, r), ld)
    res.update(rowmajor=a, firstfit=b, loads=n, bound=lower_bound(ld))
print({k: int(v) for k, v in res.items()})
print("wavefronts per quarter-warp load: row-major %.3f, first fit %.3f, lower bound over lane orders %.3f"
      % (res["rowmajor"] / res["loads"], res["firstfit"] / res["loads"], res["bound"] / res["loads"]))
