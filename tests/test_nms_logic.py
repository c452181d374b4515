"""CPU checks of the three algorithmic claims the NMS kernels (wsovod_b200/csrc/nms.cu) rest on, restated in
numpy with IEEE fp32 arithmetic (no GPU, no library call):

  1. suppresses_fast(): whenever the division-free screen calls a pair "sure", its answer equals the exact
     test `fl(inter / den) > thr` of either torchvision arithmetic -- including pairs that sit exactly on the
     threshold (grid-snapped boxes) and degenerate boxes;
  2. the head stage of det_class: iterating kept = alive & !(covered_by & kept) to its fixed point gives the
     greedy NMS survivors;
  3. merge_lists() of det_topk: "own index + lower bound in the partner list", truncated to L, merges two
     sorted lists of unique keys, and a pairwise tree of such merges yields the L best of all lists.
"""
import numpy as np
import pytest

F = np.float32


def _boxes(n, rng, grid=None, degenerate=False):
    x1 = rng.uniform(0, 500, n)
    y1 = rng.uniform(0, 400, n)
    w = np.exp(rng.uniform(np.log(4), np.log(300), n))
    h = np.exp(rng.uniform(np.log(4), np.log(300), n))
    b = np.stack([x1, y1, x1 + w, y1 + h], 1)
    if grid:
        b = np.round(b / grid) * grid
    if degenerate:
        k = n // 10
        b[:k, 2] = b[:k, 0]                     # zero width
        b[k:2 * k, [0, 2]] = b[k:2 * k, [2, 0]]  # inverted
    return b.astype(F)


def _pair_terms(bi, bj, mode):
    """inter and den exactly as suppresses() / suppresses_fast() form them (fp32, round to nearest)"""
    w = np.maximum(np.minimum(bi[:, 2], bj[:, 2]) - np.maximum(bi[:, 0], bj[:, 0]), F(0)).astype(F)
    h = np.maximum(np.minimum(bi[:, 3], bj[:, 3]) - np.maximum(bi[:, 1], bj[:, 1]), F(0)).astype(F)
    inter = (w * h).astype(F)
    ai = ((bi[:, 2] - bi[:, 0]).astype(F) * (bi[:, 3] - bi[:, 1]).astype(F)).astype(F)
    wj, hj = (bj[:, 2] - bj[:, 0]).astype(F), (bj[:, 3] - bj[:, 1]).astype(F)
    if mode == 0:      # torchvision CPU: (ai + aj) - inter
        den = ((ai + (wj * hj).astype(F)).astype(F) - inter).astype(F)
    else:              # torchvision CUDA as compiled: fma(wj, hj, ai) - inter (one rounding for the fma)
        den = ((wj.astype(np.float64) * hj.astype(np.float64) + ai.astype(np.float64)).astype(F) - inter).astype(F)
    return inter, den


def _screen(inter, den, thr):
    """suppresses_fast(): returns (sure, answer)"""
    thr = F(thr)
    with np.errstate(all="ignore"):
        t = (thr * den).astype(F)
        e = (inter - t).astype(F)
        margin = (t * F(2.0 ** -18)).astype(F)
        sure = (thr >= F(1e-6)) & (thr <= F(1e6)) & (den >= F(2.0 ** -60)) & (np.abs(e) > margin)
    return sure, e > 0


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("thr", [0.3, 0.5, 0.25, 1.0 / 3.0, 0.2, 0.7, 1e-6, 1.0, 0.999999, 123.0])
def test_screen_never_disagrees_with_the_exact_test(mode, thr):
    rng = np.random.default_rng(int(thr * 1000) + mode)
    total_sure = total = 0
    for grid, degenerate in ((None, False), (32, False), (8, False), (None, True), (16, True)):
        b = _boxes(3000, rng, grid, degenerate)
        i, j = rng.integers(0, len(b), 400000), rng.integers(0, len(b), 400000)
        inter, den = _pair_terms(b[i], b[j], mode)
        with np.errstate(all="ignore"):
            exact = (inter / den).astype(F) > F(thr)
        sure, ans = _screen(inter, den, thr)
        assert np.array_equal(ans[sure], exact[sure])
        if not degenerate and not grid:            # snapping to a coarse grid flattens small boxes too
            total_sure += int(sure.sum())
            total += len(sure)
        if grid and 1e-3 < thr < 1.0:
            on_thr = np.isclose((inter / np.where(den == 0, 1, den)).astype(np.float64), thr, rtol=1e-6, atol=0) & (inter > 0)
            assert not sure[on_thr].any()          # quotients on the threshold always go to the exact test
    if 1e-6 <= thr <= 1e6:
        assert total_sure > 0.999 * total          # ... and, among well-formed boxes, almost nothing else does


def test_screen_is_off_for_thresholds_it_cannot_vouch_for():
    rng = np.random.default_rng(5)
    b = _boxes(1000, rng)
    i, j = rng.integers(0, 1000, 10000), rng.integers(0, 1000, 10000)
    inter, den = _pair_terms(b[i], b[j], 1)
    for thr in (0.0, -0.5, 1e-7, 1e7):
        assert not _screen(inter, den, thr)[0].any()


@pytest.mark.parametrize("density", [0.02, 0.1, 0.5, 0.9])
def test_fixed_point_of_the_ballot_equals_greedy(density):
    """32 candidates in score order, cov[i, j] = "j (ahead of i) covers i", some dead on arrival"""
    rng = np.random.default_rng(int(density * 100))
    for _ in range(2000):
        cov = np.tril(rng.random((32, 32)) < density, -1)
        alive = rng.random(32) < 0.8
        greedy = np.zeros(32, bool)
        for i in range(32):
            greedy[i] = alive[i] and not (cov[i] & greedy).any()
        kept = alive.copy()
        rounds = 0
        while True:
            nk = alive & ~(cov & kept[None, :]).any(1)
            rounds += 1
            if np.array_equal(nk, kept):
                break
            kept = nk
        assert np.array_equal(kept, greedy) and rounds <= 33


def _rank_merge(a, b, L):
    out = {}
    for own, other in ((a, b), (b, a)):
        for i, x in enumerate(own):
            lo = int(np.searchsorted(other[:max(0, min(len(other), L - i))], x, side="left"))
            if i + lo < L:
                assert i + lo not in out
                out[i + lo] = x
    n = min(L, len(a) + len(b))
    assert sorted(out) == list(range(n))
    return np.array([out[k] for k in range(n)], dtype=np.uint64)


@pytest.mark.parametrize("L", [1, 7, 100])
def test_rank_merge_and_merge_tree(L):
    rng = np.random.default_rng(L)
    for _ in range(200):
        K = int(rng.integers(1, 40))
        keys = rng.permutation(np.arange(1, 5000, dtype=np.uint64))
        lists, at = [], 0
        for _c in range(K):
            n = int(rng.integers(0, L + 1))
            lists.append(np.sort(keys[at:at + n]))
            at += n
        want = np.sort(np.concatenate(lists))[:L] if lists else np.zeros(0, np.uint64)
        cur = lists
        while len(cur) > 1:                         # merge_lists(): all pairs of a round, the odd list moves on
            nxt = [_rank_merge(cur[2 * p], cur[2 * p + 1], L) for p in range(len(cur) // 2)]
            if len(cur) & 1:
                nxt.append(cur[-1])
            cur = nxt
        assert np.array_equal(cur[0][:L], want)
