"""CSC (wsovod/layers/csc/csc_cuda.cu): the oracle restatement on hand-checkable cases (CPU), and on the
GPU our kernels against the reference's OWN compiled extension (oracle/_ref/wsovod_ref_C.so::csc_forward) and the
oracle, bit for bit."""
import pytest
import torch

import oracle
from oracle import ref
from wsovod_b200 import synth


def _case(B, K, H, W, R, seed, frac_pos=0.5):
    g = synth.gen(seed)
    cpgs = torch.rand(B, K, H, W, generator=g) * (torch.rand(B, K, 1, 1, generator=g) > 0.2)
    blob = torch.zeros(B, K, H, W)
    for b in range(B):
        for c in range(K):
            y, x = int(torch.randint(0, H, (1,), generator=g)), int(torch.randint(0, W, (1,), generator=g))
            blob[b, c, max(y - H // 6, 0): y + H // 6 + 1, max(x - W // 6, 0): x + W // 6 + 1] = 0.5
    cpgs = torch.maximum(cpgs * 0.12, blob)
    labels = (torch.rand(B, K, generator=g) < frac_pos).float()
    preds = torch.rand(B, K, generator=g)
    boxes = synth.proposals(R, H, W, g) if R >= 100 else torch.rand(R, 4, generator=g) * torch.tensor([W, H, W, H]).float()
    x1, x2 = torch.minimum(boxes[:, 0], boxes[:, 2]), torch.maximum(boxes[:, 0], boxes[:, 2])
    y1, y2 = torch.minimum(boxes[:, 1], boxes[:, 3]), torch.maximum(boxes[:, 1], boxes[:, 3])
    rois = torch.stack([torch.zeros(R), x1, y1, x2, y2], 1)
    if R > 8:
        rois[:4, 1:] += torch.tensor([-30.0, -30.0, 40.0, 40.0])          # beyond the map: clamped
        rois[4:8, 1:] = (rois[4:8, 1:]).round() + 0.5                      # .5 ties of round()
    return cpgs, labels, preds, rois


def test_oracle_csc_hand_case():
    # one class, one image: a 4x4 blob of ones in an 8x8 map; the roi that frames the blob tightly scores highest
    cpgs = torch.zeros(1, 1, 8, 8)
    cpgs[0, 0, 2:6, 2:6] = 1.0
    rois = torch.tensor([[0, 2, 2, 5, 5], [0, 0, 0, 7, 7], [0, 6, 6, 7, 7]], dtype=torch.float32)
    W = oracle.csc(cpgs, torch.ones(1, 1), torch.ones(1, 1), rois, 0.1, True, 1.8)
    assert W.shape == (3, 1) and W[0, 0] == 1.0 and W[0, 0] > W[1, 0] and W[2, 0] <= 0.0
    # a class without a positive label keeps W = 1; preds = 0 blends everything to 1
    W2 = oracle.csc(cpgs, torch.zeros(1, 1), torch.ones(1, 1), rois)
    W3 = oracle.csc(cpgs, torch.ones(1, 1), torch.zeros(1, 1), rois)
    assert (W2 == 1).all() and (W3 == 1).all()


@pytest.mark.gpu
@pytest.mark.parametrize("B,K,H,W,R,seed", [(1, 20, 60, 80, 2000, 1), (2, 8, 100, 152, 700, 2), (3, 5, 17, 9, 300, 3),
                                            (1, 3, 480, 640, 4000, 4), (2, 4, 33, 65, 6, 5)])
@pytest.mark.parametrize("area_sqrt", [True, False])
def test_csc_vs_reference_extension_and_oracle(B, K, H, W, R, seed, area_sqrt):
    from wsovod_b200 import ops
    from wsovod_b200.layers import CSC
    cpgs, labels, preds, rois = _case(B, K, H, W, R, seed)
    dev = "cuda:0"
    args = [t.to(dev) for t in (cpgs, labels, preds, rois)]
    ours = ops.csc(*args, 0.1, area_sqrt, 1.8)
    o = oracle.csc(cpgs, labels, preds, rois, 0.1, area_sqrt, 1.8)
    assert torch.equal(ours.cpu(), o)
    ext = ref.cuda()
    if ext is not None:                      # the reference's own compiled csc_forward (GPU only)
        r = ext.csc_forward(*args, 0.7, False, 0.1, 0.2, 0.0, area_sqrt, 1.8)
        assert torch.equal(ours, r)
    Wm, PL, NL = CSC(area_sqrt=area_sqrt)(*args)
    assert torch.equal(Wm, ours) and torch.equal(PL, args[1]) and not NL.any()
    assert (ours[:, labels.max(0).values < 0.5] == 1).all()
