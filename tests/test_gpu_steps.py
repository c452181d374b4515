"""The benchmark's steps (wsovod_b200/steps.py) on small workloads: the training step really trains (every real
parameter, the "rand" text matrix included, receives a finite gradient; the FC stand-in a zero one of the FC layers'
size), the mixed-dataset step alternates sources and ends with an inference pass, the inference step equals the ops."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from wsovod_b200 import ops, steps, synth  # noqa: E402

DEV = torch.device("cuda", 0)


def _small(name, **over):
    cfg = dict(synth.CONFIGS[name])
    cfg.update(over)
    synth.CONFIGS["_t"] = cfg
    try:
        return synth.workload("_t", seed=3)
    finally:
        del synth.CONFIGS["_t"]


def test_training_step_c3_like():
    w = _small("c3", C=8, H=30, W=40, R=4500, K=20, D=64)            # R > 4096: the subsampling branch runs
    st = steps.TrainStep(w, DEV, world=1, mixed=False, width=32)
    a = st()
    b = st()
    assert set(a) == {"loss_cls_object_mining", "loss_cls_r0", "loss_box_reg_r0"}
    assert all(torch.isfinite(v) for v in a.values()) and all(torch.isfinite(v) for v in b.values())
    named = dict(st.heads.named_parameters())
    for n, p in named.items():
        assert p.grad is not None and torch.isfinite(p.grad).all(), n
        if "standin" in n:
            assert not p.grad.any()
        elif "bias" not in n:
            assert p.grad.abs().sum() > 0, n
    assert "box_refinery_0.cls.class_weight" in named
    assert st.head.grad_bytes() == 4 * (8 * 49 * 4096 + 4096 * 4096) and st.grad_bytes() > st.head.grad_bytes()


def test_training_step_c5_like_mixed():
    w = _small("c5", C=8, H=30, W=40, R=600, K=40, D=64)
    st = steps.TrainStep(w, DEV, world=1, mixed=True, width=32)
    o0, o1 = st(), st()
    assert st.classes == [20, 40]
    assert o0["detections"] > 0 and o1["detections"] > 0
    assert torch.isfinite(o0["loss_cls_r0"]) and torch.isfinite(o1["loss_cls_object_mining"])


def test_inference_step_matches_ops():
    w = _small("c2", N=2, C=8, H=30, W=40, R=1300, K=20, D=64)
    st = steps.InferenceStep(w, DEV)
    pooled, det = st()
    ref, _ = ops.roi_pool(st.feat, st.rois, w["spatial_scale"], 7, st.obj, 1.0, False)
    assert torch.equal(pooled, ref)
    probs = ops.align(st.emb, st.text, 50.0, 1, True, None, ops.ALIGN_TF32, False, True)[1]
    d2 = ops.detections(probs, st.boxes, st.off, st.sizes, w["R"], 1e-5, 0.3, 100, ops.IOU_TV_CUDA)
    assert all(torch.equal(det[k], d2[k]) for k in det)
