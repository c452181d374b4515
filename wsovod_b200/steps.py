"""One step of the region-scoring path per BASELINE config, driven through this package's public surface
(``wsovod_b200.modeling.WSOVODROIHeads``), for ``bench.py`` and the GPU tests.

  inference (c1, c2, c4)   pool (+ objectness scale) -> [box-head FCs: out of scope, region embeddings are synthetic]
                           -> alignment + softmax (tcgen05 TF32) -> per-class NMS + top-100
                           (wsovod/modeling/roi_heads/roi_heads.py:727-746,886-907)
  training (c3, c5)        ``WSOVODROIHeads.forward`` in training mode (roi_heads.py:648-884, trainer.py:37-84):
                           pool -> MIL forward -> BCE -> seeds -> pseudo-label assignment (+ torch-RNG subsample above 4096
                           proposals) -> alignment logits -> weighted CE / smooth-L1 -> backward of all of it, under
                           DistributedDataParallel when world > 1.

What stands in for the out-of-scope parts in the training step (north star: the box-head FCs stay PyTorch):
  * ``StandInBoxHead``: features = a strided slice of the pooled tensor (width ``F``): the data dependency on kernel 1 is
    kept, the 205 MFLOP / proposal of fc1 is not paid ("FC stub excluded");
  * its two parameters have the SIZE of fc1 / fc2 of the config's backbone (R18: 25088x4096 + 4096x4096 = 0.48 GB,
    R50: 100352x4096 + 4096x4096 = 1.71 GB of fp32 gradients, engine/defaults.py:146-148) and receive a zero
    gradient at the end of backward, where fc1's real gradient becomes ready, so DDP all-reduces the same bytes
    at the same point of the step; the small real Linears (cls / det / bbox_pred / projection MLP) train normally.
"""
import os

import torch
from torch import nn

from . import ops, synth
from .modeling import (InstanceRefinementOutputLayers, ObjectMiningOutputLayers, OpenVocabularyClassifier, ROIPooler,
                       WSOVODMixedDatasetsROIHeads, WSOVODROIHeads)
from .structures import Boxes, Instances

TRAIN_CONFIGS = ("c3", "c5")
FC_WIDTH = 4096                     # MODEL.ROI_BOX_HEAD.FC_DIM of the shipped configs


class _ZeroGrad(torch.autograd.Function):
    """identity on x whose backward hands every stand-in parameter a zero gradient of its own shape"""

    @staticmethod
    def forward(ctx, x, *params):
        ctx.shapes = [p.shape for p in params]
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        # a freshly zero-filled tensor per parameter: autograd takes it over as .grad without another copy, so the
        # stand-in costs what writing fc1's real gradient would cost (one pass over its bytes)
        return (g,) + tuple(g.new_zeros(s) for s in ctx.shapes)


class StandInBoxHead(nn.Module):
    def __init__(self, in_features, width, fc_in):
        super().__init__()
        self.width, self.step = width, max(in_features // width, 1)
        self.fc1_standin = nn.Parameter(torch.zeros(fc_in, FC_WIDTH))        # gradient bytes of fc1
        self.fc2_standin = nn.Parameter(torch.zeros(FC_WIDTH, FC_WIDTH))     # gradient bytes of fc2

    def forward(self, pooled):
        x = torch.flatten(pooled, start_dim=1)[:, :: self.step][:, : self.width].contiguous()
        return _ZeroGrad.apply(x, self.fc1_standin, self.fc2_standin)

    def grad_bytes(self):
        return 4 * (self.fc1_standin.numel() + self.fc2_standin.numel())


def _proposals(w, dev):
    props = []
    for n in range(w["N"]):
        a, b = w["offsets"][n], w["offsets"][n + 1]
        hw = (int(w["image_sizes"][n, 0]), int(w["image_sizes"][n, 1]))
        props.append(Instances(hw, proposal_boxes=Boxes(w["rois"][a:b, 1:].contiguous().to(dev)),
                               objectness_logits=w["objectness"][a:b].to(dev)))
    return props


class InferenceStep:
    """pool -> alignment + softmax -> detections on device-resident tensors of workload ``w``"""

    def __init__(self, w, dev, precision=ops.ALIGN_TF32, with_argmax=False):
        self.w, self.dev, self.precision, self.with_argmax = w, dev, precision, with_argmax
        self.feat, self.rois, self.obj = w["features"].to(dev), w["rois"].to(dev), w["objectness"].to(dev)
        self.emb, self.text = w["region_emb"].to(dev), w["text_emb"].to(dev)
        self.off = torch.tensor(w["offsets"], dtype=torch.int64, device=dev)
        self.sizes = w["image_sizes"].to(dev)
        self.boxes = self.rois[:, 1:].contiguous()
        self.proposals = w["N"] * w["R"]

    def pool(self, with_argmax=None):
        return ops.roi_pool(self.feat, self.rois, self.w["spatial_scale"], 7, self.obj, 1.0,
                            self.with_argmax if with_argmax is None else with_argmax)

    def align(self):
        return ops.align(self.emb, self.text, self.w["temperature"], 1, True, None, self.precision, False, True)[1]

    def detections(self, probs):
        w = self.w
        return ops.detections(probs, self.boxes, self.off, self.sizes, w["R"], w["score_thresh"], w["nms_thresh"], w["topk"],
                              ops.IOU_TV_CUDA)

    def eager(self):
        pooled, _ = self.pool()
        det = self.detections(self.align())
        return pooled, det

    def capture(self):
        """record the step's launches (12 kernels + 2 memsets, all on the current stream, no host sync) in a CUDA graph:
        ``__call__`` then replays it, so a launch-bound step (c1: 0.15 ms of kernels) is not timed at the pace of
        the Python wrappers.  The outputs live in the graph's private pool and are overwritten by the next replay."""
        from . import _lib
        self.eager()
        n0 = _lib.launch_count()
        self.eager()
        self.launches_per_step = _lib.launch_count() - n0      # kernels of libwsovod_b200.so in one step = one replay
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._out = self.eager()
        self._graph = g
        return self

    def __call__(self):
        if getattr(self, "_graph", None) is not None:
            self._graph.replay()
            return self._out
        return self.eager()


class TrainStep:
    """``WSOVODROIHeads.forward`` (training) + backward on workload ``w`` (c3: R50 COCO; c5: mixed VOC + COCO, where the
    miner / class count / text matrix follow the step's ``source_id`` and an inference pass with NMS follows)"""

    def __init__(self, w, dev, world=1, mixed=False, width=256, precision=ops.ALIGN_TF32, seed=0):
        self.w, self.dev, self.world, self.mixed = w, dev, world, mixed
        N, C, K, D = w["N"], w["C"], w["K"], w["D"]
        g = synth.gen(seed + 17)
        torch.manual_seed(seed)
        torch.backends.cuda.matmul.allow_tf32 = True          # the stand-in Linears are not what is measured
        head = StandInBoxHead(C * 49, width, C * 49)
        pooler = ROIPooler(7, (w["spatial_scale"],), 0, "ROIPool")
        ovc = OpenVocabularyClassifier(width, num_classes=K, weight_path="rand", weight_dim=D, precision=precision)
        refinery = [InstanceRefinementOutputLayers(width, K, ovc, test_score_thresh=w["score_thresh"],
                                                   test_nms_thresh=w["nms_thresh"], test_topk_per_image=w["topk"],
                                                   refine_reg=True)]
        common = dict(box_in_features=["res5"], box_pooler=pooler, box_head=head, box_refinery=refinery, refine_reg=[True],
                      sampling_on=True, batch_size_per_images=[4096], positive_sample_fractions=[1.0], pooler_type="ROIPool")
        if mixed:
            self.classes = [20, K]                                # VOC + COCO class counts
            miners = [ObjectMiningOutputLayers(width, k) for k in self.classes]
            heads = WSOVODMixedDatasetsROIHeads(object_miners=miners, num_classes_list=self.classes, **common)
        else:
            self.classes = [K]
            heads = WSOVODROIHeads(num_classes=K, object_miner=ObjectMiningOutputLayers(width, K), **common)
        self.heads = heads.to(dev)
        self.head = head
        self.model = self.heads
        if world > 1:
            from torch.nn.parallel import DistributedDataParallel as DDP
            # Buckets: the FC-sized gradients become ready last, nothing is left to overlap them with, and NCCL moves one
            # 1.7 GB message faster than 68 buckets of 25 MB (N = 2, c3: 8.0 -> 7.0 ms per step); WSOVOD_BUCKET_MB overrides.
            # The mixed-dataset step keeps DDP's default: with unused parameters (the other dataset's miner) in one
            # bucket with used ones the reducer rejects the undefined gradients.
            self.model = DDP(self.heads, device_ids=[dev.index], find_unused_parameters=mixed, gradient_as_bucket_view=True,
                             bucket_cap_mb=int(os.environ.get("WSOVOD_BUCKET_MB", "25" if mixed else "2048")))
        self.features = {"res5": w["features"].to(dev)}
        self.props = _proposals(w, dev)
        self.texts = [synth.text_embeddings(k, D, g).to(dev) for k in self.classes]
        self.targets = [[Instances(p.image_size, gt_classes=c.to(dev)) for p, c in zip(self.props, synth.image_labels(N, k, g, 8))]
                        for k in self.classes]
        self.proposals = N * w["R"] * (2 if mixed else 1)       # c5: the batch also goes through the inference pass
        self.count = 0

    def grad_bytes(self):
        return sum(4 * p.numel() for p in self.heads.parameters() if p.requires_grad)

    def __call__(self, sync=True):
        src = self.count % len(self.classes)
        self.count += 1
        self.heads.train()
        for p in self.heads.parameters():
            p.grad = None
        # single-dataset training scores against the head's stored text embeddings (classifier=None, the "rand"
        # Parameter here: its gradient comes from align_bwd); mixed-dataset training passes the batch's matrix
        kw = dict(targets=self.targets[src], classifier=self.texts[src] if self.mixed else None)
        if self.mixed:
            kw["source_id"] = src
        ctx = self.model.no_sync() if (not sync and self.world > 1) else _null()
        with ctx:
            _, losses = self.model(None, self.features, self.props, None, **kw)
            total = sum(losses.values())
            total.backward()
        out = {k: v.detach() for k, v in losses.items()}
        if self.mixed:                                            # "... stress with NMS and refinement" (BASELINE configs[4])
            self.heads.eval()
            with torch.no_grad():
                inst, _, _, _ = self.heads(None, self.features, self.props, None, classifier=self.texts[src])
            out["detections"] = sum(len(i) for i in inst)
        return out


class _null:
    def __enter__(self):
        return None

    def __exit__(self, *a):
        return False


def make(config, w, dev, world=1, precision=ops.ALIGN_TF32, graph=True):
    if config in TRAIN_CONFIGS:
        return TrainStep(w, dev, world, mixed=(config == "c5"), precision=precision)
    st = InferenceStep(w, dev, precision)
    return st.capture() if graph else st
