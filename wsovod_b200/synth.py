"""Seeded synthetic inputs of the region-scoring path: features, SAM-like proposals,
region/text embeddings, MIL logits, image-level labels.  Shared by tests/ and bench.py so that the
CUDA path, the oracle and the CPU baseline see identical tensors.  CPU tensors; callers move them."""
import math

import torch

CONFIGS = {
    # name: images/GPU, channels, map H, W, proposals/image, classes, text dim
    "c1": dict(N=1, C=512, H=60, W=80, R=2000, K=20, D=768),      # VOC R18 480x640 (CPU-runnable)
    "c2": dict(N=8, C=512, H=86, W=128, R=4000, K=80, D=768),     # COCO R18 688x1024, batch 8/GPU
    "c3": dict(N=1, C=2048, H=100, W=152, R=5024, K=80, D=768),   # COCO R50 train 800x1216
    "c4": dict(N=8, C=512, H=86, W=128, R=4000, K=1203, D=768),   # LVIS-scale concepts
    "c5": dict(N=1, C=512, H=100, W=152, R=5000, K=80, D=768),    # mixed-dataset stress
}
STRIDE = 8


def gen(seed):
    return torch.Generator().manual_seed(int(seed))


def features(N, C, H, W, g, relu=True):
    x = torch.randn(N, C, H, W, generator=g)
    return torch.relu(x) if relu else x


def proposals(R, img_h, img_w, g, stress=True):
    """(R,4) XYXY fp32: x1~U(0,.85W) y1~U(0,.85H) w~LogU(16,.6W) h~LogU(16,.6H), clipped,
    min side 8 px; stress adds 5% exact duplicates and 1% degenerate (zero-width) boxes."""
    x1 = torch.rand(R, generator=g) * 0.85 * img_w
    y1 = torch.rand(R, generator=g) * 0.85 * img_h
    w = torch.exp(torch.rand(R, generator=g) * (math.log(0.6 * img_w) - math.log(16)) + math.log(16))
    h = torch.exp(torch.rand(R, generator=g) * (math.log(0.6 * img_h) - math.log(16)) + math.log(16))
    x2 = torch.minimum(x1 + w, torch.tensor(float(img_w)))
    y2 = torch.minimum(y1 + h, torch.tensor(float(img_h)))
    x1 = torch.minimum(x1, x2 - 8).clamp(min=0)
    y1 = torch.minimum(y1, y2 - 8).clamp(min=0)
    b = torch.stack([x1, y1, x2, y2], 1)
    if stress and R >= 100:
        nd, nz = R // 20, R // 100
        src = torch.randint(0, R, (nd,), generator=g)
        dst = torch.randint(0, R, (nd,), generator=g)
        b[dst] = b[src]
        z = torch.randint(0, R, (nz,), generator=g)
        b[z, 2] = b[z, 0]
    return b


def rois_from(boxes_per_image):
    """list of (Ri,4) -> (sum Ri, 5) pooler format (poolers.py:81-108) and offsets list."""
    rows, off = [], [0]
    for i, b in enumerate(boxes_per_image):
        rows.append(torch.cat([torch.full((b.size(0), 1), float(i)), b], 1))
        off.append(off[-1] + b.size(0))
    return torch.cat(rows, 0), off


def objectness(R, g):
    return 0.79 + 0.21 * torch.rand(R, generator=g)


def region_embeddings(M, D, g):
    x = torch.relu(torch.randn(M, D, generator=g))
    if M >= 1000:
        idx = torch.randint(0, M, (max(M // 1000, 1),), generator=g)
        x[idx] = 0            # 0.1% all-zero rows (eps path of F.normalize)
    return x


def text_embeddings(K, D, g):
    return torch.randn(K, D, generator=g)


def mil_logits(M, K, g):
    return torch.randn(M, K, generator=g) * 3, torch.randn(M, K, generator=g) * 3


def image_labels(N, K, g, max_labels=4):
    out = []
    for _ in range(N):
        n = int(torch.randint(1, max_labels + 1, (1,), generator=g))
        out.append(torch.randperm(K, generator=g)[:n].sort().values)
    return out


def workload(name, seed=1234, rank=0, stress=True):
    """All tensors of one step of config `name` for one GPU (CPU tensors)."""
    cfg = dict(CONFIGS[name])
    g = gen(seed + rank)
    N, C, H, W, R, K, D = (cfg[k] for k in "NCHWRKD")
    img_h, img_w = H * STRIDE, W * STRIDE
    boxes = [proposals(R, img_h, img_w, g, stress) for _ in range(N)]
    rois, offsets = rois_from(boxes)
    cfg.update(
        features=features(N, C, H, W, g), boxes=boxes, rois=rois, offsets=offsets,
        objectness=objectness(N * R, g), region_emb=region_embeddings(N * R, D, g),
        text_emb=text_embeddings(K, D, g), image_sizes=torch.tensor([[img_h, img_w]] * N, dtype=torch.float32),
        spatial_scale=1.0 / STRIDE, temperature=50.0, score_thresh=1e-5, nms_thresh=0.3, topk=100,
    )
    return cfg
