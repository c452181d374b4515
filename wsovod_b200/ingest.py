"""Proposal ingest: the on-disk format and the per-image preparation that feed kernel 1.

Host-side, numpy / torch CPU only -- this is the data format on the input side of the path, not a kernel.
Mirrors, with the same numpy calls where their tie behaviour matters:

  * `load_proposals_into_dataset` (wsovod/data/build.py:112-173): a pickled dict
    {"boxes": list[np.ndarray Nx4], "objectness_logits" | "scores": list[np.ndarray N],
     "ids" | "indexes": list, ["bbox_mode": int]}, read with encoding="latin1"; per image the proposals are
    sorted by score, descending, as `argsort()[::-1]` does (ties end up in reversed index order);
  * `unique_boxes` (wsovod/data/detection_utils.py:206-217): hash round(box * scale) . [1, 1e3, 1e6, 1e9],
    keep the first occurrence of every hash, in index order;
  * `transform_proposals` (:220-266): [transform] -> clip to the image -> unique -> drop boxes with a side
    <= min_box_size -> the first `proposal_topk`.

  * the concept (text) embedding file (open_vocabulary_classifier.py:51-57, rcnn_wsovod.py:299-305): whatever
    `np.load(path, encoding="bytes", allow_pickle=True)` reads -- a .npy array or a pickled (K, D) tensor --
    as a contiguous fp32 (K, D) matrix, the `classifier` argument of the alignment op.

`batch()` turns the prepared images into the (M, 5) roi tensor / offsets / objectness vector the pooling op
takes (poolers.py:81-108).
"""
import os
import pickle

import numpy as np
import torch

XYXY_ABS, XYWH_ABS = 0, 1        # detectron2.structures.BoxMode values used by proposal files


def load_proposal_file(path):
    """build.py:139-151 -- returns {"ids", "boxes", "objectness_logits", "bbox_mode"} with the D1 key names
    ("indexes", "scores") renamed"""
    with open(path, "rb") as f:
        d = pickle.load(f, encoding="latin1")
    for old, new in (("indexes", "ids"), ("scores", "objectness_logits")):
        if old in d:
            d[new] = d.pop(old)
    d["bbox_mode"] = int(d["bbox_mode"]) if "bbox_mode" in d else XYXY_ABS
    return d


def _index(pfile):
    """str(image id) -> position in the file's lists, built once per loaded file (build.py:156)"""
    if "_index" not in pfile:
        pfile["_index"] = {str(i): k for k, i in enumerate(pfile["ids"])}
    return pfile["_index"]


def image_proposals(pfile, image_id):
    """build.py:153-171 -- (boxes, objectness_logits) of one image, sorted by score descending"""
    k = _index(pfile)[str(image_id)]
    boxes, logits = pfile["boxes"][k], pfile["objectness_logits"][k]
    inds = logits.argsort()[::-1]
    return boxes[inds], logits[inds]


def load_proposals_into_dataset(dataset_dicts, proposal_file):
    """build.py:112-173 with its three forms of `proposal_file`: "" leaves the records alone (:134-135); a
    directory only records the per-image path `<dir>/<image_id>.pkl` (:137-142; `image_proposals_from_dir`
    reads such a file); a pickle file attaches proposal_boxes / proposal_objectness_logits /
    proposal_bbox_mode, score-sorted, to every record (:144-171)."""
    if proposal_file == "":
        return dataset_dicts
    if os.path.isdir(proposal_file):
        for record in dataset_dicts:
            record["proposal_file"] = proposal_file + "/" + str(record["image_id"]) + ".pkl"
        return dataset_dicts
    pfile = load_proposal_file(proposal_file)
    for record in dataset_dicts:
        boxes, logits = image_proposals(pfile, record["image_id"])
        record["proposal_boxes"] = boxes
        record["proposal_objectness_logits"] = logits
        record["proposal_bbox_mode"] = pfile["bbox_mode"]
    return dataset_dicts


def image_proposals_from_dir(record):
    """one image's `<image_id>.pkl` of the directory form (same dict layout, lists of length one, as
    tools/generate_sam_proposals_cuda.py writes per image): (boxes, objectness_logits), score-sorted"""
    pfile = load_proposal_file(record["proposal_file"])
    boxes, logits = pfile["boxes"][0], pfile["objectness_logits"][0]
    inds = logits.argsort()[::-1]
    return boxes[inds], logits[inds]


def unique_boxes(boxes, scale=1.0):
    """detection_utils.py:206-217 -- indices of the first occurrence of every rounded box, ascending"""
    boxes = boxes.numpy() if isinstance(boxes, torch.Tensor) else np.asarray(boxes)
    v = np.array([1, 1e3, 1e6, 1e9])
    hashes = np.round(boxes * scale).dot(v).astype(int)
    _, index = np.unique(hashes, return_index=True)
    return np.sort(index)


def transform_proposals(boxes, objectness_logits, image_shape, *, proposal_topk, min_box_size=0,
                        bbox_mode=XYXY_ABS, apply_box=None):
    """detection_utils.py:220-266 -- returns (boxes (n, 4) fp32 XYXY, objectness_logits (n,) fp32) as torch CPU
    tensors.  `apply_box` is the TransformList.apply_box of the image's augmentations (resize / flip), if any."""
    boxes = np.asarray(boxes)
    if bbox_mode == XYWH_ABS:                         # BoxMode.convert(XYWH_ABS -> XYXY_ABS)
        boxes = boxes.copy()
        boxes[:, 2] += boxes[:, 0]
        boxes[:, 3] += boxes[:, 1]
    elif bbox_mode != XYXY_ABS:
        raise ValueError("proposal files hold XYXY_ABS or XYWH_ABS boxes (got bbox_mode %r)" % (bbox_mode,))
    if apply_box is not None:
        boxes = apply_box(boxes)
    b = torch.as_tensor(boxes).to(torch.float32)      # Boxes(...) casts to float32
    logits = torch.as_tensor(np.asarray(objectness_logits).astype("float32"))
    if not torch.isfinite(b).all():
        raise AssertionError("Box tensor contains infinite or NaN!")
    h, w = image_shape
    b = torch.stack((b[:, 0].clamp(min=0, max=w), b[:, 1].clamp(min=0, max=h),
                     b[:, 2].clamp(min=0, max=w), b[:, 3].clamp(min=0, max=h)), dim=-1)   # Boxes.clip
    keep = torch.as_tensor(unique_boxes(b))
    b, logits = b[keep], logits[keep]
    keep = ((b[:, 2] - b[:, 0]) > min_box_size) & ((b[:, 3] - b[:, 1]) > min_box_size)       # Boxes.nonempty
    b, logits = b[keep], logits[keep]
    return b[:proposal_topk], logits[:proposal_topk]


def load_text_embeddings(path):
    """(K, D) fp32 concept embeddings, one row per class (rcnn_wsovod.py:301-303; the train-time buffer is the
    transpose, open_vocabulary_classifier.py:53-56)"""
    w = np.load(path, encoding="bytes", allow_pickle=True)
    w = (w.detach() if isinstance(w, torch.Tensor) else torch.as_tensor(np.asarray(w))).to(torch.float32).contiguous()
    if w.dim() != 2:
        raise ValueError("concept embedding file must hold a (K, D) matrix, got shape %s" % (tuple(w.shape),))
    return w


def batch(prepared):
    """list of (boxes, objectness_logits) per image -> rois (M, 5) [image index, x1, y1, x2, y2]
    (poolers.py:81-108), offsets (N + 1,) int64, objectness (M,)"""
    rows, off, obj = [], [0], []
    for i, (b, l) in enumerate(prepared):
        rows.append(torch.cat([torch.full((b.size(0), 1), float(i)), b], 1))
        off.append(off[-1] + b.size(0))
        obj.append(l)
    if not rows:
        return torch.zeros(0, 5), torch.zeros(1, dtype=torch.int64), torch.zeros(0)
    return torch.cat(rows, 0), torch.tensor(off, dtype=torch.int64), torch.cat(obj, 0)
