"""``OpenVocabularyClassifier`` with the reference's constructor and forward signature
(wsovod/modeling/class_heads/open_vocabulary_classifier.py:14-105).  The projection MLP stays PyTorch
(out of scope: "box-head FC" class); normalise . temperature . x @ W^T (+background, +bias) is one
fused kernel (tcgen05 TF32 by default, fp32 FMA on request)."""
from math import fabs

import numpy as np
import torch
from torch import nn

from .. import ops


class OpenVocabularyClassifier(nn.Module):
    def __init__(self, input_shape, *, num_classes, weight_path, weight_dim=512, use_bias=0.0, norm_weight=True,
                 norm_temperature=50.0, precision=ops.ALIGN_TF32):
        super().__init__()
        if isinstance(input_shape, int):
            input_size = input_shape
        else:
            input_size = input_shape.channels * (input_shape.width or 1) * (input_shape.height or 1)
        self.norm_weight, self.weight_dim, self.norm_temperature = norm_weight, weight_dim, norm_temperature
        self.precision = precision
        self.use_bias = fabs(use_bias) > 1e-9
        if self.use_bias:
            self.cls_bias = nn.Parameter(torch.ones(1) * use_bias)
        self.projection = nn.Sequential(nn.Linear(input_size, 1024), nn.ReLU(), nn.Linear(1024, weight_dim), nn.ReLU())
        # class_weight keeps the reference's layout and names, (D, K), Parameter for "rand" and buffer otherwise
        # (:47-65), so reference checkpoints load unchanged; the kernel consumes the (K, D) transpose.
        if weight_path == "rand":
            w = torch.randn((weight_dim, num_classes))
            nn.init.normal_(w, std=0.01)
        else:
            w = torch.tensor(np.load(weight_path, encoding="bytes", allow_pickle=True), dtype=torch.float32)
            w = w.permute(1, 0).contiguous()
        if norm_weight:                                    # normalised once at construction (:59-60)
            w = torch.nn.functional.normalize(w, p=2, dim=0)
        if weight_path == "rand":
            self.class_weight = nn.Parameter(w)
        else:
            self.register_buffer("class_weight", w)
        self._kd_cache = None

    def _stored_kd(self):
        """the stored weights as the (K, D) matrix the kernel reads; the copy of a buffer is cached"""
        w = self.class_weight
        if w.requires_grad and torch.is_grad_enabled():
            return w.t()                                   # autograd reaches the Parameter through the view
        key = (w.data_ptr(), w._version, w.device)
        if self._kd_cache is None or self._kd_cache[0] != key:
            self._kd_cache = (key, w.detach().t().contiguous())
        return self._kd_cache[1]

    def forward(self, x, classifier=None, append_background=False, want_probs=False):
        """x: B x input_size; classifier: (C', D) text embeddings or None (use the stored weights).
        Returns logits B x (C' [+1]); with ``want_probs`` also softmax(logits) from the same kernel."""
        x = self.projection(x)
        w = classifier if classifier is not None else self._stored_kd()
        # stored weights are used as they are (:91-92); a classifier passed in is normalised (:87-90)
        mode = 0 if not self.norm_weight else (1 if classifier is not None else 2)
        logits, probs = ops.align(x, w, self.norm_temperature, mode, append_background,
                                  self.cls_bias if self.use_bias else None, self.precision, True, want_probs)
        return (logits, probs) if want_probs else logits
