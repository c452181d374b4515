"""Minimal carriers with detectron2's field names so the drop-in runs without detectron2
installed.  Everything in wsovod_b200.modeling is duck-typed on ``.tensor`` / ``.proposal_boxes`` /
``.objectness_logits`` / ``.image_size`` / ``.gt_classes``; real detectron2 objects work unchanged."""
import torch


class Boxes:
    def __init__(self, tensor):
        if not isinstance(tensor, torch.Tensor):
            tensor = torch.as_tensor(tensor, dtype=torch.float32)
        self.tensor = tensor.to(torch.float32).reshape(-1, 4)

    def __len__(self):
        return self.tensor.shape[0]

    def __getitem__(self, item):
        return Boxes(self.tensor[item].reshape(-1, 4))

    def area(self):
        b = self.tensor
        return (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])

    def to(self, device):
        return Boxes(self.tensor.to(device))

    @property
    def device(self):
        return self.tensor.device


class Instances:
    def __init__(self, image_size, **fields):
        object.__setattr__(self, "_image_size", image_size)
        object.__setattr__(self, "_fields", {})
        for k, v in fields.items():
            self.set(k, v)

    @property
    def image_size(self):
        return self._image_size

    def __setattr__(self, name, value):
        self.set(name, value)

    def __getattr__(self, name):
        f = object.__getattribute__(self, "_fields")
        if name not in f:
            raise AttributeError(f"Cannot find field '{name}' in the given Instances!")
        return f[name]

    def set(self, name, value):
        self._fields[name] = value

    def has(self, name):
        return name in self._fields

    def get(self, name):
        return self._fields[name]

    def get_fields(self):
        return self._fields

    def __len__(self):
        for v in self._fields.values():
            return len(v)
        raise NotImplementedError("Empty Instances does not support __len__!")

    def __getitem__(self, item):
        r = Instances(self._image_size)
        for k, v in self._fields.items():
            r.set(k, v[item])
        return r
