"""Minimal result containers with the attribute surface eval_seg's callers use
(detectron2.structures.Instances / Boxes, reference llava_phi.py:18, 318-323, 432-446)."""


class Boxes:
    def __init__(self, tensor):
        self.tensor = tensor

    def __len__(self):
        return self.tensor.shape[0]


class BitMasks:
    """detectron2.structures.BitMasks stand-in: `.tensor` [N,H,W] bool (region prompts, llava_phi.py:792)."""

    def __init__(self, tensor):
        self.tensor = tensor

    def __len__(self):
        return self.tensor.shape[0]


class Instances:
    def __init__(self, image_size, **fields):
        object.__setattr__(self, "_image_size", tuple(image_size))
        object.__setattr__(self, "_fields", {})
        for k, v in fields.items():
            self._fields[k] = v

    @property
    def image_size(self):
        return self._image_size

    def __setattr__(self, k, v):
        self._fields[k] = v

    def __getattr__(self, k):
        f = object.__getattribute__(self, "_fields")
        if k in f:
            return f[k]
        raise AttributeError("Cannot find field '%s' in the given Instances!" % k)

    def has(self, k):
        return k in self._fields

    def get_fields(self):
        return self._fields

    def __len__(self):
        for v in self._fields.values():
            return len(v)
        return 0
