"""Image processors = the `vision_tower.image_processor` dict of the reference (language_model/llava_phi.py:66-69,
keys panoptic | instance | semantic), consumed by the eval datasets as `processor.preprocess(data_dict, ...)`
(psalm/train/train_datasets.py:203-207).

Restates the INFERENCE part of the reference mappers (datasets_mapper/coco_panoptic_mapper.py:134-164 and siblings;
build_transform_gen :81-89): read the image, ResizeShortestEdge(short = max = IMAGE_SIZE), FixedSizeCrop(IMAGE_SIZE^2,
pad value 128, padding mask from a ones-mask padded with 0), CHW tensor, `(image - pixel_mean) / pixel_std`.
detectron2's transforms (un-vendored third party, version unpinned upstream) are restated from their documented
semantics: new size = round-half-up of the scaled size, PIL bilinear resampling for uint8 images, padding at the
bottom / right.  Ground-truth handling (panoptic PNGs, polygons) belongs to training / evaluation and is not restated.

On top of the reference contract, `preprocess` also returns `image_u8` (the padded uint8 CHW image): `PSALM.eval_seg`
accepts it directly and normalises on the device (csrc/preproc.cu), a 4x smaller upload."""
import numpy as np
import torch

PIXEL_MEAN = (123.675, 116.28, 103.53)   # coco_panoptic_mapper.py:118-119
PIXEL_STD = (58.395, 57.12, 57.375)


def resize_shortest_edge_shape(h, w, short, max_size):
    """detectron2 ResizeShortestEdge.get_output_shape: scale the short side to `short`, cap the long side at
    `max_size`, round half up."""
    scale = short * 1.0 / min(h, w)
    if h < w:
        newh, neww = short, scale * w
    else:
        newh, neww = scale * h, short
    if max(newh, neww) > max_size:
        scale = max_size * 1.0 / max(newh, neww)
        newh, neww = newh * scale, neww * scale
    return int(newh + 0.5), int(neww + 0.5)


class SegImageProcessor:
    """One class for the three mappers: they differ in ground-truth handling only (not restated)."""

    def __init__(self, image_size=1024, task="panoptic", pad_value=128.0):
        self.image_size, self.task, self.pad_value = int(image_size), task, float(pad_value)
        self.pixel_mean = torch.tensor(PIXEL_MEAN).view(-1, 1, 1)
        self.pixel_std = torch.tensor(PIXEL_STD).view(-1, 1, 1)

    def _read(self, dataset_dict):
        if "image_array" in dataset_dict:           # HWC uint8 RGB, already decoded
            return np.asarray(dataset_dict["image_array"], dtype=np.uint8)
        from PIL import Image
        with Image.open(dataset_dict["file_name"]) as im:
            return np.asarray(im.convert("RGB"), dtype=np.uint8)

    def preprocess(self, dataset_dict, region_mask_type=None, mask_format="polygon"):
        d = dict(dataset_dict)
        image = self._read(d)
        h, w = image.shape[:2]
        d.setdefault("height", h)
        d.setdefault("width", w)
        S = self.image_size
        nh, nw = resize_shortest_edge_shape(h, w, S, S)
        if (nh, nw) != (h, w):
            from PIL import Image
            image = np.asarray(Image.fromarray(image).resize((nw, nh), Image.BILINEAR))
        padded = np.full((S, S, 3), self.pad_value, dtype=np.float32).astype(np.uint8)
        padded[:nh, :nw] = image[:S, :S]
        padding_mask = np.ones((S, S), dtype=bool)
        padding_mask[:nh, :nw] = False               # ~ (ones-mask padded with 0)
        u8 = torch.as_tensor(np.ascontiguousarray(padded.transpose(2, 0, 1)))
        d["image_u8"] = u8
        d["image"] = (u8 - self.pixel_mean) / self.pixel_std
        d["padding_mask"] = torch.as_tensor(padding_mask)
        d["transforms"] = None
        return d


def build_image_processors(image_size=1024):
    """The dict `PSALMModel.__init__` attaches to the vision tower (llava_phi.py:66-69)."""
    return {k: SegImageProcessor(image_size, k) for k in ("panoptic", "instance", "semantic")}
