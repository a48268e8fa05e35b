"""`load_pretrained_model` with the signature and return tuple of the reference (psalm/model/builder.py:27-72):

    tokenizer, model, image_processor, context_len = load_pretrained_model(model_path, model_base, model_name, model_args,
                                                                           mask_config=..., device="cuda")

so that the eval scripts (psalm/eval/panoptic_segmentation.py:127 and siblings) can be pointed at psalm_b200 by changing
one import.  `model` is a psalm_b200.PSALM built from the checkpoint directory's unchanged state dict; the task comes from
`model_args.seg_task` (default 'instance', builder.py:50); `image_processor` is the panoptic | instance | semantic dict."""
import os

import torch

from . import loader
from .image_processor import build_image_processors
from .layout import PsalmConfig


def load_pretrained_model(model_path, model_base=None, model_name=None, model_args=None,
                          mask_config="./psalm/mask_config/maskformer2_swin_base_384_bs16_50ep.yaml", load_8bit=False,
                          load_4bit=False, device_map="auto", device="cuda", torch_dtype=torch.float16, use_cuda_graph=False):
    if load_8bit or load_4bit:
        raise NotImplementedError("psalm_b200 runs 16-bit / fp32 storage; 8-bit / 4-bit loading (bitsandbytes) is not provided")
    from .psalm import PSALM, PSALMForDAVISEval
    model_map = {"psalm": PSALM, "psalm_video": PSALMForDAVISEval}      # builder.py:45-48
    map_name = getattr(model_args, "model_map_name", "psalm")
    if map_name not in model_map:
        raise KeyError("model_map_name %r: expected one of %s" % (map_name, sorted(model_map)))
    seg_task = getattr(model_args, "seg_task", "instance")          # builder.py:50
    tokenizer = None
    try:
        from transformers import AutoTokenizer
        tokenizer = AutoTokenizer.from_pretrained(model_path, use_fast=True)
    except Exception as e:   # no tokenizer files / no transformers: the segmentation path takes token ids
        tokenizer = None
        _ = e
    model = model_map[map_name].from_pretrained(model_path, torch_dtype=torch_dtype, device=device, seg_task=seg_task,
                                  use_cuda_graph=use_cuda_graph)
    image_processor = model.get_vision_tower().image_processor
    cfgj = os.path.join(model_path, "config.json")
    context_len = 2048
    if os.path.exists(cfgj):
        import json
        context_len = json.load(open(cfgj)).get("max_sequence_length", 2048)
    return tokenizer, model, image_processor, context_len


def from_pretrained(cls, model_path, mask_decoder_cfg=None, torch_dtype=torch.float16, device="cuda", seg_task="panoptic",
                    use_cuda_graph=False, **_ignored):
    """Body of `PSALM.from_pretrained` (the reference inherits it from transformers.PreTrainedModel, builder.py:55)."""
    sd = loader.load_state_dict(model_path)
    cfg = loader.config_from_json(model_path) if os.path.exists(os.path.join(model_path, "config.json")) else PsalmConfig()
    missing, unexpected = loader.validate_state_dict(sd, cfg)
    if missing:
        raise KeyError("checkpoint %s misses %d tensors of the PSALM layout, e.g. %s" % (model_path, len(missing), missing[:5]))
    model = cls(sd, cfg, torch_dtype, device, seg_task, use_cuda_graph=use_cuda_graph)
    model.unexpected_keys = unexpected
    model.name_or_path = model_path
    return model


__all__ = ["load_pretrained_model", "from_pretrained", "build_image_processors"]
