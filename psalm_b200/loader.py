"""Checkpoint loading for the reference's Hugging Face layout (psalm/model/builder.py:55 loads the same
directory with `PSALM.from_pretrained`): `config.json` + `model*.safetensors` or `pytorch_model*.bin`
shards (+ optional `*.index.json`).  Keys are used as they are (layout.py); nothing is renamed."""
import glob
import json
import os

import torch

from .layout import PsalmConfig, checkpoint_layout


def load_state_dict(path, map_location="cpu"):
    """Return the flat state dict of a checkpoint directory (or a single file)."""
    files = []
    if os.path.isdir(path):
        for pat in ("*.safetensors", "pytorch_model*.bin", "*.pt", "*.pth"):
            files = sorted(glob.glob(os.path.join(path, pat)))
            if files:
                break
        if not files:
            raise FileNotFoundError("no checkpoint shards (*.safetensors / pytorch_model*.bin) under %s" % path)
    else:
        files = [path]
    sd = {}
    for f in files:
        if f.endswith(".safetensors"):
            from safetensors.torch import load_file
            part = load_file(f, device=str(map_location))
        else:
            part = torch.load(f, map_location=map_location, weights_only=True)
            if isinstance(part, dict) and "state_dict" in part:
                part = part["state_dict"]
        sd.update(part)
    return sd


def validate_state_dict(sd, cfg: PsalmConfig = PsalmConfig(), strict_shapes=True):
    """Check `sd` against the layout contract; returns (missing, unexpected) like load_state_dict.
    lm_head / region_projector / unused decoder embeddings may be absent: the hot path never reads them."""
    optional = ("lm_head.", "region_projector.", "predictor.query_feat.", "predictor.SEG_query_embed.",
                "predictor.REGION_proj.")
    lay = checkpoint_layout(cfg)
    missing = [k for k in lay if k not in sd and not k.startswith(optional) and not k.endswith("num_batches_tracked")
               and not k.endswith("relative_position_index")]
    unexpected = [k for k in sd if k not in lay and "rotary_emb" not in k and "inv_freq" not in k]
    if strict_shapes:
        for k, (shape, _, _) in lay.items():
            if k in sd and tuple(sd[k].shape) != tuple(shape):
                raise ValueError("checkpoint tensor %s has shape %s, layout expects %s" % (k, tuple(sd[k].shape), shape))
    return missing, unexpected


def config_from_json(path):
    """Read config.json of the reference checkpoint (model_type llava_phi) into a PsalmConfig."""
    from .layout import PhiConfig
    with open(os.path.join(path, "config.json")) as f:
        c = json.load(f)
    phi = PhiConfig(hidden=c.get("hidden_size", 2048), layers=c.get("num_hidden_layers", 24),
                    heads=c.get("num_attention_heads", 32), inter=c.get("intermediate_size", 8192),
                    vocab=c.get("vocab_size", 51200), eps=c.get("layer_norm_eps", 1e-5),
                    rope_theta=c.get("rope_theta", 10000.0), rotary_frac=c.get("partial_rotary_factor", 0.5))
    if c.get("swin_type", "base") != "base":
        raise NotImplementedError("only the Swin-B tower of the released checkpoint is built (swin_type=%r)" % c.get("swin_type"))
    return PsalmConfig(phi=phi)
