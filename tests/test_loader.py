import os

import pytest
import torch

from psalm_b200 import loader, synth
from psalm_b200.layout import PhiConfig, PsalmConfig

SMALL = PsalmConfig(phi=PhiConfig(hidden=64, layers=1, heads=1, inter=128, vocab=300))


def test_roundtrip_bin_and_safetensors(tmp_path):
    sd = synth.synth_state_dict(SMALL, seed=2, only_prefix="predictor.")
    sd.update(synth.synth_state_dict(SMALL, seed=2, only_prefix="model.layers"))
    d = tmp_path / "ckpt"
    os.makedirs(d)
    keys = sorted(sd)
    half = len(keys) // 2
    torch.save({k: sd[k] for k in keys[:half]}, d / "pytorch_model-00001-of-00002.bin")
    torch.save({k: sd[k] for k in keys[half:]}, d / "pytorch_model-00002-of-00002.bin")
    got = loader.load_state_dict(str(d))
    assert set(got) == set(sd) and all(torch.equal(got[k], sd[k]) for k in sd)
    from safetensors.torch import save_file
    d2 = tmp_path / "st"
    os.makedirs(d2)
    save_file({k: v.contiguous() for k, v in sd.items()}, str(d2 / "model.safetensors"))
    got = loader.load_state_dict(str(d2))
    assert all(torch.equal(got[k], sd[k]) for k in sd)


def test_validate_reports_missing_and_shape_errors():
    sd = synth.synth_state_dict(SMALL, seed=0)
    missing, unexpected = loader.validate_state_dict(sd, SMALL)
    assert missing == [] and unexpected == []
    del sd["seg_query"]
    sd["bogus.weight"] = torch.zeros(1)
    missing, unexpected = loader.validate_state_dict(sd, SMALL)
    assert missing == ["seg_query"] and unexpected == ["bogus.weight"]
    sd["pixel_decoder.mask_features.bias"] = torch.zeros(3)
    with pytest.raises(ValueError):
        loader.validate_state_dict(sd, SMALL)
