"""Reference-facing surface: types, image processors, `load_pretrained_model`,
`from_pretrained`, `prepare_inputs_labels_for_multimodal` - host logic, with the CUDA entry points emulated
(tests/emu.py) where a model is needed."""
import json
import os
import types

import numpy as np
import pytest
import torch

import emu
from oracle import psalm_oracle as O
from psalm_b200 import builder, hf_types, image_processor, synth
from psalm_b200.layout import PhiConfig, PsalmConfig

SMALL = PsalmConfig(phi=PhiConfig(hidden=256, layers=2, heads=4, inter=1024))


def test_types_of_the_reference_surface():
    assert hf_types.LlavaConfig.model_type == "llava_phi"                       # llava_phi.py:34-35
    cfg = hf_types.LlavaConfig()
    assert cfg.hidden_size == 2048 and cfg.num_hidden_layers == 24             # PhiConfig() == Phi-1.5
    out = hf_types.CausalOutputWithMask(loss=torch.tensor(1.0), loss_mask=torch.tensor(2.0))
    for f in ("loss", "logits", "past_key_values", "hidden_states", "attentions", "loss_mask", "loss_dice", "loss_SEG_class",
              "loss_class_name_class", "loss_region_class", "loss_llm"):         # llava_phi.py:37-49
        assert hasattr(out, f)
    hf_types.register()
    hf_types.register()   # idempotent


@pytest.mark.parametrize("hw,expect", [((480, 640), (768, 1024)), ((640, 480), (1024, 768)), ((1024, 1024), (1024, 1024)),
                                       ((333, 500), (682, 1024)), ((2000, 1000), (1024, 512))])
def test_resize_shortest_edge_shape(hw, expect):
    assert image_processor.resize_shortest_edge_shape(hw[0], hw[1], 1024, 1024) == expect


def test_image_processor_contract():
    """Keys / shapes / arithmetic of the mappers' inference part (coco_panoptic_mapper.py:134-164)."""
    procs = image_processor.build_image_processors(256)
    assert set(procs) == {"panoptic", "instance", "semantic"}                  # llava_phi.py:66-69
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (120, 160, 3), dtype=np.uint8)
    d = procs["panoptic"].preprocess({"image_array": img, "image_id": 7})
    assert d["height"] == 120 and d["width"] == 160 and d["image_id"] == 7
    assert tuple(d["image"].shape) == (3, 256, 256) and d["image"].dtype == torch.float32
    assert tuple(d["padding_mask"].shape) == (256, 256) and d["padding_mask"].dtype == torch.bool
    # 120 x 160 -> 192 x 256 (long side capped at 256), padded at the bottom with 128
    assert not d["padding_mask"][:192].any() and d["padding_mask"][192:].all()
    assert (d["image_u8"][:, 192:] == 128).all()
    mean = torch.tensor(image_processor.PIXEL_MEAN).view(3, 1, 1)
    std = torch.tensor(image_processor.PIXEL_STD).view(3, 1, 1)
    assert torch.equal(d["image"], (d["image_u8"] - mean) / std)              # the mapper's normalisation, bit for bit
    # an image that needs no resize is passed through unchanged
    same = procs["instance"].preprocess({"image_array": rng.integers(0, 256, (256, 256, 3), dtype=np.uint8)})
    assert not same["padding_mask"].any()


def _emu_cls(monkeypatch):
    from psalm_b200.psalm import PSALM
    emu.install(monkeypatch)

    class _EmuPSALM(PSALM):
        @staticmethod
        def _check_runtime(device):
            pass
    return _EmuPSALM


def test_from_pretrained_and_load_pretrained_model(monkeypatch, tmp_path):
    """Checkpoint directory in the reference's Hugging Face layout -> (tokenizer, model, image_processor, context_len)
    (psalm/model/builder.py:27-72)."""
    from safetensors.torch import save_file
    sd = synth.synth_state_dict(SMALL, seed=5)
    d = tmp_path / "ckpt"
    os.makedirs(d)
    save_file({k: v.contiguous() for k, v in sd.items()}, str(d / "model.safetensors"))
    json.dump({"model_type": "llava_phi", "hidden_size": 256, "num_hidden_layers": 2, "num_attention_heads": 4,
               "intermediate_size": 1024, "max_sequence_length": 1536}, open(d / "config.json", "w"))
    cls = _emu_cls(monkeypatch)
    m = builder.from_pretrained(cls, str(d), torch_dtype=torch.float32, device="cpu", seg_task="panoptic")
    assert m.cfg.phi.hidden == 256 and m.unexpected_keys == []
    monkeypatch.setattr("psalm_b200.psalm.PSALM", cls)
    args = types.SimpleNamespace(model_map_name="psalm", seg_task="referring")
    tok, model, proc, ctx = builder.load_pretrained_model(str(d), None, "psalm", args, torch_dtype=torch.float32, device="cpu")
    assert tok is None and ctx == 1536 and model.seg_task == "referring" and set(proc) == {"panoptic", "instance", "semantic"}
    with pytest.raises(KeyError):     # a checkpoint that misses tensors fails loudly
        bad = dict(sd)
        del bad["seg_query"]
        d2 = tmp_path / "bad"
        os.makedirs(d2)
        save_file({k: v.contiguous() for k, v in bad.items()}, str(d2 / "model.safetensors"))
        json.dump(json.load(open(d / "config.json")), open(d2 / "config.json", "w"))
        builder.from_pretrained(cls, str(d2), torch_dtype=torch.float32, device="cpu")
    with pytest.raises(KeyError, match="model_map_name"):     # builder.py:45-49: 'psalm' or 'psalm_video'
        builder.load_pretrained_model(str(d), None, "psalm", types.SimpleNamespace(model_map_name="psalm_audio"), device="cpu")
    from psalm_b200.psalm import PSALMForDAVISEval

    class _EmuVideo(PSALMForDAVISEval):
        @staticmethod
        def _check_runtime(device):
            pass
    monkeypatch.setattr("psalm_b200.psalm.PSALMForDAVISEval", _EmuVideo)
    args = types.SimpleNamespace(model_map_name="psalm_video", seg_task="region")
    _, vmodel, _, _ = builder.load_pretrained_model(str(d), None, "psalm", args, torch_dtype=torch.float32, device="cpu")
    assert isinstance(vmodel, PSALMForDAVISEval) and vmodel.region_on and vmodel.eval_video.__func__ is vmodel.eval_seg.__func__


def test_prepare_inputs_labels_for_multimodal_matches_the_oracle(monkeypatch):
    """llava_phi.py:767-971 through the reference-named method."""
    cls = _emu_cls(monkeypatch)
    sd = synth.synth_state_dict(SMALL, seed=1)
    inp = synth.synth_inputs(batch=2, height=64, width=96, task="panoptic", n_classes=6, seed=2, ragged=True)
    m = cls(sd, SMALL, torch.float32, "cpu", "panoptic")
    got = m.prepare_inputs_labels_for_multimodal(inp["input_ids"], inp["attention_mask"], None, None, inp["images"],
                                                  class_name_embedding_indices=inp["class_name_embedding_indices"],
                                                  class_name_ids=inp["class_name_ids"], cls_indices=inp["cls_indices"])
    ids, am, pkv, embeds, labels, qmask, cidx, rmask, ridx = got
    with torch.no_grad():
        feats = O.swin_forward(sd, "model.vision_tower.", inp["images"])
        img_tok = O.projector_forward(sd, "model.mm_projector.", feats[3])
        ref = O.assemble_sequence(sd, inp["input_ids"], inp["attention_mask"], img_tok, inp["class_name_ids"],
                                  inp["cls_indices"], inp["class_name_embedding_indices"])
    assert ids is None and rmask is None and ridx is None
    assert torch.allclose(embeds, ref["inputs_embeds"], atol=1e-4, rtol=1e-4)
    assert torch.equal(am, ref["attention_mask"]) and torch.equal(qmask, ref["seg_query_mask"])
    assert torch.equal(cidx, ref["class_name_embedding_indices"])
    m.initialize_vision_tokenizer(types.SimpleNamespace(mm_use_im_patch_token=False, mm_use_im_start_end=False), None)
    with pytest.raises(NotImplementedError):
        m.initialize_vision_tokenizer(types.SimpleNamespace(mm_use_im_patch_token=True, mm_use_im_start_end=False), None)


@pytest.mark.parametrize("shape", [(48, 64), (480, 640), (427, 640), (1200, 1600), (333, 500), (900, 600), (1024, 1024), (100, 37)])
def test_device_side_resize_has_pillow_bits(shape):
    """image_processor.pil_bilinear_resize restates Pillow's 8-bit BILINEAR resampler exactly (integer coefficients scaled
    by 2^22, horizontal pass then vertical pass with uint8 in between): bit-identical to Image.resize for up- and
    down-scaling, and preprocess_device == preprocess on `image_u8` / `padding_mask` (torch ops only: the same arithmetic
    runs on the GPU, checked against the CPU result in tests/test_preproc_gpu.py)."""
    import numpy as np
    from PIL import Image
    from psalm_b200.image_processor import SegImageProcessor, pil_bilinear_resize, resize_shortest_edge_shape
    h, w = shape
    img = (np.random.RandomState(h * 7 + w).rand(h, w, 3) * 255).astype(np.uint8)
    nh, nw = resize_shortest_edge_shape(h, w, 1024, 1024)
    ref = np.asarray(Image.fromarray(img).resize((nw, nh), Image.BILINEAR))
    got = pil_bilinear_resize(torch.from_numpy(img), nh, nw).numpy()
    assert np.array_equal(ref, got)
    proc = SegImageProcessor(1024, "panoptic")
    a = proc.preprocess({"image_array": img})
    b = proc.preprocess_device({"image_array": img}, device="cpu")
    assert torch.equal(a["image_u8"], b["image_u8"]) and torch.equal(a["padding_mask"], b["padding_mask"])
    assert (a["height"], a["width"]) == (b["height"], b["width"]) == (h, w)
