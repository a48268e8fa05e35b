"""Device input pipeline (csrc/preproc.cu): normalise + zero-pad + unfold, against the torch restatement in
tests/emu.py; and the staged (copy-stream) uint8 upload through PSALM.eval_seg against the float contract."""
import pytest
import torch

import emu
from psalm_b200 import kernels, synth
from psalm_b200.layout import PhiConfig, PsalmConfig

pytestmark = pytest.mark.gpu
SMALL = PsalmConfig(phi=PhiConfig(hidden=256, layers=2, heads=4, inter=1024))


@pytest.mark.parametrize("H,W", [(64, 64), (37, 50), (4, 4), (129, 3)])
@pytest.mark.parametrize("src", ["u8", "f32", "bf16"])
@pytest.mark.parametrize("dst", [torch.float32, torch.bfloat16, torch.float16])
def test_patchify_matches_restatement(H, W, src, dst):
    g = torch.Generator().manual_seed(H * 131 + W)
    mean = torch.tensor([123.675, 116.28, 103.53])
    std = torch.tensor([58.395, 57.12, 57.375])
    if src == "u8":
        img = torch.randint(0, 256, (2, 3, H, W), generator=g, dtype=torch.uint8)
        ref, _ = emu.patchify(img, dst, mean, std)
        out, (Wh, Ww) = kernels.patchify(img.cuda(), dst, mean.cuda(), std.cuda())
    else:
        img = torch.randn(2, 3, H, W, generator=g).to(torch.float32 if src == "f32" else torch.bfloat16)
        ref, _ = emu.patchify(img, dst)
        out, (Wh, Ww) = kernels.patchify(img.cuda(), dst)
    assert (Wh, Ww) == (-(-H // 4), -(-W // 4))
    assert torch.equal(out.cpu(), ref)      # byte shuffling + two fp32 operations: bit exact


def test_uint8_staged_eval_equals_float_contract():
    """eval_seg(images=uint8 staged on the copy stream) == eval_seg(images=(u8 - mean) / std computed on the host)."""
    from psalm_b200.psalm import PSALM
    sd = synth.synth_state_dict(SMALL, seed=0)
    inp = synth.synth_inputs(batch=2, height=192, width=192, task="panoptic", n_classes=20, seed=1)
    g = torch.Generator().manual_seed(5)
    u8 = torch.randint(0, 256, (2, 3, 192, 192), generator=g, dtype=torch.uint8)
    mean = torch.tensor([123.675, 116.28, 103.53]).view(1, 3, 1, 1)
    std = torch.tensor([58.395, 57.12, 57.375]).view(1, 3, 1, 1)
    flt = (u8 - mean) / std
    kw = {k: inp[k] for k in ("class_name_ids", "cls_indices", "class_name_embedding_indices", "is_thing_list")}
    for graph in (False, True):
        m = PSALM(sd, SMALL, torch.bfloat16, "cuda", "panoptic", use_cuda_graph=graph)
        a = m.eval_seg(input_ids=inp["input_ids"], attention_mask=inp["attention_mask"], images=flt,
                       seg_info=inp["seg_info"], **kw)
        sem_a = [r["sem_seg"].clone() for r in a]
        pinned = u8.pin_memory()
        for _ in range(3):   # ring of two staging buffers: reuse must wait for the consumer
            st = m.stage_images(pinned)
            b = m.eval_seg(input_ids=inp["input_ids"], attention_mask=inp["attention_mask"], images=st,
                           seg_info=inp["seg_info"], **kw)
            torch.cuda.synchronize()
            for x, r in zip(sem_a, b):
                assert torch.equal(x, r["sem_seg"])


def test_device_side_resize_equals_the_host_path():
    """SegImageProcessor.preprocess_device on the GPU: the float64 GEMMs evaluate Pillow's integer sums exactly, so the
    padded uint8 image equals the host (PIL) path bit for bit."""
    import numpy as np
    from psalm_b200.image_processor import SegImageProcessor
    proc = SegImageProcessor(1024, "panoptic")
    for h, w in ((480, 640), (1333, 800), (300, 1100)):
        img = (np.random.RandomState(h + w).rand(h, w, 3) * 255).astype(np.uint8)
        a = proc.preprocess({"image_array": img})
        b = proc.preprocess_device({"image_array": img}, device="cuda")
        assert b["image_u8"].is_cuda and torch.equal(a["image_u8"], b["image_u8"].cpu())
        assert torch.equal(a["padding_mask"], b["padding_mask"])
