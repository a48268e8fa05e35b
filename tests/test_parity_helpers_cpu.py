"""The parity metrics of oracle/parity.py (used by the full-size GPU parity tests, smoke() and bench.py) exercised
on CPU: the host pipeline with emulated kernels (tests/emu.py) must score ~0 against the oracle at every stage,
free-running and teacher-forced, and a perturbed output must be caught."""
import torch

import emu
from oracle import parity
from oracle import psalm_oracle as O
from psalm_b200 import synth
from psalm_b200.layout import PhiConfig, PsalmConfig

SMALL = PsalmConfig(phi=PhiConfig(hidden=256, layers=2, heads=4, inter=1024))
SMALL_O = dict(hidden=256, layers=2, heads=4, inter=1024, eps=1e-5, theta=10000.0, rotary_frac=0.5)


def test_parity_metrics_on_emulated_pipeline(monkeypatch):
    from psalm_b200.psalm import PSALM
    emu.install(monkeypatch)

    class _EmuPSALM(PSALM):
        @staticmethod
        def _check_runtime(device):
            pass
    task, H, W = "panoptic", 96, 128
    sd = parity.round_sd(synth.synth_state_dict(SMALL, seed=3), torch.bfloat16)
    assert all((v.to(torch.bfloat16).float() == v).all() for v in sd.values() if v.is_floating_point())
    inp = synth.synth_inputs(batch=1, height=H, width=W, task=task, n_classes=7, seed=4)
    with torch.no_grad():
        ores, it = O.eval_seg(sd, inp["input_ids"], inp["attention_mask"], inp["images"], inp["seg_info"],
                              class_name_ids=inp["class_name_ids"], cls_indices=inp["cls_indices"],
                              class_name_embedding_indices=inp["class_name_embedding_indices"],
                              is_thing_list=inp["is_thing_list"], task=task, phi_cfg=SMALL_O, return_intermediates=True)
    m = _EmuPSALM(sd, SMALL, torch.float32, "cpu", task)
    plan = m.make_plan(inp["input_ids"], inp["attention_mask"], (H, W), inp["class_name_ids"], inp["cls_indices"],
                       inp["class_name_embedding_indices"])
    trace = {}
    out = m.forward_core(inp["images"], plan, trace=trace)
    m.is_thing_list = inp["is_thing_list"]
    res = m.post_process(out, (H, W), inp["seg_info"])
    fm = parity.final_metrics(out, it)
    assert fm["mask_logit_maxrel"] < 1e-3 and fm["class_argmax_agree"] == 1.0 and fm["mask_sign_agree"] > 0.999
    rm = parity.result_metrics(res, ores, task)
    assert rm["panoptic_pixel_agree"] > 0.998 and rm["sem_argmax_agree"] > 0.998
    free = parity.free_stage_errors(trace, it)
    forced = parity.forced_stage_errors(m, it, inp["images"])
    assert max(free.values()) < 1e-3, free
    assert max(forced.values()) < 1e-3, forced
    layers = parity.forced_layer_errors(m, it)
    assert layers["layer_l2rel_max"] < 1e-4 and layers["mask_bit_agree_min"] > 0.9999, layers
    sens = parity.predictor_sensitivity(sd, it, torch.bfloat16)
    assert sens["pred_masks_l2rel"] > 1e-4      # bf16 input rounding alone moves the reference's own output
    # a corrupted result must show up
    bad = dict(out, pred_masks=out["pred_masks"] * 1.05)
    assert parity.final_metrics(bad, it)["mask_logit_maxrel"] > 1e-2
