"""Host logic of the hot path, checked on CPU against the oracle and the reference-generated golden
fixtures with the CUDA entry points emulated by tests/emu.py (test-only monkeypatch)."""
import numpy as np
import pytest
import torch

import emu
from oracle import psalm_oracle as O
from psalm_b200 import sequence as SEQ
from psalm_b200 import synth
from psalm_b200.layout import PhiConfig, PsalmConfig

SMALL = PsalmConfig(phi=PhiConfig(hidden=256, layers=2, heads=4, inter=1024))
SMALL_O = dict(hidden=256, layers=2, heads=4, inter=1024, eps=1e-5, theta=10000.0, rotary_frac=0.5)


@pytest.mark.parametrize("task,ragged", [("panoptic", False), ("panoptic", True), ("referring", True)])
def test_sequence_plan_matches_oracle(task, ragged):
    """build_plan + materialize_embeds == the reference's per-token splice (oracle.assemble_sequence)."""
    inp = synth.synth_inputs(batch=3, height=64, width=64, task=task, n_classes=9, seed=4, ragged=ragged)
    g = torch.Generator().manual_seed(0)
    C, n_img, n_q = 32, 6, 100
    sd = {"model.embed_tokens.weight": torch.randn(51200, C, generator=g), "seg_query": torch.randn(n_q, C, generator=g)}
    img = torch.randn(3, n_img, C, generator=g)
    ref = O.assemble_sequence(sd, inp["input_ids"], inp["attention_mask"], img, inp.get("class_name_ids"),
                              inp.get("cls_indices"), inp.get("class_name_embedding_indices"),
                              inp.get("token_refer_id"), inp.get("refer_embedding_indices"))
    plan = SEQ.build_plan(inp["input_ids"], inp["attention_mask"], n_img, n_q, inp.get("class_name_ids"),
                          inp.get("cls_indices"), inp.get("class_name_embedding_indices"),
                          inp.get("token_refer_id"), inp.get("refer_embedding_indices"))
    emb = SEQ.materialize_embeds(plan, sd["model.embed_tokens.weight"], img, sd["seg_query"])
    assert torch.equal(emb, ref["inputs_embeds"])
    assert torch.equal(plan.attention_mask, ref["attention_mask"])
    hidden = torch.randn(3, plan.T, C, generator=g)
    assert torch.equal(SEQ.gather_seg_query(plan, hidden), O.get_seg_query(hidden, ref["seg_query_mask"]))
    if task == "panoptic":
        assert torch.allclose(SEQ.pool(plan.cls_pool, hidden),
                              O.get_class_name_embedding(hidden, ref["class_name_embedding_indices"]), atol=1e-6)
    else:
        assert torch.allclose(SEQ.pool(plan.refer_pool, hidden),
                              O.get_SEG_embedding(hidden, ref["refer_embedding_indices"]), atol=1e-6)


def _emu_model(monkeypatch, sd, task):
    from psalm_b200.psalm import PSALM
    emu.install(monkeypatch)

    class _EmuPSALM(PSALM):
        @staticmethod
        def _check_runtime(device):   # tests only: host-logic check with emulated kernels
            pass
    return _EmuPSALM(sd, SMALL, torch.float32, "cpu", task)


CASES = [("panoptic", 192, 192, 20, 0, 1, False), ("panoptic", 200, 264, 12, 3, 1, False),
         ("referring", 192, 192, 0, 5, 1, False), ("panoptic", 96, 128, 7, 7, 2, True)]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "%s_%dx%d_b%d" % (c[0], c[1], c[2], c[5]))
def test_host_pipeline_matches_golden(monkeypatch, golden, case):
    task, H, W, ncls, seed, batch, ragged = case
    sd = synth.synth_state_dict(SMALL, seed=seed)
    inp = synth.synth_inputs(batch=batch, height=H, width=W, task=task, n_classes=ncls, seed=seed + 1, ragged=ragged)
    m = _emu_model(monkeypatch, sd, task)
    g = golden("e2e_%s_%dx%d_b%d.npz" % (task, H, W, batch))
    plan = m.make_plan(inp["input_ids"], inp["attention_mask"], (H, W), inp.get("class_name_ids"), inp.get("cls_indices"),
                       inp.get("class_name_embedding_indices"), inp.get("token_refer_id"), inp.get("refer_embedding_indices"))
    out = m.forward_core(inp["images"], plan)
    H4, W4 = out["mask_size"]
    pm = out["pred_masks"].view(batch, -1, H4, W4)
    assert list(pm.shape) == g["pred_masks_shape"].tolist()
    got = pm.reshape(-1)[torch.from_numpy(g["pred_masks_idx"])].numpy()
    scale = np.abs(g["pred_masks"]).max()
    assert np.abs(got - g["pred_masks"]).max() / scale < 2e-3
    if "pred_class_name_logits" in g:
        cl = out["pred_class_name_logits"].numpy()
        assert np.allclose(cl, g["pred_class_name_logits"], rtol=1e-3, atol=2e-3)
        assert np.array_equal(cl.argmax(-1), g["pred_class_name_logits"].argmax(-1))
    if "pred_SEG_logits" in g:
        assert np.allclose(out["pred_SEG_logits"].numpy(), g["pred_SEG_logits"], rtol=1e-3, atol=2e-3)
    m.is_thing_list = inp.get("is_thing_list")
    res = m.post_process(out, (H, W), inp["seg_info"])
    if "panoptic_seg" in g:
        pan, info = res[0]["panoptic_seg"]
        assert (pan.numpy() != g["panoptic_seg"]).mean() < 2e-3
        assert [[d["id"], int(d["isthing"]), d["category_id"]] for d in info] == g["panoptic_info"].tolist()
        assert np.array_equal(res[0]["sem_seg"].argmax(0).numpy().astype(np.uint8), g["sem_seg_argmax"]) or \
            (res[0]["sem_seg"].argmax(0).numpy().astype(np.uint8) != g["sem_seg_argmax"]).mean() < 2e-3
    if "inst_scores_sorted" in g:
        sc = res[0]["instances"].scores
        order = torch.argsort(sc, descending=True, stable=True)
        assert np.allclose(sc[order].numpy(), g["inst_scores_sorted"], rtol=1e-3, atol=1e-4)


def test_region_sequence_plan_matches_oracle():
    """<region> rows: build_plan + materialize_embeds == oracle.assemble_sequence with region features, and the
    region-embedding rows gathered after the LLM are the masked rows of the reference (llava_phi.py:302-307)."""
    inp = synth.synth_inputs(batch=3, height=64, width=64, task="region", seed=4, ragged=True)
    g = torch.Generator().manual_seed(0)
    C, n_img, n_q = 32, 4, 100
    sd = {"model.embed_tokens.weight": torch.randn(51200, C, generator=g), "seg_query": torch.randn(n_q, C, generator=g)}
    img = torch.randn(3, n_img, C, generator=g)
    counts = [int((inp["input_ids"][b] == -203).sum()) for b in range(3)]
    feats = [torch.randn(k, 1, C, generator=g) for k in counts]
    ref = O.assemble_sequence(sd, inp["input_ids"], inp["attention_mask"], img, region_features=feats)
    plan = SEQ.build_plan(inp["input_ids"], inp["attention_mask"], n_img, n_q)
    assert plan.region_counts == tuple(counts)
    emb = SEQ.materialize_embeds(plan, sd["model.embed_tokens.weight"], img, sd["seg_query"], torch.cat(feats, 0).squeeze(1))
    assert torch.equal(emb, ref["inputs_embeds"]) and torch.equal(plan.attention_mask, ref["attention_mask"])
    hidden = torch.randn(3, plan.T, C, generator=g)
    want = torch.cat([h[m.bool()] for h, m in zip(hidden, ref["region_embedding_masks"])], 0)
    assert torch.equal(SEQ.gather_region_rows(plan, hidden), want)


def test_region_points_follow_the_reference_rng_stream():
    """psalm_b200.region.sample_region_points makes the reference's draws in the reference's order
    (context_cluster.py:31-40): with the same seed it returns the oracle's points (which are pinned to the reference
    by tests/golden/e2e_region_192x192_b1.npz, generated with the reference's region_pooling under the same seed)."""
    from psalm_b200 import region as R
    inp = synth.synth_inputs(batch=1, height=192, width=192, task="region", seed=10)
    masks = inp["seg_info"][0]["instances"].region_masks.tensor
    torch.manual_seed(1234)
    a = R.sample_region_points(masks)
    torch.manual_seed(1234)
    b = O.sample_region_points(masks)
    assert a.shape == (masks.shape[0], 256, 2) and torch.equal(a, b.float())
    # small masks: every mask pixel is present (plus repeats); large masks: 256 distinct pixels of the mask
    for m, p in zip(masks, a):
        yx = (p * torch.tensor([192.0, 192.0])).round().long()
        assert bool(m[yx[:, 0], yx[:, 1]].all())
        n = int(m.sum())
        assert len({tuple(t) for t in yx.tolist()}) == min(n, 256)


def test_region_host_pipeline_matches_reference_golden(monkeypatch, golden):
    """The whole region path with emulated kernels against the fixture the UNMODIFIED reference produced
    (oracle/gen_golden_modules.py, case region_192x192): region logits [K,Q], scores [Q,K], thresholded masks, gt."""
    H = W = 192
    sd = synth.synth_state_dict(SMALL, seed=9)
    inp = synth.synth_inputs(batch=1, height=H, width=W, task="region", seed=10)
    m = _emu_model(monkeypatch, sd, "region")
    g = golden("e2e_region_192x192_b1.npz")
    pts = [torch.from_numpy(g["region_points_0"])]
    plan = m.make_plan(inp["input_ids"], inp["attention_mask"], (H, W))
    from psalm_b200.region import region_inputs
    plan.region_points, plan.region_image, counts = region_inputs(inp["seg_info"], pts)
    assert counts == plan.region_counts
    out = m.forward_core(inp["images"], plan)
    rl = out["pred_region_logits"][0].numpy()
    assert rl.shape == g["pred_region_logits_0"].shape
    assert np.abs(rl - g["pred_region_logits_0"]).max() / np.abs(g["pred_region_logits_0"]).max() < 2e-3
    res = m.post_process(out, (H, W), inp["seg_info"])
    inst = res[0]["instances"]
    assert np.allclose(inst.scores.numpy(), g["region_scores"], rtol=2e-3, atol=2e-4)
    area = inst.pred_masks.flatten(1).sum(1).numpy()
    assert np.abs(area - g["region_mask_area"]).max() <= 2          # |logit| ~ 0 ties only
    gt = res[0]["gt"].reshape(-1)[torch.from_numpy(g["region_gt_idx"])].numpy()
    assert np.allclose(gt, g["region_gt"], atol=1e-6)


def test_davis_variant_matches_reference_golden(monkeypatch, golden):
    """PSALMForDAVISEval: region features pooled from the visual-prompt frame (`vp_images`, `vp_region_masks`), against
    the fixture produced by the UNMODIFIED reference class (oracle/gen_golden_modules.py, case davis_192x192)."""
    from psalm_b200.psalm import PSALMForDAVISEval
    H = W = 192
    sd = synth.synth_state_dict(SMALL, seed=13)
    inp = synth.synth_inputs(batch=1, height=H, width=W, task="region", seed=14, visual_prompt_frame=True)
    emu.install(monkeypatch)

    class _Emu(PSALMForDAVISEval):
        @staticmethod
        def _check_runtime(device):
            pass
    m = _Emu(sd, SMALL, torch.float32, "cpu", "region")
    g = golden("e2e_davis_192x192_b1.npz")
    from psalm_b200.region import region_inputs
    plan = m.make_plan(inp["input_ids"], inp["attention_mask"], (H, W))       # (eval_video itself needs CUDA streams)
    plan.region_points, plan.region_image, _ = region_inputs(inp["seg_info"], [torch.from_numpy(g["region_points_0"])],
                                                             "vp_region_masks")
    plan.vp_images = inp["vp_images"]
    res = m.post_process(m.forward_core(inp["images"], plan), (H, W), inp["seg_info"])
    inst = res[0]["instances"]
    assert np.allclose(inst.scores.numpy(), g["region_scores"], rtol=2e-3, atol=2e-4)
    assert np.abs(inst.pred_masks.flatten(1).sum(1).numpy() - g["region_mask_area"]).max() <= 2
    with pytest.raises(ValueError, match="vp_images"):
        m.eval_seg(input_ids=inp["input_ids"], attention_mask=inp["attention_mask"], images=inp["images"], seg_info=inp["seg_info"])
