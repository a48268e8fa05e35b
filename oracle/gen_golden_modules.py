"""TEST INFRASTRUCTURE — golden outputs of the UNMODIFIED reference modules (run here on CPU through
oracle/ref_shims.py) on synthetic weights/inputs that can be regenerated anywhere from seeds
(psalm_b200/synth.py).  Only subsampled outputs are stored (fixtures stay small).

Also prints the max deviation of oracle/psalm_oracle.py from the reference on the same data, i.e. it
is the pinning run of the restatement.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")

from oracle import psalm_oracle as O  # noqa: E402
from oracle import ref_shims  # noqa: E402
from psalm_b200 import synth  # noqa: E402
from psalm_b200.layout import MaskConfig, PhiConfig, PsalmConfig, SwinConfig  # noqa: E402

SMALL_PHI = PhiConfig(hidden=256, layers=2, heads=4, inter=1024)
SMALL = PsalmConfig(phi=SMALL_PHI)
SMALL_PHI_ORACLE = dict(hidden=256, layers=2, heads=4, inter=1024, eps=1e-5, theta=10000.0, rotary_frac=0.5)


def subsample(t, n=4096, seed=0):
    """Deterministic subset of a tensor: (flat indices, values)."""
    t = t.detach().float().reshape(-1)
    g = torch.Generator().manual_seed(seed)
    idx = torch.randperm(t.numel(), generator=g)[: min(n, t.numel())].sort().values
    return idx.numpy().astype(np.int64), t[idx].numpy()


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30)), float((a - b).abs().max())


def build_ref(task, seed=0):
    davis = task == "davis"
    task = "region" if davis else task
    m = ref_shims.build_reference_psalm(task, num_hidden_layers=SMALL_PHI.layers, hidden_size=SMALL_PHI.hidden, davis=davis)
    sd = synth.synth_state_dict(SMALL, seed=seed, include_lm_head=True)
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all("rotary" in k or "inv_freq" in k for k in missing), missing
    m.train(False)
    return m, sd


def run_case(task, H, W, n_classes, seed, batch=1, ragged=False):
    m, sd = build_ref(task, seed)
    name = "e2e_%s_%dx%d_b%d" % (task, H, W, batch)
    davis = task == "davis"     # PSALMForDAVISEval: region prompts pooled from a visual-prompt frame
    task = "region" if davis else task
    inp = synth.synth_inputs(batch=batch, height=H, width=W, task=task, n_classes=n_classes, seed=seed + 1,
                             ragged=ragged, visual_prompt_frame=davis)
    gold = {}
    # ---- reference, module by module (hooks on the real modules) ----
    caps = {}

    def cap(key):
        def hook(mod, args, out):
            caps.setdefault(key, []).append(out)
        return hook

    hs = [m.model.vision_tower.register_forward_hook(cap("swin")),
          m.model.mm_projector.register_forward_hook(cap("proj")),
          m.predictor.register_forward_hook(cap("predictor")),
          m.seg_query_projector.register_forward_hook(cap("seg_query"))]
    orig_pd = m.pixel_decoder.forward_features

    def pd_wrap(feats):
        out = orig_pd(feats)
        caps["pixdec"] = out
        return out
    m.pixel_decoder.forward_features = pd_wrap
    orig_model_fwd = m.model.forward

    def model_wrap(*a, **k):
        out = orig_model_fwd(*a, **k)
        caps["hidden"] = out.last_hidden_state
        caps["inputs_embeds"] = k.get("inputs_embeds")
        caps["attn_mask"] = k.get("attention_mask")
        return out
    m.model.forward = model_wrap

    kw = dict(input_ids=inp["input_ids"], attention_mask=inp["attention_mask"], images=inp["images"],
              seg_info=inp["seg_info"], labels=inp["input_ids"].clone())
    for k in ("class_name_ids", "cls_indices", "class_name_embedding_indices", "token_refer_id",
              "refer_embedding_indices", "is_thing_list", "vp_images"):
        if k in inp:
            kw[k] = inp[k]
    if task == "region":
        torch.manual_seed(1234)   # region_pooling draws its sample points from the global CPU generator
    ref_results = m.eval_seg(**kw)
    for h in hs:
        h.remove()
    if task == "region":
        torch.manual_seed(1234)   # the oracle makes the same draws in the same order

    # ---- oracle on the same data ----
    ores, oi = O.eval_seg(sd, inp["input_ids"], inp["attention_mask"], inp["images"], inp["seg_info"],
                          class_name_ids=inp.get("class_name_ids"), cls_indices=inp.get("cls_indices"),
                          class_name_embedding_indices=inp.get("class_name_embedding_indices"),
                          token_refer_id=inp.get("token_refer_id"),
                          refer_embedding_indices=inp.get("refer_embedding_indices"),
                          is_thing_list=inp.get("is_thing_list"), task=task, phi_cfg=SMALL_PHI_ORACLE,
                          return_intermediates=True, vp_images=inp.get("vp_images"))
    rep = {}
    swin_ref = caps["swin"][0]
    for i in range(4):
        rep["swin_res%d" % (i + 2)] = rel(oi["feats"][i], swin_ref[i])
        gold["swin_res%d_idx" % (i + 2)], gold["swin_res%d" % (i + 2)] = subsample(swin_ref[i], seed=i)
    rep["proj"] = rel(oi["img_tok"], caps["proj"][0])
    gold["proj_idx"], gold["proj"] = subsample(caps["proj"][0])
    rep["inputs_embeds"] = rel(oi["seq"]["inputs_embeds"], caps["inputs_embeds"])
    assert torch.equal(oi["seq"]["attention_mask"], caps["attn_mask"].bool()), "attention mask mismatch"
    gold["attn_mask"] = caps["attn_mask"].numpy()
    gold["inputs_embeds_idx"], gold["inputs_embeds"] = subsample(caps["inputs_embeds"])
    rep["hidden"] = rel(oi["hidden"], caps["hidden"])
    gold["hidden_idx"], gold["hidden"] = subsample(caps["hidden"], n=16384)
    rep["seg_query"] = rel(oi["seg_query"], caps["seg_query"][0])
    gold["seg_query"] = caps["seg_query"][0].numpy()
    mf, _, ms = caps["pixdec"]
    rep["mask_features"] = rel(oi["mask_features"], mf)
    gold["mask_features_idx"], gold["mask_features"] = subsample(mf, n=16384)
    for i in range(3):
        rep["ms%d" % i] = rel(oi["ms"][i], ms[i])
        gold["ms%d_idx" % i], gold["ms%d" % i] = subsample(ms[i], n=8192, seed=i)
    pr = caps["predictor"][0]
    rep["pred_masks"] = rel(oi["predictor"]["pred_masks"], pr["pred_masks"])
    gold["pred_masks_idx"], gold["pred_masks"] = subsample(pr["pred_masks"], n=32768)
    gold["pred_masks_shape"] = np.array(pr["pred_masks"].shape)
    if pr["pred_class_name_logits"] is not None:
        rep["cls"] = rel(oi["predictor"]["pred_class_name_logits"], pr["pred_class_name_logits"])
        gold["pred_class_name_logits"] = pr["pred_class_name_logits"].numpy()
    if pr["pred_SEG_logits"] is not None:
        rep["SEG"] = rel(oi["predictor"]["pred_SEG_logits"], pr["pred_SEG_logits"])
        gold["pred_SEG_logits"] = pr["pred_SEG_logits"].numpy()
    if pr.get("pred_region_logits") is not None:
        for b, (a, r_) in enumerate(zip(oi["predictor"]["pred_region_logits"], pr["pred_region_logits"])):
            rep["region_logits_%d" % b] = rel(a, r_)
            gold["pred_region_logits_%d" % b] = r_.numpy()
            gold["region_points_%d" % b] = oi["region_points"][b].numpy()
    # ---- post-processed results (image 0 only: the reference returns inside the loop, LP:1472) ----
    r0, o0 = ref_results[0], ores[0]
    if "sem_seg" in r0:
        rep["sem_seg"] = rel(o0["sem_seg"], r0["sem_seg"])
        gold["sem_seg_argmax"] = r0["sem_seg"].argmax(0).numpy().astype(np.uint8)
        gold["sem_seg_idx"], gold["sem_seg"] = subsample(r0["sem_seg"], n=16384)
    if "panoptic_seg" in r0:
        pan, info = r0["panoptic_seg"]
        gold["panoptic_seg"] = pan.numpy().astype(np.int32)
        gold["panoptic_info"] = np.array([[d["id"], int(d["isthing"]), d["category_id"]] for d in info], dtype=np.int64).reshape(-1, 3)
        opan, oinfo = o0["panoptic_seg"]
        rep["panoptic_equal"] = (bool(torch.equal(opan, pan)), oinfo == info)
    if task == "region":
        inst = r0["instances"]
        gold["region_scores"] = inst.scores.numpy()
        gold["region_mask_area"] = inst.pred_masks.flatten(1).sum(1).numpy()
        gold["region_gt_idx"], gold["region_gt"] = subsample(r0["gt"], n=8192)
        rep["region_scores"] = rel(o0["instances"]["scores"], inst.scores)
        rep["region_masks_equal"] = bool(torch.equal(o0["instances"]["pred_masks"], inst.pred_masks))
        rep["region_gt"] = rel(o0["gt"], r0["gt"])
    elif "instances" in r0:
        inst = r0["instances"]
        sc = inst.scores
        order = torch.argsort(sc, descending=True, stable=True)
        gold["inst_scores_sorted"] = sc[order].numpy()
        if task != "referring":
            gold["inst_classes_sorted"] = inst.pred_classes[order].numpy()
        gold["inst_mask_area_sorted"] = inst.pred_masks.flatten(1).sum(1)[order].numpy()
        oinst = o0["instances"]
        oorder = torch.argsort(oinst["scores"], descending=True, stable=True)
        rep["inst_scores"] = rel(oinst["scores"][oorder], sc[order])
        rep["inst_area_equal"] = bool(torch.equal(oinst["pred_masks"].flatten(1).sum(1)[oorder],
                                                  inst.pred_masks.flatten(1).sum(1)[order]))
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **gold)
    print(name, json.dumps(rep, indent=None))
    return rep


CASES = [
    ("panoptic", 192, 192, 20, 0, 1, False),     # window-aligned grid (48 = 4 x 12): no Swin padding
    ("panoptic", 200, 264, 12, 3, 1, False),     # ragged: window padding, odd patch-merge sizes, 32-padding crop
    ("referring", 192, 192, 0, 5, 1, False),
    ("panoptic", 96, 128, 7, 7, 2, True),        # batch 2, ragged prompts -> right padding + attention mask
    ("region", 192, 192, 0, 9, 1, False),        # <region> prompts: point-sampled region features + REGION_proj head
    ("davis", 192, 192, 0, 13, 1, False),        # PSALMForDAVISEval: the regions come from a visual-prompt frame
]


def main():
    reps = {}
    only = sys.argv[1] if len(sys.argv) > 1 else None      # e.g. `python oracle/gen_golden_modules.py region`
    rp = os.path.join(GOLD, "oracle_vs_reference_report.json")
    if only and os.path.exists(rp):
        reps = json.load(open(rp))
    for c in CASES:
        if only and c[0] != only:
            continue
        reps["%s_%dx%d_b%d" % (c[0], c[1], c[2], c[5])] = run_case(*c)
    with open(os.path.join(GOLD, "oracle_vs_reference_report.json"), "w") as f:
        json.dump(reps, f, indent=1)


if __name__ == "__main__":
    with torch.no_grad():
        main()
