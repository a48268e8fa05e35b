"""TEST INFRASTRUCTURE — parity metrics between psalm_b200.PSALM (CUDA) and the CPU oracle
(oracle/psalm_oracle.py).  Imported only by tests/, `__graft_entry__.smoke()` and bench.py's cpu_baseline
leg (as the checker of the measured path, never as the path).

What is compared, and why it is split this way ("what's weak" 1-3):
  * `final_metrics`      final mask logits / class logits of a free-running GPU pass vs the oracle.
  * `result_metrics`     the task outputs (panoptic map, semantic arg-max, instance scores).
  * `free_stage_errors`  l2-relative error of every stage output of the free-running pass (Swin levels,
                         image tokens, LLM hidden states, pixel decoder, mask logits): shows WHERE a
                         16-bit run drifts.
  * `forced_stage_errors` every GPU stage fed the ORACLE's inputs of that stage (rounded to the storage
                         type): the residual is that stage's kernels + its own storage rounding, with no
                         error inherited from earlier stages ("teacher forced").
16-bit runs are compared with an oracle that was given the SAME 16-bit-rounded weights (`round_sd`), so
weight quantisation is not part of the reported error.
"""
import numpy as np
import torch


def round_sd(sd, dtype):
    """State dict whose floating tensors hold `dtype`-representable values (stored as fp32)."""
    if dtype == torch.float32:
        return sd
    return {k: (v.to(dtype).float() if v.is_floating_point() else v) for k, v in sd.items()}


def _l2rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-300))


def _maxrel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-300))


def _tok(nchw):
    """[B,C,H,W] -> token-major [B,H*W,C]."""
    B, C, H, W = nchw.shape
    return nchw.permute(0, 2, 3, 1).reshape(B, H * W, C)


def final_metrics(out, it):
    """out: PSALM.forward_core result (device); it: oracle intermediates."""
    po = it["predictor"]
    ref = po["pred_masks"].flatten(2)
    got = out["pred_masks"].float().cpu().view_as(ref)
    m = dict(mask_logit_maxrel=_maxrel(got, ref), mask_logit_l2rel=_l2rel(got, ref),
             mask_sign_agree=float(((got > 0) == (ref > 0)).double().mean()))
    if po.get("pred_class_name_logits") is not None:
        cg, cr = out["pred_class_name_logits"].float().cpu(), po["pred_class_name_logits"]
        m["class_logit_maxrel"] = _maxrel(cg, cr)
        m["class_argmax_agree"] = float((cg.argmax(-1) == cr.argmax(-1)).double().mean())
    if po.get("pred_SEG_logits") is not None:
        sg, sr = out["pred_SEG_logits"].float().cpu(), po["pred_SEG_logits"]
        m["seg_logit_maxrel"] = _maxrel(sg, sr)
        m["seg_top1_agree"] = float((sg.flatten(1).argmax(-1) == sr.flatten(1).argmax(-1)).double().mean())
    return m


def result_metrics(res, ores, task):
    """res: PSALM.eval_seg results; ores: oracle results (lists, one dict per image)."""
    m = {}
    pan_ag, sem_ag, seg_eq, sc = [], [], [], []
    for r, o in zip(res, ores):
        if "panoptic_seg" in o:
            pa, pb = r["panoptic_seg"][0].cpu(), o["panoptic_seg"][0]
            pan_ag.append(float((pa == pb).double().mean()))
            seg_eq.append(float([(d["id"], bool(d["isthing"]), d["category_id"]) for d in r["panoptic_seg"][1]] ==
                                [(d["id"], bool(d["isthing"]), d["category_id"]) for d in o["panoptic_seg"][1]]))
        if "sem_seg" in o:
            sem_ag.append(float((r["sem_seg"].argmax(0).cpu() == o["sem_seg"].argmax(0)).double().mean()))
        if "instances" in o:
            sa = torch.sort(r["instances"].scores.float().cpu(), descending=True).values
            sb = torch.sort(o["instances"]["scores"].float(), descending=True).values
            n = min(len(sa), len(sb))
            sc.append(float((sa[:n] - sb[:n]).abs().max() / (sb.abs().max() + 1e-30)) if n else 0.0)
            m.setdefault("n_instances", []).append((int(len(sa)), int(len(sb))))
    if pan_ag:
        m["panoptic_pixel_agree"] = float(np.mean(pan_ag))
        m["panoptic_segments_equal"] = float(np.mean(seg_eq))
        m["n_segments"] = [len(o["panoptic_seg"][1]) for o in ores]
    if sem_ag:
        m["sem_argmax_agree"] = float(np.mean(sem_ag))
    if sc:
        m["instance_score_maxrel"] = float(np.max(sc))
    return m


def free_stage_errors(trace, it):
    """Stage outputs of a free-running GPU pass (`trace` from PSALM.forward_core) vs the oracle's."""
    e = {}
    for i, name in enumerate(("res2", "res3", "res4", "res5")):
        e["swin_" + name] = _l2rel(trace["swin"][i].float().cpu(), _tok(it["feats"][i]))
    e["img_tok"] = _l2rel(trace["img_tok"].float().cpu(), it["img_tok"])
    e["hidden"] = _l2rel(trace["hidden"].float().cpu(), it["hidden"])
    e["seg_query"] = _l2rel(trace["seg_query"].float().cpu(), it["seg_query"])
    if it.get("cls_emb") is not None:
        e["cls_emb"] = _l2rel(trace["cls_emb"].float().cpu(), it["cls_emb"])
    e["mask_features"] = _l2rel(trace["mask_features"].float().cpu(), _tok(it["mask_features"]))
    for i in range(3):
        e["ms%d" % i] = _l2rel(trace["ms"][i].float().cpu(), _tok(it["ms"][i]))
    return e


def forced_stage_errors(model, it, images):
    """Each GPU stage on the ORACLE's inputs of that stage (cast to the model's storage type)."""
    dev, dt = model.device, model.dtype
    g = lambda t: t.to(device=dev, dtype=dt)  # noqa: E731
    e = {}
    with torch.no_grad(), model._precision_scope():
        toks, sizes = model.model.vision_tower.forward_tokens(images.to(dev))
        for i, name in enumerate(("res2", "res3", "res4", "res5")):
            e["swin_" + name] = _l2rel(toks[i].float().cpu(), _tok(it["feats"][i]))
        e["img_tok"] = _l2rel(model.model.mm_projector(g(it["feats"][3])).float().cpu(), it["img_tok"])
        seq = it["seq"]
        am = seq["attention_mask"]
        hidden = model.model.phi(g(seq["inputs_embeds"]).contiguous(), am.to(dev) if not bool(am.all()) else None)
        e["hidden"] = _l2rel(hidden.float().cpu(), it["hidden"])
        feats = dict(res2=g(it["feats"][0]), res3=g(it["feats"][1]), res4=g(it["feats"][2]), res5=g(it["feats"][3]))
        mf, _, ms = model.pixel_decoder.forward_features(feats)
        e["mask_features"] = _l2rel(mf.float().cpu(), it["mask_features"])
        for i in range(3):
            e["ms%d" % i] = _l2rel(ms[i].float().cpu(), it["ms"][i])
        po = model.predictor([g(t) for t in it["ms"]], g(it["mask_features"]), None, g(it["seg_query"]),
                             g(it["SEG_emb"]) if it.get("SEG_emb") is not None else None,
                             g(it["cls_emb"]) if it.get("cls_emb") is not None else None, None)
        ref = it["predictor"]
        e["pred_masks"] = _l2rel(po["pred_masks"].float().cpu(), ref["pred_masks"])
        e["pred_masks_maxrel"] = _maxrel(po["pred_masks"].float().cpu(), ref["pred_masks"])
        if ref.get("pred_class_name_logits") is not None:
            e["class_logits"] = _l2rel(po["pred_class_name_logits"].float().cpu(), ref["pred_class_name_logits"])
        if torch.device(dev).type == "cuda":
            torch.cuda.synchronize()
    return e


def pack_mask(blocked):
    """bool [B,Q,P] (True = blocked) -> (bits int32 [B,Q,ceil(P/32)], row_open uint8 [B,Q]) in the kernels' format."""
    B, Q, P = blocked.shape
    W32 = (P + 31) // 32
    w = torch.zeros(B, Q, W32 * 32, dtype=torch.int64)
    w[..., :P] = blocked.long()
    words = (w.view(B, Q, W32, 32) << torch.arange(32)).sum(-1)
    words = torch.where(words >= 2 ** 31, words - 2 ** 32, words).to(torch.int32)
    return words, blocked.all(-1).to(torch.uint8)


class _LayerTeacher:
    """Layer-wise teacher forcing of the masked decoder: every layer starts from the ORACLE's decoder state and the
    ORACLE's attention mask, so the error of a layer's output is that layer's kernels + storage rounding only; the
    mask-prediction kernels are scored separately as the bit agreement of the mask they derive from the oracle's
    state.  (End to end, one flipped mask bit changes which keys a query attends to - a discontinuity the reference
    has too, see `predictor_sensitivity`.)"""

    def __init__(self, model, it, nh=8):
        self.m, self.po = model, it["predictor"]
        self.sizes = [tuple(t.shape[-2:]) for t in it["ms"]]
        self.nh = nh
        self.layer_l2, self.mask_agree, self.mask_agree_off_threshold = [], [], []

    def _g(self, t):
        return t.to(device=self.m.device, dtype=self.m.dtype)

    def before_layer(self, i, output, bits, row_open, mask_for):
        state = self._g(self.po["layer_outputs"][i].permute(1, 0, 2).contiguous())      # [B,Q,C]
        ref_mask = self.po["trace"][i][3]                                               # bool [B*nh, Q, HW], heads identical
        B = state.shape[0]
        blocked = ref_mask.view(B, self.nh, ref_mask.shape[1], ref_mask.shape[2])[:, 0]
        rb, ro = pack_mask(blocked)
        # mask-prediction kernels on the oracle's state
        pb, _ = mask_for(state)
        flips = (pb.cpu() ^ rb)
        nbits = blocked.numel()
        nflip = sum(int(((flips >> k) & 1).sum()) for k in range(32))
        self.mask_agree.append(1.0 - nflip / nbits)
        # flips away from the decision threshold (|oracle mask logit| at the target size > 5e-2 of its rms)
        lg = torch.nn.functional.interpolate(self.po["trace"][i][2].float(), size=self.sizes[i % 3], mode="bilinear",
                                             align_corners=False).flatten(2)
        hw = blocked.shape[-1]
        firm = lg.abs() > 5e-2 * lg.pow(2).mean().sqrt()
        bit = torch.stack([((flips >> k) & 1) for k in range(32)], -1).view(B, blocked.shape[1], -1)[..., :hw].bool()
        self.mask_agree_off_threshold.append(1.0 - float((bit & firm).sum()) / max(1, int(firm.sum())))
        return state, rb.to(self.m.device), ro.to(self.m.device)

    def after_layer(self, i, output):
        ref = self.po["layer_outputs"][i + 1].permute(1, 0, 2)
        self.layer_l2.append(_l2rel(output.float().cpu(), ref))


def forced_layer_errors(model, it):
    """Per decoder layer: l2-relative error of the layer output and mask-bit agreement, layer-wise teacher forced."""
    dev, dt = model.device, model.dtype
    g = lambda t: t.to(device=dev, dtype=dt)  # noqa: E731
    teacher = _LayerTeacher(model, it)
    pred = model.predictor
    ms = [g(_tok(t)).contiguous() for t in it["ms"]]
    sizes = [tuple(t.shape[-2:]) for t in it["ms"]]
    mf = g(_tok(it["mask_features"])).contiguous()
    with torch.no_grad(), model._precision_scope():
        pred.forward_tokens(ms, sizes, mf, tuple(it["mask_features"].shape[-2:]), g(it["seg_query"]),
                            g(it["SEG_emb"]) if it.get("SEG_emb") is not None else None,
                            g(it["cls_emb"]) if it.get("cls_emb") is not None else None, hooks=teacher)
    e = {"layer_l2rel_max": max(teacher.layer_l2), "layer_l2rel": [round(x, 5) for x in teacher.layer_l2],
         "mask_bit_agree_min": min(teacher.mask_agree)}
    if teacher.mask_agree_off_threshold:
        e["mask_bit_agree_off_threshold_min"] = min(teacher.mask_agree_off_threshold)
    return e


def predictor_sensitivity(sd, it, dtype):
    """Conditioning of the REFERENCE function itself: the oracle's masked decoder re-run (CPU, fp32 arithmetic) on its
    own inputs rounded once to `dtype`, against the un-rounded run.  Thresholded attention masks make the decoder
    discontinuous; on random weights many mask logits sit at the threshold, so a 2^-9 input perturbation can move the
    output by tens of percent.  A GPU run in 16-bit storage cannot be closer to the fp32 oracle than this."""
    from oracle import psalm_oracle as O
    r = lambda t: None if t is None else t.to(dtype).float()  # noqa: E731
    with torch.no_grad():
        po = O.predictor_forward(sd, "predictor.", [r(t) for t in it["ms"]], r(it["mask_features"]), r(it["seg_query"]),
                                 r(it.get("SEG_emb")), r(it.get("cls_emb")))
    ref = it["predictor"]
    e = {"pred_masks_l2rel": _l2rel(po["pred_masks"], ref["pred_masks"]), "pred_masks_maxrel": _maxrel(po["pred_masks"], ref["pred_masks"]),
         "mask_sign_agree": float(((po["pred_masks"] > 0) == (ref["pred_masks"] > 0)).double().mean())}
    if ref.get("pred_class_name_logits") is not None:
        e["class_logits_l2rel"] = _l2rel(po["pred_class_name_logits"], ref["pred_class_name_logits"])
    return e


def fmt(d):
    return ", ".join("%s %s" % (k, ("%.3e" % v) if isinstance(v, float) else v) for k, v in d.items())
