"""Post-processing of eval_seg (reference language_model/llava_phi.py:1395-1472 and the task heads
:308-447), on the GPU.  detectron2's `sem_seg_postprocess` / `ImageList.from_tensors` (third-party,
un-vendored, version unpinned upstream) are restated: crop to the un-padded box, bilinear resize.

Panoptic merging: the reference loops over kept segments in Python with several `.item()` syncs per
segment (:356-384).  Here the per-segment areas are three batched reductions, ONE small D2H copy
brings (keep, class, areas) to the host, the sequential merge rule (stuff classes share an id, overlap
threshold 0.8) runs on <= 100 integers, and a lookup table maps the arg-max map to segment ids on the
device."""
import numpy as np
import torch
import torch.nn.functional as F

from .structures import Boxes, Instances


def unpadded_box(padding_mask):
    """llava_phi.py:1418-1423: bounding box of ~padding_mask (host side, like the reference)."""
    pm = padding_mask.cpu().numpy() if isinstance(padding_mask, torch.Tensor) else np.asarray(padding_mask)
    nz = np.where(~pm.astype(bool))
    return int(nz[0].max() - nz[0].min() + 1), int(nz[1].max() - nz[1].min() + 1)


def sem_seg_postprocess(result, img_size, out_h, out_w):
    result = result[:, : img_size[0], : img_size[1]].unsqueeze(0)
    if result.shape[-2:] == (out_h, out_w):
        return result[0]
    return F.interpolate(result, size=(out_h, out_w), mode="bilinear", align_corners=False)[0]


def semantic_inference(cls, mask_pred):
    """softmax(cls)[:, :-1]^T . sigmoid(mask)  (llava_phi.py:402-406) as one library GEMM."""
    probs = F.softmax(cls.float(), dim=-1)[:, :-1]
    Q, H, W = mask_pred.shape
    return torch.matmul(probs.t(), mask_pred.float().sigmoid().view(Q, H * W)).view(-1, H, W)


def _mask_scores(mp):
    pm = (mp > 0).float()
    return pm, (mp.sigmoid().flatten(1) * pm.flatten(1)).sum(1) / (pm.flatten(1).sum(1) + 1e-6)


def instance_inference(cls, mask_pred, topk, is_thing_list=None, panoptic_on=False):
    """llava_phi.py:407-447 (topk(sorted=False): order is implementation defined; compare as sets)."""
    scores = F.softmax(cls.float(), dim=-1)[:, :-1]
    nq, nc = scores.shape
    s, idx = scores.flatten(0, 1).topk(topk, sorted=False)
    lab = idx % nc
    qi = idx // nc
    if panoptic_on:
        thing = torch.as_tensor([bool(t) for t in is_thing_list], device=cls.device)
        keep = thing[lab]
        s, lab, qi = s[keep], lab[keep], qi[keep]
    mp = mask_pred[qi].float()
    pm, ms = _mask_scores(mp)
    r = Instances(tuple(mask_pred.shape[-2:]))
    r.pred_masks = pm
    r.pred_boxes = Boxes(torch.zeros(mp.size(0), 4))
    r.scores = s * ms
    r.pred_classes = lab
    r.query_index = qi
    return r


def seg_instance_inference(SEG_cls, mask_pred, topk):
    """llava_phi.py:308-324 (referring segmentation)."""
    scores = torch.sigmoid(SEG_cls.float())
    s, idx = scores.flatten(0, 1).topk(topk, sorted=False)
    mp = mask_pred[idx].float()
    pm, ms = _mask_scores(mp)
    r = Instances(tuple(mask_pred.shape[-2:]))
    r.pred_masks = pm
    r.pred_boxes = Boxes(torch.zeros(mp.size(0), 4))
    r.scores = s * ms
    r.query_index = idx
    return r


def panoptic_inference(cls, mask_pred, is_thing_list, obj_thr=0.8, ovl_thr=0.8):
    """llava_phi.py:325-386 -> (panoptic_seg int32 [H,W], segments_info list)."""
    scores, labels = F.softmax(cls.float(), dim=-1).max(-1)
    nc = cls.shape[-1] - 1
    Q, H, W = mask_pred.shape
    sig = mask_pred.float().sigmoid()
    keep = labels.ne(nc) & (scores > obj_thr)
    # argmax over kept queries only: non-kept rows get -1 (< any kept score*sigmoid >= 0)
    prob = torch.where(keep.view(-1, 1, 1), scores.view(-1, 1, 1) * sig, torch.full_like(sig, -1.0))
    ids = prob.argmax(0)                                   # [H,W] query index
    ge = sig >= 0.5
    onehot = F.one_hot(ids.view(-1), Q).t().view(Q, H, W).bool()
    area = onehot.flatten(1).sum(1)
    orig = ge.flatten(1).sum(1)
    inter = (onehot & ge).flatten(1).sum(1)
    host = torch.stack([keep.long(), labels.long(), area, orig, inter], 0).cpu().numpy()   # the one D2H copy
    seg_of_query = np.zeros(Q + 1, np.int32)
    info, stuff, cur = [], {}, 0
    if host[0].sum() == 0:
        return torch.zeros((H, W), dtype=torch.int32, device=cls.device), info
    for q in range(Q):
        if not host[0, q]:
            continue
        pc, a, o, it = int(host[1, q]), int(host[2, q]), int(host[3, q]), int(host[4, q])
        if a > 0 and o > 0 and it > 0:
            if a / o < ovl_thr:
                continue
            isthing = bool(is_thing_list[pc])
            if not isthing:
                if pc in stuff:
                    seg_of_query[q] = stuff[pc]
                    continue
                stuff[pc] = cur + 1
            cur += 1
            seg_of_query[q] = cur
            info.append(dict(id=cur, isthing=isthing, category_id=pc))
    lut = torch.from_numpy(seg_of_query).to(cls.device)
    in_mask = torch.gather(ge.view(Q, -1), 0, ids.view(1, -1)).view(H, W)   # sig[ids[p], p] >= 0.5
    pan = torch.where(in_mask, lut[ids], torch.zeros((), dtype=torch.int32, device=cls.device))
    return pan.to(torch.int32), info
