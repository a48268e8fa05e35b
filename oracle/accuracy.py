"""TEST INFRASTRUCTURE — segmentation accuracy of GPU outputs scored against the CPU oracle's outputs on held
inputs (the oracle plays the role of the ground truth; no dataset is reachable offline).

Metric definitions follow what the reference's evaluators feed:
  * PQ / SQ / RQ: panopticapi `pq_compute` as driven by psalm/eval/segmentation_evaluation/
    panoptic_evaluation.py:179-221 (per-image (id map, segments_info) pairs): a predicted and a reference
    segment of the SAME category match when IoU > 0.5; per category PQ = sum(IoU of matches) /
    (TP + FP/2 + FN/2); the reported value is the mean over categories that occur.
  * mIoU / pixel accuracy: the confusion-matrix accumulation of the semantic evaluator
    (panoptic_evaluation.py:136-144: bincount((C + 1) * pred + gt)).
  * mask IoU: mean IoU of the 100 per-query binary masks (logit > 0), matched by query index.
Accumulators are plain int64 / float64 arrays so that ranks can all_reduce(SUM) them.
"""
import numpy as np
import torch


class Accumulator:
    def __init__(self, num_classes):
        self.C = num_classes
        self.pq = np.zeros((num_classes, 4), np.float64)          # iou_sum, tp, fp, fn per category
        self.conf = np.zeros((num_classes + 1, num_classes + 1), np.int64)   # [pred, gt]
        self.misc = np.zeros(6, np.float64)   # pan_agree_px, pan_px, mask_iou_sum, mask_iou_n, images, segments_ref

    # ---- per image ----------------------------------------------------------------------------
    def add_panoptic(self, pred_map, pred_info, ref_map, ref_info):
        p = np.asarray(pred_map).astype(np.int64).ravel()
        g = np.asarray(ref_map).astype(np.int64).ravel()
        self.misc[0] += float((p == g).sum())
        self.misc[1] += p.size
        self.misc[5] += len(ref_info)
        pcat = {int(d["id"]): int(d["category_id"]) for d in pred_info}
        gcat = {int(d["id"]): int(d["category_id"]) for d in ref_info}
        parea = dict(zip(*np.unique(p, return_counts=True)))
        garea = dict(zip(*np.unique(g, return_counts=True)))
        K = int(max(p.max(), g.max())) + 1
        pair, cnt = np.unique(p * K + g, return_counts=True)
        matched_p, matched_g = set(), set()
        for pr, n in zip(pair, cnt):
            pi, gi = int(pr // K), int(pr % K)
            if pi == 0 or gi == 0 or pi not in pcat or gi not in gcat or pcat[pi] != gcat[gi]:
                continue
            union = parea[pi] + garea[gi] - n
            iou = n / union
            if iou > 0.5:
                self.pq[gcat[gi], 0] += iou
                self.pq[gcat[gi], 1] += 1
                matched_p.add(pi)
                matched_g.add(gi)
        for gi, c in gcat.items():
            if gi not in matched_g:
                self.pq[c, 3] += 1
        for pi, c in pcat.items():
            if pi not in matched_p:
                self.pq[c, 2] += 1

    def add_semantic(self, pred_argmax, ref_argmax):
        p = np.asarray(pred_argmax).astype(np.int64).ravel()
        g = np.asarray(ref_argmax).astype(np.int64).ravel()
        self.conf += np.bincount((self.C + 1) * p + g, minlength=self.conf.size).reshape(self.conf.shape)

    def add_masks(self, pred_pos, ref_pos):
        """pred_pos / ref_pos: bool [Q, H, W] (logit > 0)."""
        inter = (pred_pos & ref_pos).flatten(1).sum(1).double()
        union = (pred_pos | ref_pos).flatten(1).sum(1).double()
        ok = union > 0
        self.misc[2] += float((inter[ok] / union[ok]).sum())
        self.misc[3] += float(ok.sum())

    def add_image(self):
        self.misc[4] += 1

    # ---- reduction / report -------------------------------------------------------------------
    def tensors(self):
        return [torch.from_numpy(self.pq), torch.from_numpy(self.conf), torch.from_numpy(self.misc)]

    def load(self, pq, conf, misc):
        self.pq, self.conf, self.misc = pq.cpu().numpy(), conf.cpu().numpy(), misc.cpu().numpy()

    def report(self):
        iou, tp, fp, fn = self.pq.T
        seen = (tp + fp + fn) > 0
        r = {"images": int(self.misc[4])}
        if seen.any():
            den = tp[seen] + 0.5 * fp[seen] + 0.5 * fn[seen]
            r["PQ"] = float(np.mean(iou[seen] / den))
            r["SQ"] = float(np.mean(np.where(tp[seen] > 0, iou[seen] / np.maximum(tp[seen], 1), 0.0)))
            r["RQ"] = float(np.mean(tp[seen] / den))
            r["segments_ref_per_image"] = float(self.misc[5] / max(1.0, self.misc[4]))
            r["pq_categories"] = int(seen.sum())
        if self.misc[1] > 0:
            r["panoptic_pixel_agree"] = float(self.misc[0] / self.misc[1])
        if self.conf.sum() > 0:
            tpc = np.diag(self.conf).astype(np.float64)
            gt = self.conf.sum(0).astype(np.float64)
            pr = self.conf.sum(1).astype(np.float64)
            valid = gt > 0
            r["mIoU"] = float(np.mean(tpc[valid] / np.maximum(gt[valid] + pr[valid] - tpc[valid], 1)))
            r["sem_pixel_acc"] = float(tpc.sum() / self.conf.sum())
            r["sem_classes_ref"] = int(valid.sum())
        if self.misc[3] > 0:
            r["query_mask_mIoU"] = float(self.misc[2] / self.misc[3])
        return r
