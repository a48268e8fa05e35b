"""Sequence assembly for the LLM prefill — host logic of the hot path.

Restates `PSALM.prepare_inputs_labels_for_multimodal` / `concat_image_seg_cls_embeds`
(reference language_model/llava_phi.py:767-971, 581-766) and the embedding-extraction helpers
`get_seg_query` (:1299-1316), `get_class_name_embedding` (:552-565), `get_SEG_embedding` (:972-978).

The reference walks every token id in Python with `.item()` (llava_phi.py:614-623) and launches
~270 tiny kernels for class-name pooling.  Here the prompt is turned ONCE, on the host, into an index
*plan* (numpy, loops only over the ~140 sentinel ids); the device side is then four library calls:
one embedding gather, two row scatters (image tokens, seg queries) and one pooling matmul.
Sentinel ids (psalm/constants.py:8-12): <image> -200, <seg> -201, <cls> -202, <region> -203, <refer> -204.
"""
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

IMAGE_TOKEN_INDEX = -200
SEG_TOKEN_INDEX = -201
CLS_TOKEN_INDEX = -202
REGION_TOKEN_INDEX = -203
REFER_TOKEN_INDEX = -204


@dataclass
class SequencePlan:
    B: int
    T: int
    tok_ids: torch.Tensor            # [B,T] int64  token id to embed (0 where the row comes from elsewhere)
    img_pos: torch.Tensor            # [B*n_img] int64 flat row index (b*T + t) of every image token
    seg_pos: torch.Tensor            # [B*n_q]  int64 flat row index of every seg-query token
    pad_pos: Optional[torch.Tensor]  # flat rows that are right padding (zero embeddings), or None
    attention_mask: torch.Tensor     # [B,T] bool
    any_padding: bool
    cls_pool: Optional[torch.Tensor]   # [B,n_cls,T] fp32 averaging matrix (rows sum to 1), or None
    refer_pool: Optional[torch.Tensor]  # [B,1,T] fp32, or None
    n_img: int
    n_q: int
    region_pos: Optional[torch.Tensor] = None   # [R] int64 flat rows of the <region> tokens (sample-major, prompt order)
    region_counts: Optional[tuple] = None        # regions per sample
    region_points: Optional[torch.Tensor] = None  # [R,P,2] fp32 sample points (y, x) in [0,1], set by the caller
    region_image: Optional[torch.Tensor] = None   # [R] int32 image index of every region
    vp_images: Optional[torch.Tensor] = None      # [B,3,H,W] visual-prompt frames the regions are pooled from (DAVIS variant)

    def to(self, device):
        mv = lambda t: None if t is None else t.to(device, non_blocking=True)  # noqa: E731
        return SequencePlan(self.B, self.T, mv(self.tok_ids), mv(self.img_pos), mv(self.seg_pos), mv(self.pad_pos),
                            mv(self.attention_mask), self.any_padding, mv(self.cls_pool), mv(self.refer_pool),
                            self.n_img, self.n_q, mv(self.region_pos), self.region_counts, mv(self.region_points),
                            mv(self.region_image), mv(self.vp_images))


def build_plan(input_ids, attention_mask, n_img, n_q, class_name_ids=None, cls_indices=None,
               class_name_embedding_indices=None, token_refer_id=None, refer_embedding_indices=None):
    """All arguments are HOST tensors in the reference's input contract (train_datasets.py:186-234)."""
    ids_all = input_ids.cpu().numpy()
    B, T0 = ids_all.shape
    am_all = np.ones((B, T0), bool) if attention_mask is None else attention_mask.cpu().numpy().astype(bool)
    has_region = bool((ids_all == REGION_TOKEN_INDEX).any())
    rows = []
    for b in range(B):
        ids = ids_all[b]
        assert (ids == IMAGE_TOKEN_INDEX).sum() == 1, "not supporting multi image index"   # llava_phi.py:588
        assert (ids == SEG_TOKEN_INDEX).sum() == 1, "not supporting multi seg index"       # llava_phi.py:589
        names = None
        if class_name_ids is not None:  # embed_class_ids, llava_phi.py:566-575
            cn = class_name_ids[b].cpu().numpy()
            ci = cls_indices[b].cpu().numpy()
            names = []
            prev = None
            for u in ci:  # unique_consecutive, then drop the -1 padding
                if u != prev:
                    prev = u
                    if u >= 0:
                        names.append(cn[ci == u])
            assert (ids == CLS_TOKEN_INDEX).sum() == len(names), \
                "the number of <cls> tokens and class_embed needs to be same"               # llava_phi.py:590-591
        refer = token_refer_id[b].cpu().numpy() if token_refer_id is not None else None
        cei = class_name_embedding_indices[b].cpu().numpy() if class_name_embedding_indices is not None else None
        rei = refer_embedding_indices[b].cpu().numpy() if refer_embedding_indices is not None else None
        tok, kind, cidx, ridx = [], [], [], []   # kind: 1 text, 2 image, 3 seg, 4 region feature
        special = np.nonzero(ids < 0)[0]
        prev_end = 0
        cls_i = 0

        def text(lo, hi):
            if hi > lo:
                tok.append(ids[lo:hi]); kind.append(np.full(hi - lo, 1))
                cidx.append(cei[lo:hi] if cei is not None else np.zeros(hi - lo, np.int64))
                ridx.append(rei[lo:hi] if rei is not None else np.zeros(hi - lo, np.int64))

        for sp in special:
            text(prev_end, sp)
            t = ids[sp]
            if t == IMAGE_TOKEN_INDEX:
                n, k, tk, cv, rv = n_img, 2, np.zeros(n_img, np.int64), 0, 0
            elif t == SEG_TOKEN_INDEX:
                n, k, tk, cv, rv = n_q, 3, np.zeros(n_q, np.int64), 0, 0
            elif t == CLS_TOKEN_INDEX:
                tk = names[cls_i]
                cls_i += 1
                n, k, cv, rv = len(tk), 1, cls_i, 0   # 1-based running class counter (llava_phi.py:671-673)
            elif t == REFER_TOKEN_INDEX:
                tk = refer
                n, k, cv, rv = len(tk), 1, 0, 1
            elif t == REGION_TOKEN_INDEX:   # one row per region, filled with its pooled feature (llava_phi.py:684-703)
                n, k, tk, cv, rv = 1, 4, np.zeros(1, np.int64), 0, 0
            else:
                raise ValueError("unknown sentinel id %d" % t)
            tok.append(tk); kind.append(np.full(n, k)); cidx.append(np.full(n, cv)); ridx.append(np.full(n, rv))
            prev_end = sp + 1
        text(prev_end, T0)
        tok, kind, cidx, ridx = map(np.concatenate, (tok, kind, cidx, ridx))
        # attention mask: inserted tokens attendable, then the caller's mask (llava_phi.py:935-949, 964-969)
        am = np.concatenate([np.ones(len(tok) - T0, bool), am_all[b]])
        rows.append((tok, kind, cidx, ridx, am))
    T = max(len(r[0]) for r in rows)
    tok_ids = np.zeros((B, T), np.int64)
    attn = np.zeros((B, T), bool)
    cls_idx = np.zeros((B, T), np.int64)
    ref_idx = np.zeros((B, T), np.int64)
    img_pos, seg_pos, pad_pos, region_pos, region_counts = [], [], [], [], []
    for b, (tok, kind, cidx, ridx, am) in enumerate(rows):
        n = len(tok)
        tok_ids[b, :n] = np.where(kind == 1, tok, 0)
        attn[b, :n] = am
        cls_idx[b, :n] = cidx
        ref_idx[b, :n] = ridx
        img_pos.append(b * T + np.nonzero(kind == 2)[0])
        seg_pos.append(b * T + np.nonzero(kind == 3)[0])
        region_pos.append(b * T + np.nonzero(kind == 4)[0])
        region_counts.append(int((kind == 4).sum()))
        pad_pos.append(b * T + np.arange(n, T))
    pad_pos = np.concatenate(pad_pos)
    cls_pool = refer_pool = None
    if class_name_embedding_indices is not None:
        ncls = int(cls_idx.max())
        cls_pool = np.zeros((B, ncls, T), np.float32)
        for b in range(B):
            for c in range(1, ncls + 1):
                m = cls_idx[b] == c
                assert m.any(), "class %d has no tokens in sample %d" % (c, b)
                cls_pool[b, c - 1, m] = 1.0 / m.sum()   # AdaptiveAvgPool1d(1), llava_phi.py:561-563
    if refer_embedding_indices is not None:
        refer_pool = np.zeros((B, 1, T), np.float32)
        for b in range(B):
            m = ref_idx[b] != 0
            refer_pool[b, 0, m] = 1.0 / m.sum()
    any_padding = bool((~attn).any())
    ft = torch.from_numpy
    return SequencePlan(B, T, ft(tok_ids), ft(np.concatenate(img_pos)), ft(np.concatenate(seg_pos)),
                        ft(pad_pos) if len(pad_pos) else None, ft(attn), any_padding,
                        ft(cls_pool) if cls_pool is not None else None,
                        ft(refer_pool) if refer_pool is not None else None, n_img, n_q,
                        ft(np.concatenate(region_pos)) if has_region else None,
                        tuple(region_counts) if has_region else None)


def materialize_embeds(plan, embed_tokens, image_tokens, seg_query, region_features=None):
    """plan on device; embed_tokens [V,C]; image_tokens [B,n_img,C]; seg_query [n_q,C]; region_features [R,C] (one
    row per <region> token, plan.region_pos order) -> inputs_embeds [B,T,C]."""
    B, T = plan.B, plan.T
    C = embed_tokens.shape[1]
    flat = embed_tokens.index_select(0, plan.tok_ids.view(-1))
    flat.index_copy_(0, plan.img_pos, image_tokens.reshape(-1, C).to(flat.dtype))
    flat.index_copy_(0, plan.seg_pos, seg_query.to(flat.dtype).repeat(B, 1))
    if plan.region_pos is not None:
        flat.index_copy_(0, plan.region_pos, region_features.to(flat.dtype))
    if plan.pad_pos is not None:
        flat.index_fill_(0, plan.pad_pos, 0)   # right padding rows are zeros (llava_phi.py:878-883)
    return flat.view(B, T, C)


def gather_seg_query(plan, hidden):
    """hidden [B,T,C] -> [B,n_q,C]  (get_seg_query, llava_phi.py:1299-1316)."""
    return hidden.reshape(plan.B * plan.T, -1).index_select(0, plan.seg_pos).view(plan.B, plan.n_q, -1)


def pool(pool_matrix, hidden):
    """[B,n,T] averaging matrix x [B,T,C] -> [B,n,C] (class-name / [SEG] mean pooling)."""
    return torch.bmm(pool_matrix.to(hidden.dtype), hidden)


def gather_region_rows(plan, hidden):
    """hidden [B,T,C] -> [R,C] hidden states at the <region> positions (get_region_embedding, llava_phi.py:302-307)."""
    return hidden.reshape(plan.B * plan.T, -1).index_select(0, plan.region_pos)
