"""Deterministic synthetic weights and inputs for the PSALM hot path (no checkpoint / dataset is
available offline).  Shapes follow the reference checkpoint layout (layout.py) and the reference
input contract (train_datasets.py:186-234, collator :968-1044).

Every tensor is drawn from its own CPU generator seeded by crc32(name) ^ seed, so the same weights
can be regenerated anywhere (build container, GPU box) without shipping them.
"""
import math
import zlib

import torch

from .layout import PsalmConfig, checkpoint_layout

# token id used for the "[SEG]" marker that terminates every class name (train_datasets.py:175-184);
# the released tokenizer appends it after the 51200 base vocabulary, synthetic weights keep it in range.
SEG_MARKER_ID = 50295


def _gen(name, seed, device="cpu"):
    g = torch.Generator(device=device)
    g.manual_seed((zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFFFFFFFFFF)
    return g


def relative_position_index(ws):
    """swin_trans.py:93-103 (buffer of the checkpoint)."""
    coords = torch.stack(torch.meshgrid([torch.arange(ws), torch.arange(ws)], indexing="ij"))
    cf = torch.flatten(coords, 1)
    rel = (cf[:, :, None] - cf[:, None, :]).permute(1, 2, 0).contiguous()
    rel[:, :, 0] += ws - 1
    rel[:, :, 1] += ws - 1
    rel[:, :, 0] *= 2 * ws - 1
    return rel.sum(-1)


def synth_tensor(name, shape, dtype, kind, seed=0, window=12, device="cpu"):
    """device="cpu" (default) is bit-reproducible everywhere; device="cuda" draws on the GPU (fast, used
    by bench.py for the 1.59 B-parameter model; values differ from the CPU stream)."""
    g = _gen(name, seed, device)
    _randn, _rand = torch.randn, torch.rand
    if device != "cpu":
        _randn = lambda shape, generator: torch.randn(shape, generator=generator, device=device)  # noqa: E731
        _rand = lambda shape, generator: torch.rand(shape, generator=generator, device=device)  # noqa: E731
    return _synth(name, shape, kind, window, g, _randn, _rand, device)


def _synth(name, shape, kind, window, g, randn, rand, device):
    if kind == "index":
        return relative_position_index(window).to(device)
    if kind == "count":
        return torch.zeros((), dtype=torch.int64, device=device)
    if kind == "w":
        fan_in = max(1, math.prod(shape[1:]))
        std = 1.0 / math.sqrt(fan_in)
        if name.endswith("sampling_offsets.weight"):
            std = 0.02      # reference init is zero (ms_deform_attn.py:67); keep offsets within a few pixels
        elif name.endswith("attention_weights.weight"):
            std = 0.05
        return randn(shape, generator=g) * std
    if kind == "b":
        std = 2.0 if name.endswith("sampling_offsets.bias") else 0.02
        return randn(shape, generator=g) * std
    if kind == "g":
        return 1.0 + 0.1 * randn(shape, generator=g)
    if kind == "emb":
        std = 1.0 if ("level_embed" in name or "query_embed" in name or "query_feat" in name) else 0.05
        return randn(shape, generator=g) * std
    if kind == "table":
        return randn(shape, generator=g) * 0.2
    if kind == "mean":
        return randn(shape, generator=g) * 0.1
    if kind == "var":
        return 0.5 + rand(shape, generator=g)
    raise ValueError(kind)


def synth_state_dict(cfg: PsalmConfig = PsalmConfig(), seed=0, include_lm_head=False, dtype=torch.float32,
                     only_prefix=None, device="cpu"):
    """Synthetic checkpoint with the reference key layout.  lm_head is skipped by default: eval_seg
    never evaluates it (llava_phi.py:1354-1366 takes last_hidden_state)."""
    sd = {}
    for name, (shape, dt, kind) in checkpoint_layout(cfg, include_lm_head=include_lm_head).items():
        if only_prefix is not None and not name.startswith(only_prefix):
            continue
        t = synth_tensor(name, shape, dt, kind, seed, cfg.swin.window, device)
        sd[name] = t if dt == "int64" else t.to(dtype)
    return sd


def synth_inputs(batch=1, height=1024, width=1024, task="panoptic", n_classes=134, seed=1, text_len=(12, 8, 9),
                 refer_len=12, vocab_hi=50000, ragged=False, n_regions=3, visual_prompt_frame=False):
    """Synthetic request in the reference's input contract.

    panoptic / instance / semantic: 12 text ids + <image> + 8 ids + n_classes x (<cls>, one id) + 9 ids +
    <seg> + 1 id; class names of 1-3 random ids each terminated by the [SEG] marker
    (train_datasets.py:175-184, 214-227).  referring: text + <image> + text + <refer> + text + <seg>
    (train_datasets.py:644-).  `ragged=True` varies the text lengths per sample and right-pads
    (collator, train_datasets.py:968-985) so the padding / attention-mask path is exercised."""
    g = torch.Generator().manual_seed(seed)
    images = torch.randn(batch, 3, height, width, generator=g)

    def rnd(n):
        return torch.randint(5, vocab_hi, (n,), generator=g).tolist()

    ids_list = []
    out = {}
    if task in ("panoptic", "instance", "semantic"):
        lens = torch.randint(1, 4, (n_classes,), generator=g).tolist()
        name_ids, cls_idx = [], []
        for c, n in enumerate(lens):
            toks = rnd(n) + [SEG_MARKER_ID]
            name_ids += toks
            cls_idx += [c] * len(toks)
        for b in range(batch):
            extra = int(torch.randint(0, 7, (1,), generator=g)) if (ragged and b > 0) else 0
            ids = rnd(text_len[0] + extra) + [-200] + rnd(text_len[1])
            for _ in range(n_classes):
                ids += [-202] + rnd(1)
            ids += rnd(text_len[2]) + [-201] + rnd(1)
            ids_list.append(ids)
        out["class_name_ids"] = torch.tensor([name_ids] * batch, dtype=torch.long)
        out["cls_indices"] = torch.tensor([cls_idx] * batch, dtype=torch.long)
    elif task == "referring":
        refer = []
        for b in range(batch):
            n = refer_len + (int(torch.randint(0, 5, (1,), generator=g)) if (ragged and b > 0) else 0)
            ids = rnd(text_len[0]) + [-200] + rnd(text_len[1]) + [-204] + rnd(text_len[2]) + [-201] + rnd(1)
            ids_list.append(ids)
            refer.append(torch.tensor(rnd(n) + [SEG_MARKER_ID], dtype=torch.long))
        out["token_refer_id"] = refer
    elif task == "region":   # train_datasets.py:339-341: text <image> text (<region> text) x K text <seg> text
        for b in range(batch):
            k = n_regions + (b if ragged else 0)
            ids = rnd(text_len[0]) + [-200] + rnd(text_len[1])
            for _ in range(k):
                ids += [-203] + rnd(1)
            ids += rnd(text_len[2]) + [-201] + rnd(1)
            ids_list.append(ids)
    else:
        raise ValueError(task)
    T = max(len(i) for i in ids_list)
    pad_id = 0
    input_ids = torch.full((batch, T), pad_id, dtype=torch.long)
    attn = torch.zeros(batch, T, dtype=torch.bool)
    for b, ids in enumerate(ids_list):
        input_ids[b, : len(ids)] = torch.tensor(ids)
        attn[b, : len(ids)] = True
    out["input_ids"] = input_ids
    out["attention_mask"] = attn
    out["images"] = images
    if task in ("panoptic", "instance", "semantic"):
        cei = torch.zeros_like(input_ids)
        cei[input_ids == -202] = 1
        out["class_name_embedding_indices"] = cei
    elif task == "referring":
        rei = torch.zeros_like(input_ids)
        rei[input_ids == -204] = 1
        out["refer_embedding_indices"] = rei
    out["seg_info"] = [dict(padding_mask=torch.zeros(height, width, dtype=torch.bool), height=height, width=width)
                       for _ in range(batch)]
    if task == "region":   # instances.region_masks (BitMasks-like .tensor) and gt_masks, one per <region> token
        from .structures import BitMasks, Instances
        for b in range(batch):
            k = int((input_ids[b] == -203).sum())
            rm = torch.zeros(k, height, width, dtype=torch.bool)
            for j in range(k):   # rectangles of very different areas: < 256 pixels (repeat path) and > 256 (randperm path)
                hh = 3 + 7 * j if j % 2 == 0 else height // 3 + 5 * j
                ww = 4 + 5 * j if j % 2 == 0 else width // 4 + 3 * j
                y0 = int(torch.randint(0, height - hh + 1, (1,), generator=g))
                x0 = int(torch.randint(0, width - ww + 1, (1,), generator=g))
                rm[j, y0:y0 + hh, x0:x0 + ww] = True
            inst = Instances((height, width))
            if visual_prompt_frame:   # DAVIS variant: the masks live on the visual-prompt frame (llava_phi.py:1664)
                inst.vp_region_masks = BitMasks(rm)
            else:
                inst.region_masks = BitMasks(rm)
            inst.gt_masks = rm.float()
            out["seg_info"][b]["instances"] = inst
        if visual_prompt_frame:
            out["vp_images"] = torch.randn(batch, 3, height, width, generator=g)
    if task == "panoptic":
        out["is_thing_list"] = [True] * min(80, n_classes - 1) + [False] * max(0, n_classes - 1 - 80)
    return out
