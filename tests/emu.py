"""TEST INFRASTRUCTURE — CPU emulation of the CUDA entry points in psalm_b200.kernels, built from
plain torch ops, so the *host orchestration* (weight preparation, layouts, sequence plans, decoder
control flow, post-processing) can be checked against the oracle on a machine without a GPU.
It lives under tests/ and is installed by monkeypatching inside tests only; the product has no such
switch and refuses to run without CUDA (tests/test_abi.py::test_no_cpu_path)."""
import math

import torch
import torch.nn.functional as F

from oracle import psalm_oracle as O


def window_attention(qkv, qkv_bias, rel_bias, B, H, W, C, nh, ws, shift):
    hd = C // nh
    Hp, Wp = -(-H // ws) * ws, -(-W // ws) * ws
    x = qkv_bias.view(1, 1, 1, 3 * C).expand(B, Hp, Wp, 3 * C).clone()
    x[:, :H, :W] = qkv.view(B, H, W, 3 * C)
    if shift:
        x = torch.roll(x, (-shift, -shift), (1, 2))
    xw = O._window_partition(x, ws).view(-1, ws * ws, 3, nh, hd).permute(2, 0, 3, 1, 4).float()
    q, k, v = xw[0], xw[1], xw[2]
    # rel_bias is the compact [nh, (2 ws - 1)^2] table; expand it the way the reference does (swin_trans.py:131-134)
    dense = rel_bias[:, O.relative_position_index(ws).view(-1)].view(nh, ws * ws, ws * ws)
    attn = (q @ k.transpose(-2, -1)) * hd ** -0.5 + dense.unsqueeze(0)
    if shift:
        m = O._shift_mask(H, W, ws, shift)
        nW = m.shape[0]
        attn = (attn.view(B, nW, nh, ws * ws, ws * ws) + m.unsqueeze(1).unsqueeze(0)).view(-1, nh, ws * ws, ws * ws)
    o = (attn.softmax(-1) @ v).transpose(1, 2).reshape(-1, ws, ws, C)
    o = O._window_reverse(o, ws, Hp, Wp)
    if shift:
        o = torch.roll(o, (shift, shift), (1, 2))
    return o[:, :H, :W].reshape(B, H * W, C).to(qkv.dtype)


def rotary_inplace(qkv, cos, sin, B, T, nh, hd, rd):
    half = rd // 2
    for which in (0, 1):
        x = qkv[:, :, which].float()
        x1, x2 = x[..., :half].clone(), x[..., half:rd].clone()
        c, s = cos[None, :, None, :], sin[None, :, None, :]
        qkv[:, :, which, :, :half] = (x1 * c - x2 * s).to(qkv.dtype)
        qkv[:, :, which, :, half:rd] = (x2 * c + x1 * s).to(qkv.dtype)


def causal_attention(qkv, key_valid, B, T, nh, hd):
    q, k, v = (qkv[:, :, i].permute(0, 2, 1, 3).float() for i in range(3))
    s = (q @ k.transpose(-2, -1)) * hd ** -0.5
    allowed = torch.tril(torch.ones(T, T, dtype=torch.bool))[None, None]
    if key_valid is not None:
        allowed = allowed & key_valid.bool()[:, None, None, :]
    s = s.masked_fill(~allowed, float("-inf"))
    p = torch.nan_to_num(s.softmax(-1))
    return (p @ v).permute(0, 2, 1, 3).reshape(B, T, nh * hd).to(qkv.dtype)


def cross_attention(q, k, v, mask_bits=None, row_open=None, nh=8, splits=None, workspace=None):
    B, Lq, C = q.shape
    Lk = k.shape[1]
    hd = C // nh
    qh, kh, vh = (t.view(B, -1, nh, hd).permute(0, 2, 1, 3).float() for t in (q, k, v))
    s = (qh @ kh.transpose(-2, -1)) / math.sqrt(hd)
    if mask_bits is not None:
        w = mask_bits.long() & 0xFFFFFFFF
        bit = (w.unsqueeze(-1) >> torch.arange(32)) & 1
        blocked = bit.view(B, Lq, -1)[..., :Lk].bool()
        if row_open is not None:
            blocked = blocked & ~row_open.bool().unsqueeze(-1)
        s = s.masked_fill(blocked.unsqueeze(1), float("-inf"))
    return (s.softmax(-1) @ vh).permute(0, 2, 1, 3).reshape(B, Lq, C).to(q.dtype)


def masked_cross_attention(q, k, v, mask_bits=None, row_open=None, nh=8, workspace=None):
    return cross_attention(q, k, v, mask_bits, row_open, nh)


def mask_logits(mask_embed, feats, out_dtype=None):
    return torch.bmm(mask_embed.float(), feats.float().transpose(1, 2)).to(out_dtype or mask_embed.dtype)


def bilinear_tokens(x, Hi, Wi, Ho, Wo, out=None, out_dtype=None, accumulate=False):
    B, _, C = x.shape
    y = F.interpolate(x.float().view(B, Hi, Wi, C).permute(0, 3, 1, 2), size=(Ho, Wo), mode="bilinear",
                      align_corners=False).permute(0, 2, 3, 1).reshape(B, Ho * Wo, C)
    if out is None:
        return y.to(out_dtype or x.dtype)
    out.copy_((out.float() + y if accumulate else y).to(out.dtype))
    return out


def attn_mask_bits(logits):
    B, Q, P = logits.shape
    W32 = (P + 31) // 32
    blocked = torch.zeros(B, Q, W32 * 32, dtype=torch.int64)
    blocked[..., :P] = (logits.float() < 0).long()
    words = (blocked.view(B, Q, W32, 32) << torch.arange(32)).sum(-1)
    words = torch.where(words >= 2 ** 31, words - 2 ** 32, words).to(torch.int32)
    row_open = (logits.float() < 0).all(-1).to(torch.uint8)
    return words, row_open


def msda_encoder_fused(value_hm, ow, spatial_shapes, level_start_index, n_points=4):
    B, M, S, D = value_hm.shape
    L = len(spatial_shapes)
    LP = L * n_points
    off = ow[..., : M * LP * 2].float().view(B, S, M, L, n_points, 2)
    aw = ow[..., M * LP * 2:].float().view(B, S, M, LP).softmax(-1).view(B, S, M, L, n_points)
    ref = O.encoder_reference_points(spatial_shapes, B)
    norm = torch.tensor([[w, h] for h, w in spatial_shapes], dtype=torch.float32)
    loc = ref[:, :, None, :, None, :] + off / norm[None, None, None, :, None, :]
    return O.msda_core(value_hm.permute(0, 2, 1, 3).float(), spatial_shapes, loc, aw).to(value_hm.dtype)


def add_layer_norm(x, weight, bias, eps=1e-5, r1=None, r2=None, return_sum=False):
    s = x.float()
    if r1 is not None:
        s = s + r1.float()
    if r2 is not None:
        s = s + r2.float()
    if return_sum:
        s = s.to(x.dtype).float()
    y = F.layer_norm(s, (x.shape[-1],), weight.float(), bias.float(), eps).to(x.dtype)
    return (s.to(x.dtype), y) if return_sum else y


def group_norm_tokens(x, weight, bias, groups=32, eps=1e-5, relu=False, pre_bias=None):
    xf = x.float() if pre_bias is None else x.float() + pre_bias.float()
    y = F.group_norm(xf.transpose(1, 2), groups, weight.float(), bias.float(), eps).transpose(1, 2)
    if relu:
        y = F.relu(y)
    return y.contiguous().to(x.dtype)


def mask_bits(mask_embed, feats):
    return attn_mask_bits(mask_logits(mask_embed, feats, torch.float32))


def patchify(images, out_dtype, mean=None, std=None, patch=4):
    x = images.float()
    if mean is not None:
        x = (x - mean.view(1, -1, 1, 1)) / std.view(1, -1, 1, 1)
    B, C, H, W = x.shape
    x = F.pad(x, (0, (-W) % patch, 0, (-H) % patch))
    Wh, Ww = x.shape[2] // patch, x.shape[3] // patch
    x = x.view(B, C, Wh, patch, Ww, patch).permute(0, 2, 4, 1, 3, 5).reshape(B, Wh * Ww, C * patch * patch)
    return x.to(out_dtype), (Wh, Ww)


def kv_cache_write(qkv, kcache, vcache, block_table, start_pos):
    B, T, _, nh, hd = qkv.shape
    ps = kcache.shape[2]
    for b in range(B):
        for t in range(T):
            pos = int(start_pos[b]) + t
            page, slot = int(block_table[b, pos // ps]), pos % ps
            kcache[page, :, slot] = qkv[b, t, 1]
            vcache[page, :, slot] = qkv[b, t, 2]


def paged_decode_attention(qkv, kcache, vcache, block_table, seq_lens):
    B, _, _, nh, hd = qkv.shape
    ps = kcache.shape[2]
    out = torch.empty(B, 1, nh * hd, dtype=qkv.dtype)
    for b in range(B):
        n = int(seq_lens[b])
        pages = block_table[b, : (n + ps - 1) // ps].long()
        k = kcache[pages].permute(1, 0, 2, 3).reshape(nh, -1, hd)[:, :n].float()
        v = vcache[pages].permute(1, 0, 2, 3).reshape(nh, -1, hd)[:, :n].float()
        s = torch.einsum("hd,hkd->hk", qkv[b, 0, 0].float(), k) * hd ** -0.5
        out[b, 0] = torch.einsum("hk,hkd->hd", s.softmax(-1), v).reshape(-1).to(qkv.dtype)
    return out


def install(monkeypatch):
    from psalm_b200 import kernels
    for name in ("window_attention", "rotary_inplace", "causal_attention", "cross_attention", "mask_logits",
                 "bilinear_tokens", "attn_mask_bits", "msda_encoder_fused", "add_layer_norm", "group_norm_tokens", "mask_bits", "patchify", "masked_cross_attention", "kv_cache_write",
                 "paged_decode_attention", "linear_fused_supported", "patch_merge_layer_norm", "region_pool"):
        monkeypatch.setattr(kernels, name, globals()[name])


def postproc_crop_supported(*a):
    return True


def postproc_fused(logits, H, W, probsT=None, wq=None, negq=None, slot_query=None, ncls=0, crop=None):
    """torch restatement of csrc/postproc.cu (one tile = the whole image)."""
    Q = logits.shape[0]
    if crop is not None:   # up-sample to the padded size, crop, resize (sem_seg_postprocess)
        Hp, Wp, oh, ow = crop
        x = F.interpolate(logits.float()[None], size=(Hp, Wp), mode="bilinear", align_corners=False)[:, :, :oh, :ow]
        x = F.interpolate(x, size=(H, W), mode="bilinear", align_corners=False)[0]
    else:
        x = F.interpolate(logits.float()[None], size=(H, W), mode="bilinear", align_corners=False)[0]
    s = torch.sigmoid(x)
    out = dict(sem_seg=None, ids=None, in_mask=None, inst_masks=None)
    stats = torch.zeros(Q, 5)
    stats[:, 0] = (x > 0).flatten(1).sum(1)
    stats[:, 1] = (s * (x > 0)).flatten(1).sum(1)
    stats[:, 2] = (x >= 0).flatten(1).sum(1)
    if probsT is not None:
        out["sem_seg"] = torch.einsum("cq,qhw->chw", probsT[:ncls, :Q].float(), s.half().float())
    if wq is not None:
        ids = (wq.view(-1, 1, 1) * s + negq.view(-1, 1, 1)).argmax(0)
        inm = x.flatten(1).gather(0, ids.view(1, -1)).view(H, W) >= 0
        stats[:, 3] = torch.bincount(ids.view(-1), minlength=Q)
        stats[:, 4] = torch.bincount(ids.view(-1), weights=inm.view(-1).float(), minlength=Q)
        out["ids"], out["in_mask"] = ids.to(torch.int32), inm.to(torch.uint8)
    if slot_query is not None:
        m = torch.zeros(slot_query.shape[0], H, W)
        for k, q in enumerate(slot_query.tolist()):
            if q >= 0:
                m[k] = (x[q] > 0).float()
        out["inst_masks"] = m
    out["stats"] = stats
    return out


def linear_fused_supported(*a, **k):
    return False


def patch_merge_layer_norm(x, H, W, weight, bias, eps=1e-5):
    B, N, C = x.shape
    xm = x.view(B, H, W, C)
    if H % 2 == 1 or W % 2 == 1:
        xm = F.pad(xm, (0, 0, 0, W % 2, 0, H % 2))
    xm = torch.cat([xm[:, 0::2, 0::2], xm[:, 1::2, 0::2], xm[:, 0::2, 1::2], xm[:, 1::2, 1::2]], -1)
    xm = xm.reshape(B, -1, 4 * C)
    return add_layer_norm(xm.contiguous(), weight, bias, eps)


def region_pool(tokens, points, region_image, h, w):
    outs = []
    for r in range(points.shape[0]):
        fmap = tokens[int(region_image[r])].float().view(h, w, -1).permute(2, 0, 1)[None]
        grid = (2.0 * points[r:r + 1].flip(dims=(2,)) - 1.0).unsqueeze(2)
        fea = F.grid_sample(fmap, grid.float(), align_corners=True).squeeze(3)        # [1,C,P]
        outs.append(fea.mean(-1))
    return torch.cat(outs, 0).to(tokens.dtype)
