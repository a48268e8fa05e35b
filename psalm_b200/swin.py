"""Swin-B vision tower, token-major, for the PSALM hot path.

Mirrors `SwinTransformer.forward` (reference psalm/model/multimodal_encoder/swin_trans.py:608-633):
four stage outputs (strides 4/8/16/32), each through its own LayerNorm.  Differences in *how*:
  * activations stay token-major [B, H*W, C]; the NCHW tuple the reference returns is a strided view;
  * W-MSA / SW-MSA is ONE fused kernel per block (psalm_window_attention): cyclic shift, window
    partition, the zero padding after norm1 (swin_trans.py:207-214), the relative-position bias gather
    (:131-134) and the -100 shift mask (:370-387) are folded into the kernel's addressing, so the
    reference's roll / pad / partition / reverse copies do not exist;
  * the relative-position bias stays the compact (2 ws - 1)^2 table of the checkpoint (transposed, fp32).
Linear layers are library GEMMs (cuBLAS through torch)."""
import torch
import torch.nn.functional as F

from . import kernels
from .layout import SwinConfig


_REL_INDEX = {}


def _rel_index(ws):
    """relative_position_index of a ws x ws window (swin_trans.py:93-103)."""
    if ws not in _REL_INDEX:
        coords = torch.stack(torch.meshgrid([torch.arange(ws), torch.arange(ws)], indexing="ij")).flatten(1)
        rel = (coords[:, :, None] - coords[:, None, :]).permute(1, 2, 0).contiguous()
        rel[:, :, 0] += ws - 1
        rel[:, :, 1] += ws - 1
        rel[:, :, 0] *= 2 * ws - 1
        _REL_INDEX[ws] = rel.sum(-1)
    return _REL_INDEX[ws]


class SwinTransformer:
    def __init__(self, sd, prefix="model.vision_tower.", cfg=SwinConfig(), dtype=torch.bfloat16, device="cuda"):
        self.cfg, self.dtype, self.device = cfg, dtype, device
        from .image_processor import build_image_processors
        self.image_processor = build_image_processors()   # panoptic | instance | semantic (llava_phi.py:66-69)
        # pixel statistics of the reference's dataset mappers (datasets_mapper/coco_panoptic_mapper.py:118-119):
        # uint8 images are normalised on the device with them
        self.pixel_mean = torch.tensor([123.675, 116.28, 103.53], dtype=torch.float32, device=device)
        self.pixel_std = torch.tensor([58.395, 57.12, 57.375], dtype=torch.float32, device=device)
        self.num_features = [cfg.embed_dim * 2 ** i for i in range(len(cfg.depths))]
        g = lambda k: sd[prefix + k]  # noqa: E731
        cv = lambda t: t.to(device=device, dtype=dtype).contiguous()  # noqa: E731
        w = {}
        w["pe.w2"] = cv(g("patch_embed.proj.weight").reshape(g("patch_embed.proj.weight").shape[0], -1))   # [C, 3*ps*ps]
        w["pe.b"] = cv(g("patch_embed.proj.bias"))
        w["pe.nw"], w["pe.nb"] = cv(g("patch_embed.norm.weight")), cv(g("patch_embed.norm.bias"))
        ws = cfg.window
        for s, depth in enumerate(cfg.depths):
            nh = cfg.num_heads[s]
            for b in range(depth):
                p = "layers.%d.blocks.%d." % (s, b)
                for n in ("norm1", "norm2", "attn.qkv", "attn.proj", "mlp.fc1", "mlp.fc2"):
                    w[p + n + ".w"] = cv(g(p + n + ".weight"))
                    w[p + n + ".b"] = cv(g(p + n + ".bias"))
                # compact table [nh, (2 ws - 1)^2] fp32; the kernels index it like relative_position_index does
                # (swin_trans.py:93-103, a deterministic buffer: it is validated if the checkpoint carries it,
                # never needed)
                tab = g(p + "attn.relative_position_bias_table").float()
                idx = sd.get(prefix + p + "attn.relative_position_index")
                if idx is not None and not torch.equal(idx.cpu().long(), _rel_index(ws)):
                    raise ValueError("%sattn.relative_position_index differs from swin_trans.py:93-103" % (prefix + p))
                w[p + "rel"] = tab.t().contiguous().to(device)
            if s < len(cfg.depths) - 1:
                p = "layers.%d.downsample." % s
                w[p + "red.w"] = cv(g(p + "reduction.weight"))
                w[p + "norm.w"], w[p + "norm.b"] = cv(g(p + "norm.weight")), cv(g(p + "norm.bias"))
            w["norm%d.w" % s], w["norm%d.b" % s] = cv(g("norm%d.weight" % s)), cv(g("norm%d.bias" % s))
        self.w = w

    # -- public surface of the reference -----------------------------------------------------------
    def __call__(self, images):
        return self.forward(images)

    def forward(self, images):
        """images [B,3,H,W] -> tuple of 4 NCHW maps (strided views of the token-major results)."""
        toks, sizes = self.forward_tokens(images)
        return tuple(t.view(t.shape[0], h, w_, t.shape[2]).permute(0, 3, 1, 2) for t, (h, w_) in zip(toks, sizes))

    # -- token-major pipeline ----------------------------------------------------------------------
    def forward_tokens(self, images):
        cfg, w = self.cfg, self.w
        H, W = images.shape[-2:]
        ps = cfg.patch
        # One kernel: (uint8 input) pixel normalisation of the mapper (coco_panoptic_mapper.py:161), zero padding to
        # the patch grid (swin_trans.py:431-434) and the unfold of the stride-ps convolution (:427-441), which is
        # then a GEMM over the unfolded patches with the bias in its epilogue.
        img = images.to(device=self.device).contiguous()
        norm = (self.pixel_mean, self.pixel_std) if img.dtype == torch.uint8 else (None, None)
        patches, (Wh, Ww) = kernels.patchify(img, self.dtype, norm[0], norm[1], ps)
        B = img.shape[0]
        x = F.linear(patches, w["pe.w2"], w["pe.b"])
        C = x.shape[-1]
        x = kernels.add_layer_norm(x.contiguous(), w["pe.nw"], w["pe.nb"])
        outs, sizes = [], []
        ws = cfg.window
        LN = kernels.add_layer_norm
        for s, depth in enumerate(cfg.depths):
            nh = cfg.num_heads[s]
            delta = None   # residual branch not yet added to x (folded into the next fused add + LayerNorm)
            for b in range(depth):
                p = "layers.%d.blocks.%d." % (s, b)
                shift = 0 if b % 2 == 0 else ws // 2
                if delta is None:
                    h = LN(x, w[p + "norm1.w"], w[p + "norm1.b"])
                else:
                    x, h = LN(x, w[p + "norm1.w"], w[p + "norm1.b"], r1=delta, return_sum=True)
                qkv = F.linear(h, w[p + "attn.qkv.w"], w[p + "attn.qkv.b"])
                a = kernels.timed("window_attention_stage%d" % s, kernels.window_attention, qkv, w[p + "attn.qkv.b"],
                                  w[p + "rel"], B, Wh, Ww, C, nh, ws, shift)
                a = F.linear(a, w[p + "attn.proj.w"], w[p + "attn.proj.b"])
                x, h = LN(x, w[p + "norm2.w"], w[p + "norm2.b"], r1=a, return_sum=True)
                if kernels.linear_fused_supported(h, w[p + "mlp.fc1.w"], "gelu_erf"):
                    h = kernels.linear_fused(h, w[p + "mlp.fc1.w"], w[p + "mlp.fc1.b"], "gelu_erf")   # GELU in the epilogue
                else:
                    h = F.gelu(F.linear(h, w[p + "mlp.fc1.w"], w[p + "mlp.fc1.b"]))
                delta = F.linear(h, w[p + "mlp.fc2.w"], w[p + "mlp.fc2.b"])
            x, out = LN(x, w["norm%d.w" % s], w["norm%d.b" % s], r1=delta, return_sum=True)
            outs.append(out)
            sizes.append((Wh, Ww))
            if s < len(cfg.depths) - 1:  # PatchMerging swin_trans.py:269-296
                p = "layers.%d.downsample." % s
                # pad to even size + the four strided slices + cat + norm: one gather-LayerNorm kernel
                xm = kernels.patch_merge_layer_norm(x.contiguous(), Wh, Ww, w[p + "norm.w"], w[p + "norm.b"])
                Wh, Ww = (Wh + 1) // 2, (Ww + 1) // 2
                x = F.linear(xm, w[p + "red.w"])
                C = 2 * C
        return outs, sizes
