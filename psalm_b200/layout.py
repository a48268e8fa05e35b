"""Checkpoint layout contract of the reference `PSALM` state dict (SURVEY.md Appendix B;
reference psalm/model/builder.py:55 loads exactly these keys with `from_pretrained`).

`checkpoint_layout()` enumerates every tensor name / shape / dtype the hot path owns, so that
  * a Hugging Face checkpoint directory of the reference loads unchanged (loader.py),
  * synthetic weights of the same architecture can be produced without the reference (synth.py),
  * tests can pin the layout against a manifest dumped from the reference constructors
    (tests/golden/state_dict_manifest_phi2layers.json).
"""
from collections import OrderedDict
from dataclasses import dataclass, field
from typing import Tuple


@dataclass(frozen=True)
class SwinConfig:  # build_swin_b, multimodal_encoder/swin_trans.py:660-678
    embed_dim: int = 128
    depths: Tuple[int, ...] = (2, 2, 18, 2)
    num_heads: Tuple[int, ...] = (4, 8, 16, 32)
    window: int = 12
    patch: int = 4
    mlp_ratio: int = 4


@dataclass(frozen=True)
class PhiConfig:  # transformers PhiConfig defaults == Phi-1.5 (llava_phi.py:34-35)
    hidden: int = 2048
    layers: int = 24
    heads: int = 32
    inter: int = 8192
    vocab: int = 51200
    eps: float = 1e-5
    rope_theta: float = 10000.0
    rotary_frac: float = 0.5

    @property
    def head_dim(self):
        return self.hidden // self.heads


@dataclass(frozen=True)
class MaskConfig:  # psalm/mask_config/maskformer2_R50_bs16_50ep.yaml:18-55 frozen to constants
    hidden: int = 256
    num_queries: int = 100
    nheads: int = 8
    dim_ff: int = 2048
    dec_layers: int = 9          # DEC_LAYERS 10 -> 9 (llava_phi.py:459)
    enc_layers: int = 6
    enc_ff: int = 1024           # hard-coded, llava_phi.py:516
    enc_levels: int = 3
    enc_points: int = 4
    mask_dim: int = 256
    size_divisibility: int = 32
    object_mask_threshold: float = 0.8   # llava_phi.py:331-332
    overlap_threshold: float = 0.8


@dataclass(frozen=True)
class PsalmConfig:
    swin: SwinConfig = field(default_factory=SwinConfig)
    phi: PhiConfig = field(default_factory=PhiConfig)
    mask: MaskConfig = field(default_factory=MaskConfig)
    projector_mid: int = 2048    # ResNetSwin BasicBlock planes, multimodal_projector/builder.py:335


F32, I64 = "float32", "int64"


def checkpoint_layout(cfg: PsalmConfig = PsalmConfig(), include_lm_head=True):
    """OrderedDict name -> (shape tuple, dtype str, kind).  kind drives synthetic init:
    'w' matrix/conv weight, 'b' bias, 'g' norm gain, 'emb' embedding, 'table', 'index', 'mean', 'var', 'count'."""
    L = OrderedDict()

    def lin(name, out_f, in_f, bias=True):
        L[name + ".weight"] = ((out_f, in_f), F32, "w")
        if bias:
            L[name + ".bias"] = ((out_f,), F32, "b")

    def norm(name, c):
        L[name + ".weight"] = ((c,), F32, "g")
        L[name + ".bias"] = ((c,), F32, "b")

    def conv(name, out_c, in_c, k, bias=True):
        L[name + ".weight"] = ((out_c, in_c, k, k), F32, "w")
        if bias:
            L[name + ".bias"] = ((out_c,), F32, "b")

    def bn(name, c):
        norm(name, c)
        L[name + ".running_mean"] = ((c,), F32, "mean")
        L[name + ".running_var"] = ((c,), F32, "var")
        L[name + ".num_batches_tracked"] = ((), I64, "count")

    ph, sw, mk = cfg.phi, cfg.swin, cfg.mask
    # --- Phi (transformers PhiModel key names) ---
    L["model.embed_tokens.weight"] = ((ph.vocab, ph.hidden), F32, "emb")
    for i in range(ph.layers):
        p = "model.layers.%d." % i
        for n in ("q_proj", "k_proj", "v_proj", "dense"):
            lin(p + "self_attn." + n, ph.hidden, ph.hidden)
        lin(p + "mlp.fc1", ph.inter, ph.hidden)
        lin(p + "mlp.fc2", ph.hidden, ph.inter)
        norm(p + "input_layernorm", ph.hidden)
    norm("model.final_layernorm", ph.hidden)
    # --- Swin (multimodal_encoder/swin_trans.py) ---
    v = "model.vision_tower."
    conv(v + "patch_embed.proj", sw.embed_dim, 3, sw.patch)
    norm(v + "patch_embed.norm", sw.embed_dim)
    ws = sw.window
    for s, depth in enumerate(sw.depths):
        C = sw.embed_dim * 2 ** s
        for b in range(depth):
            p = "%slayers.%d.blocks.%d." % (v, s, b)
            norm(p + "norm1", C)
            L[p + "attn.relative_position_bias_table"] = (((2 * ws - 1) ** 2, sw.num_heads[s]), F32, "table")
            L[p + "attn.relative_position_index"] = ((ws * ws, ws * ws), I64, "index")
            lin(p + "attn.qkv", 3 * C, C)
            lin(p + "attn.proj", C, C)
            norm(p + "norm2", C)
            lin(p + "mlp.fc1", sw.mlp_ratio * C, C)
            lin(p + "mlp.fc2", C, sw.mlp_ratio * C)
        if s < len(sw.depths) - 1:
            lin("%slayers.%d.downsample.reduction" % (v, s), 2 * C, 4 * C, bias=False)
            norm("%slayers.%d.downsample.norm" % (v, s), 4 * C)
    for s in range(len(sw.depths)):
        norm("%snorm%d" % (v, s), sw.embed_dim * 2 ** s)
    # --- projector (ResNetSwin, multimodal_projector/builder.py:326-375) ---
    pj = "model.mm_projector."
    cin = sw.embed_dim * 2 ** (len(sw.depths) - 1)
    mid = cfg.projector_mid
    conv(pj + "layer1.0.conv1", mid, cin, 3, bias=False)
    bn(pj + "layer1.0.bn1", mid)
    conv(pj + "layer1.0.conv2", mid, mid, 3, bias=False)
    bn(pj + "layer1.0.bn2", mid)
    conv(pj + "layer1.0.downsample.0", mid, cin, 1, bias=False)
    bn(pj + "layer1.0.downsample.1", mid)
    lin(pj + "fc", ph.hidden, mid)
    if include_lm_head:
        L["lm_head.weight"] = ((ph.vocab, ph.hidden), F32, "w")
    # --- PSALM glue (llava_phi.py:163,174-185) ---
    L["seg_query"] = ((mk.num_queries, ph.hidden), F32, "emb")
    for n in ("region_projector", "seg_query_projector", "SEG_token_projector", "class_name_projector"):
        lin(n, mk.hidden, ph.hidden)
    # --- pixel decoder (pixel_decoder/msdeformattn.py) ---
    pd = "pixel_decoder."
    H = mk.hidden
    chans = [sw.embed_dim * 2 ** s for s in range(len(sw.depths))]
    for i, c in enumerate(chans[:0:-1]):  # res5, res4, res3
        conv("%sinput_proj.%d.0" % (pd, i), H, c, 1)
        norm("%sinput_proj.%d.1" % (pd, i), H)
    L[pd + "transformer.level_embed"] = ((mk.enc_levels, H), F32, "emb")
    for i in range(mk.enc_layers):
        p = "%stransformer.encoder.layers.%d." % (pd, i)
        lin(p + "self_attn.sampling_offsets", mk.nheads * mk.enc_levels * mk.enc_points * 2, H)
        lin(p + "self_attn.attention_weights", mk.nheads * mk.enc_levels * mk.enc_points, H)
        lin(p + "self_attn.value_proj", H, H)
        lin(p + "self_attn.output_proj", H, H)
        norm(p + "norm1", H)
        lin(p + "linear1", mk.enc_ff, H)
        lin(p + "linear2", H, mk.enc_ff)
        norm(p + "norm2", H)
    conv(pd + "mask_features", mk.mask_dim, H, 1)
    conv(pd + "adapter_1.0", H, chans[0], 1)
    norm(pd + "adapter_1.1", H)
    conv(pd + "layer_1.0", H, H, 3)
    norm(pd + "layer_1.1", H)
    # --- masked-attention decoder (transformer_decoder/mask2former_transformer_decoder.py:394-486) ---
    pr = "predictor."
    for i in range(mk.dec_layers):
        for kind, attn in (("self", "self_attn"), ("cross", "multihead_attn")):
            p = "%stransformer_%s_attention_layers.%d." % (pr, kind, i)
            L[p + attn + ".in_proj_weight"] = ((3 * H, H), F32, "w")
            L[p + attn + ".in_proj_bias"] = ((3 * H,), F32, "b")
            lin(p + attn + ".out_proj", H, H)
            norm(p + "norm", H)
        p = "%stransformer_ffn_layers.%d." % (pr, i)
        lin(p + "linear1", mk.dim_ff, H)
        lin(p + "linear2", H, mk.dim_ff)
        norm(p + "norm", H)
    norm(pr + "decoder_norm", H)
    L[pr + "query_feat.weight"] = ((mk.num_queries, H), F32, "emb")
    L[pr + "query_embed.weight"] = ((mk.num_queries, H), F32, "emb")
    L[pr + "SEG_query_embed.weight"] = ((mk.num_queries + 1, H), F32, "emb")
    L[pr + "level_embed.weight"] = ((3, H), F32, "emb")
    for i in range(3):
        lin("%smask_embed.layers.%d" % (pr, i), H, H)
    for n in ("SEG_proj", "CLASS_proj", "REGION_proj"):
        for i in range(2):
            lin("%s%s.layers.%d" % (pr, n, i), H, H)
    return L
