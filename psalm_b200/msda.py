"""Host-side operator for multi-scale deformable attention (forward only, inference tier).

Mirrors the reference's pybind module `MultiScaleDeformableAttention`
(mask_decoder/Mask2Former_Simplify/modeling/pixel_decoder/ops/src/vision.cpp:18-21):

    ms_deform_attn_forward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step)
    ms_deform_attn_backward(...)   -> NotImplementedError (training is out of scope)

Argument meaning, dtype rules and error behaviour follow ops/src/cuda/ms_deform_attn_cuda.cu:25-85:
all tensors contiguous and CUDA (else RuntimeError), `batch % min(batch, im2col_step) == 0`, output
`[B, Lq, M*D]` in `value.dtype`.  CPU tensors raise, as in ops/src/ms_deform_attn.h:43.
"""
import torch

from . import _lib


def _shapes_on_host(spatial_shapes, level_start_index):
    """Accept python lists / CPU tensors (host path: enables the 2-D patch schedule) or CUDA int64
    tensors exactly like the reference (device path, no host sync)."""
    if isinstance(spatial_shapes, torch.Tensor) and spatial_shapes.is_cuda:
        return None
    if isinstance(spatial_shapes, torch.Tensor):
        spatial_shapes = spatial_shapes.tolist()
    if isinstance(level_start_index, torch.Tensor):
        level_start_index = level_start_index.tolist()
    flat = [int(v) for hw in spatial_shapes for v in hw]
    return _lib.i64_array(flat), _lib.i64_array(level_start_index), len(level_start_index)


def ms_deform_attn_forward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight,
                           im2col_step=128, value_layout=0):
    """value [B,S,M,D] (value_layout=0, reference) or [B,M,S,D] (value_layout=1, head-major)."""
    for name, t in (("value", value), ("sampling_loc", sampling_loc), ("attn_weight", attn_weight)):
        if not t.is_cuda:
            raise RuntimeError("Not implemented on the CPU (%s is on %s)" % (name, t.device))
        if not t.is_contiguous():
            raise RuntimeError("%s tensor has to be contiguous" % name)
    if value_layout == 0:
        B, S, M, D = value.shape
    else:
        B, M, S, D = value.shape
    _, Lq, M2, L, P, two = sampling_loc.shape
    if M2 != M or two != 2 or tuple(attn_weight.shape) != (B, Lq, M, L, P):
        raise RuntimeError("ms_deform_attn_forward: inconsistent shapes value=%s loc=%s w=%s"
                           % (tuple(value.shape), tuple(sampling_loc.shape), tuple(attn_weight.shape)))
    step = min(B, int(im2col_step))
    if B % step != 0:
        raise RuntimeError("batch(%d) must divide im2col_step(%d)" % (B, step))
    if sampling_loc.dtype != attn_weight.dtype:
        raise RuntimeError("sampling_loc and attn_weight must share a dtype")
    out = torch.empty((B, Lq, M * D), dtype=value.dtype, device=value.device)
    host = _shapes_on_host(spatial_shapes, level_start_index)
    with torch.cuda.device(value.device):
        if host is None:
            if not (spatial_shapes.is_contiguous() and level_start_index.is_contiguous()):
                raise RuntimeError("spatial_shapes / level_start_index tensor has to be contiguous")
            if spatial_shapes.dtype != torch.int64 or level_start_index.dtype != torch.int64:
                raise RuntimeError("spatial_shapes / level_start_index must be int64")
            sh, st, on_host = _lib.ptr(spatial_shapes), _lib.ptr(level_start_index), 0
            if spatial_shapes.shape[0] != L:
                raise RuntimeError("spatial_shapes has %d levels, sampling_loc has %d" % (spatial_shapes.shape[0], L))
        else:
            sh, st, nl = host
            on_host = 1
            if nl != L:
                raise RuntimeError("spatial_shapes has %d levels, sampling_loc has %d" % (nl, L))
        rc = _lib.lib().psalm_msda_forward(
            _lib.ptr(value), sh, st, _lib.ptr(sampling_loc), _lib.ptr(attn_weight), _lib.ptr(out),
            B, S, M, D, L, Lq, P, _lib.dtype_code(value.dtype), _lib.dtype_code(sampling_loc.dtype),
            value_layout, on_host, _lib.stream_ptr(value.device))
    _lib.check(rc, "psalm_msda_forward")
    return out


def ms_deform_attn_backward(*args, **kwargs):
    raise NotImplementedError("psalm_b200 is an inference build: ms_deform_attn_backward is not provided "
                              "(reference ops/src/ms_deform_attn.h:46-65 is training-only)")


def msda_encoder_fused(value_hm, ow, spatial_shapes, level_start_index, n_points=4):
    """Fused softmax + reference points + sampling for the pixel-decoder encoder.

    value_hm [B,M,S,D] head-major; ow [B,S,M*L*P*3] raw (offsets | logits) Linear output."""
    _lib.require_cuda(value_hm, ow)
    if not (value_hm.is_contiguous() and ow.is_contiguous()):
        raise RuntimeError("msda_encoder_fused: tensors must be contiguous")
    B, M, S, D = value_hm.shape
    L = len(level_start_index)
    if tuple(ow.shape) != (B, S, M * L * n_points * 3):
        raise RuntimeError("msda_encoder_fused: ow shape %s != %s" % (tuple(ow.shape), (B, S, M * L * n_points * 3)))
    out = torch.empty((B, S, M * D), dtype=value_hm.dtype, device=value_hm.device)
    flat = [int(v) for hw in spatial_shapes for v in hw]
    with torch.cuda.device(value_hm.device):
        rc = _lib.lib().psalm_msda_encoder_fused(
            _lib.ptr(value_hm), _lib.ptr(ow), _lib.ptr(out), _lib.i64_array(flat),
            _lib.i64_array(level_start_index), B, S, M, D, L, n_points,
            _lib.dtype_code(value_hm.dtype), _lib.dtype_code(ow.dtype), _lib.stream_ptr(value_hm.device))
    _lib.check(rc, "psalm_msda_encoder_fused")
    return out
