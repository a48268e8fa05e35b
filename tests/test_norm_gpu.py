"""Fused residual-add + LayerNorm and token-major GroupNorm kernels vs torch restatements."""
import pytest
import torch

import emu
from psalm_b200 import kernels

pytestmark = pytest.mark.gpu
DT = {"f32": torch.float32, "f16": torch.float16, "bf16": torch.bfloat16}
TOL = {"f32": 1e-5, "f16": 2e-3, "bf16": 1.6e-2}


def _close(out, ref, dt):
    err = (out.float().cpu() - ref.float()).abs().max() / (ref.float().abs().max() + 1e-30)
    assert err < TOL[dt], "rel-to-max error %.3e" % err


@pytest.mark.parametrize("dt", ["f32", "f16", "bf16"])
@pytest.mark.parametrize("rows,C", [(1000, 128), (333, 256), (77, 512), (50, 1024), (921, 2048)])
@pytest.mark.parametrize("nres", [0, 1, 2])
def test_add_layernorm(dt, rows, C, nres):
    torch.manual_seed(rows + C + nres)
    x = (torch.randn(2, rows, C) * 2 + 0.5).to(DT[dt])
    rs = [(torch.randn(2, rows, C)).to(DT[dt]) for _ in range(nres)]
    w = (1 + 0.1 * torch.randn(C)).to(DT[dt])
    b = (0.1 * torch.randn(C)).to(DT[dt])
    r1 = rs[0] if nres > 0 else None
    r2 = rs[1] if nres > 1 else None
    ref_s, ref_y = emu.add_layer_norm(x, w, b, 1e-5, r1, r2, return_sum=True)
    s, y = kernels.add_layer_norm(x.cuda(), w.cuda(), b.cuda(), 1e-5, r1.cuda() if r1 is not None else None,
                                  r2.cuda() if r2 is not None else None, return_sum=True)
    _close(s, ref_s, dt)
    _close(y, ref_y, dt)
    y2 = kernels.add_layer_norm(x.cuda(), w.cuda(), b.cuda(), 1e-5, r1.cuda() if r1 is not None else None,
                                r2.cuda() if r2 is not None else None)
    _close(y2, emu.add_layer_norm(x, w, b, 1e-5, r1, r2), dt)


@pytest.mark.parametrize("dt", ["f32", "f16", "bf16"])
@pytest.mark.parametrize("N,C,relu", [(1024, 256, False), (5000, 256, True), (333, 128, True)])
def test_group_norm_tokens(dt, N, C, relu):
    torch.manual_seed(N + C)
    x = (torch.randn(2, N, C) * 3 + 1).to(DT[dt])
    w = (1 + 0.1 * torch.randn(C)).to(DT[dt])
    b = (0.1 * torch.randn(C)).to(DT[dt])
    ref = emu.group_norm_tokens(x, w, b, 32, 1e-5, relu)
    out = kernels.group_norm_tokens(x.cuda(), w.cuda(), b.cuda(), 32, 1e-5, relu)
    _close(out, ref, dt)
    pb = (torch.randn(C) * 2).to(DT[dt])          # bias of the producing conv, folded into the kernel
    ref = emu.group_norm_tokens(x, w, b, 32, 1e-5, relu, pre_bias=pb)
    out = kernels.group_norm_tokens(x.cuda(), w.cuda(), b.cuda(), 32, 1e-5, relu, pre_bias=pb.cuda())
    _close(out, ref, dt)


def test_unsupported_width_raises():
    from psalm_b200._lib import PsalmKernelError
    x = torch.zeros(4, 96, device="cuda")
    with pytest.raises(PsalmKernelError):
        kernels.add_layer_norm(x, torch.ones(96, device="cuda"), torch.zeros(96, device="cuda"))


@pytest.mark.parametrize("dt", ["f32", "f16", "bf16"])
@pytest.mark.parametrize("H,W,C", [(16, 24, 128), (15, 9, 256), (7, 8, 512), (1, 5, 128)])
def test_patch_merge_layernorm(dt, H, W, C):
    """Swin PatchMerging gather + LayerNorm in one kernel vs pad + strided slices + cat + LayerNorm
    (swin_trans.py:280-293), including odd H / W (zero padding takes part in the statistics)."""
    torch.manual_seed(H * 100 + W + C)
    x = (torch.randn(3, H * W, C) * 1.5 + 0.3).to(DT[dt])
    w = (1 + 0.1 * torch.randn(4 * C)).to(DT[dt])
    b = (0.1 * torch.randn(4 * C)).to(DT[dt])
    ref = emu.patch_merge_layer_norm(x, H, W, w, b)
    y = kernels.patch_merge_layer_norm(x.cuda(), H, W, w.cuda(), b.cuda())
    assert y.shape == ref.shape
    _close(y, ref, dt)


@pytest.mark.parametrize("dt", ["f16", "bf16", "f32"])
def test_bilinear_tokens_vector_kernel_equals_scalar_kernel(dt):
    """The 16-byte-vector resampling kernel evaluates the scalar kernel's expression (reached through an fp32-output
    call): equal up to the 16-bit rounding of the output."""
    torch.manual_seed(5)
    x = torch.randn(2, 24 * 20, 64).to(DT[dt]).cuda()
    for Ho, Wo in ((12, 10), (6, 5), (48, 40), (17, 13)):
        vec = kernels.bilinear_tokens(x, 24, 20, Ho, Wo)
        if dt == "f32":
            ref = emu.bilinear_tokens(x.cpu(), 24, 20, Ho, Wo)
            assert (vec.cpu() - ref).abs().max() < 1e-5
        else:
            scalar = kernels.bilinear_tokens(x, 24, 20, Ho, Wo, out_dtype=torch.float32)     # fp32 result of the scalar kernel
            ulp = 2.0 ** (-10 if dt == "f16" else -7)                                         # FMA contraction may differ: 1 ulp
            assert bool(((vec.float() - scalar).abs() <= ulp * scalar.abs() + 1e-6).all())
        acc0 = torch.randn(2, Ho * Wo, 64).to(DT[dt]).cuda()
        acc = kernels.bilinear_tokens(x, 24, 20, Ho, Wo, out=acc0.clone(), accumulate=True)
        ref = emu.bilinear_tokens(x.cpu(), 24, 20, Ho, Wo, out=acc0.cpu().clone(), accumulate=True)
        assert (acc.float().cpu() - ref.float()).abs().max() <= (2e-5 if dt == "f32" else 4e-2)
