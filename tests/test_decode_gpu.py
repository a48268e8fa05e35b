"""Paged-KV decode kernels (csrc/decode.cu) against the torch restatement, and PSALM.generate on the GPU against the
oracle's full-recompute greedy decoding."""
import pytest
import torch

import emu
from psalm_b200 import kernels, synth
from psalm_b200.layout import PhiConfig, PsalmConfig

pytestmark = pytest.mark.gpu
DT = {"f32": torch.float32, "f16": torch.float16, "bf16": torch.bfloat16}
TOL = {"f32": 2e-5, "f16": 2e-3, "bf16": 1.5e-2}


@pytest.mark.parametrize("dt", ["f32", "f16", "bf16"])
@pytest.mark.parametrize("B,nh,hd,ps,lens", [(1, 32, 64, 16, [921]), (3, 4, 64, 16, [1, 17, 300]), (2, 8, 32, 8, [64, 5]),
                                             (2, 2, 64, 4, [2048, 1023])])
def test_cache_write_and_paged_decode(dt, B, nh, hd, ps, lens):
    torch.manual_seed(B * 7 + nh)
    max_len = max(lens) + 3
    max_pages = -(-max_len // ps)
    perm = torch.randperm(B * max_pages)                       # physical pages in random order
    bt = perm.view(B, max_pages).to(torch.int32)
    kc = torch.zeros(B * max_pages, nh, ps, hd, dtype=DT[dt])
    vc = torch.zeros_like(kc)
    kc_g, vc_g = kc.cuda(), vc.cuda()
    # prefill of different lengths written token block by token block, then one decode step
    T0 = max(lens)
    qkv_pre = torch.randn(B, T0, 3, nh, hd).to(DT[dt])
    start = torch.zeros(B, dtype=torch.int32)
    emu.kv_cache_write(qkv_pre, kc, vc, bt, start)
    kernels.kv_cache_write(qkv_pre.cuda(), kc_g, vc_g, bt.cuda(), start.cuda())
    assert torch.equal(kc_g.cpu(), kc) and torch.equal(vc_g.cpu(), vc)        # pure data movement: bit exact
    seq = torch.tensor(lens, dtype=torch.int32)
    qkv_new = torch.randn(B, 1, 3, nh, hd).to(DT[dt])
    emu.kv_cache_write(qkv_new, kc, vc, bt, seq)
    kernels.kv_cache_write(qkv_new.cuda(), kc_g, vc_g, bt.cuda(), seq.cuda())
    ref = emu.paged_decode_attention(qkv_new.float(), kc.float(), vc.float(), bt, seq + 1)
    out = kernels.paged_decode_attention(qkv_new.cuda(), kc_g, vc_g, bt.cuda(), (seq + 1).cuda())
    torch.cuda.synchronize()
    err = (out.float().cpu() - ref).abs().max() / ref.abs().max()
    assert err < TOL[dt], err


def test_generate_on_gpu_matches_full_recompute():
    """fp32 storage: the greedy tokens of prefill + paged decode equal the oracle's re-run of the whole sequence per token."""
    from test_generate_cpu import SMALL, oracle_generate
    from psalm_b200.psalm import PSALM
    sd = synth.synth_state_dict(SMALL, seed=3, include_lm_head=True)
    m = PSALM(sd, SMALL, torch.float32, "cuda", "panoptic")
    g = torch.Generator().manual_seed(0)
    ids = torch.randint(5, 1000, (2, 300), generator=g)       # T >= 256: the prefill takes the large-T attention path
    got = m.generate(ids, max_new_tokens=8)
    assert torch.equal(got.cpu(), oracle_generate(sd, ids, None, 8))
    ids2 = torch.cat([ids[:1, :7], torch.tensor([[-200]]), ids[:1, 7:40]], 1)
    img = torch.randn(1, 3, 96, 128, generator=g)
    got2 = m.generate(ids2, images=img, max_new_tokens=6)
    assert torch.equal(got2.cpu(), oracle_generate(sd, ids2, img, 6))
    # 16-bit storage runs the same path (tokens may legitimately differ from fp32 at near ties): shape / range only
    mb = PSALM(sd, SMALL, torch.bfloat16, "cuda", "panoptic")
    gb = mb.generate(ids, max_new_tokens=4)
    assert tuple(gb.shape) == (2, 4) and int(gb.min()) >= 0 and int(gb.max()) < 1000
