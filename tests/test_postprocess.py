"""Task heads: the fused post-processing (one kernel on the low-resolution logits) equals the step-by-step
path that mirrors the reference (llava_phi.py:308-447) on the up-sampled map.
CPU: python glue with the kernel emulated (tests/emu.py); GPU: the real kernel."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import emu
from psalm_b200 import postprocess as PP


def _inputs(seed, Q=100, H4=24, W4=40, ncls=20, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    logits = (torch.randn(Q, H4, W4, generator=g) * 4).to(dtype)
    cls = torch.randn(Q, ncls + 1, generator=g) * 4
    cls[:30, : ncls] += 6 * F.one_hot(torch.randint(0, ncls, (30,), generator=g), ncls)   # confident queries
    seg = torch.randn(Q, 1, generator=g)
    thing = [True] * (ncls // 2) + [False] * (ncls - ncls // 2)
    return logits, cls, seg, thing


def _reference_path(logits, cls, seg, thing, H, W, task):
    mp = F.interpolate(logits.float()[None], size=(H, W), mode="bilinear", align_corners=False)[0]
    r = {}
    if task in ("semantic", "panoptic"):
        r["sem_seg"] = PP.semantic_inference(cls, mp)
    if task in ("instance", "panoptic"):
        r["instances"] = PP.instance_inference(cls, mp, 100, thing, task == "panoptic")
    if task == "panoptic":
        r["panoptic_seg"] = PP.panoptic_inference(cls, mp, thing)
    if task == "referring":
        r["instances"] = PP.seg_instance_inference(seg, mp, 100)
    return r


def _compare(ref, got, task, sem_tol):
    if "sem_seg" in ref:
        err = (got["sem_seg"].float().cpu() - ref["sem_seg"].cpu()).abs().max() / ref["sem_seg"].abs().max()
        assert err < sem_tol, err
    if "panoptic_seg" in ref:
        pr, ir = ref["panoptic_seg"]
        pg, ig = got["panoptic_seg"]
        assert ir == ig
        assert (pr.cpu() != pg.cpu()).float().mean() < 1e-3
    if "instances" in ref:
        a, b = ref["instances"], got["instances"]
        assert len(a.scores) == len(b.scores)
        key = lambda inst: sorted(zip(inst.query_index.cpu().tolist(),  # noqa: E731
                                      (inst.pred_classes.cpu().tolist() if inst.has("pred_classes") else [0] * len(inst.scores)),
                                      range(len(inst.scores))))
        ka, kb = key(a), key(b)
        assert [k[:2] for k in ka] == [k[:2] for k in kb]
        ia, ib = [k[2] for k in ka], [k[2] for k in kb]
        assert torch.allclose(a.scores.cpu()[ia], b.scores.cpu()[ib], rtol=2e-3, atol=1e-5)
        assert (a.pred_masks.cpu()[ia] != b.pred_masks.cpu()[ib]).float().mean() < 1e-4


@pytest.mark.parametrize("task", ["panoptic", "instance", "semantic", "referring"])
def test_fused_glue_cpu(monkeypatch, task):
    from psalm_b200 import kernels
    monkeypatch.setattr(kernels, "postproc_fused", emu.postproc_fused)
    logits, cls, seg, thing = _inputs(3)
    H, W = 96, 160
    ref = _reference_path(logits, cls, seg, thing, H, W, task)
    got = PP.fused_postprocess(kernels, logits, H, W, None if task == "referring" else cls, seg if task == "referring" else None,
                               thing, task in ("semantic", "panoptic"), task in ("instance", "panoptic"),
                               task == "panoptic", task == "referring")
    _compare(ref, got, task, 2e-3)


@pytest.mark.gpu
@pytest.mark.parametrize("task", ["panoptic", "instance", "semantic", "referring"])
@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16, torch.float32])
@pytest.mark.parametrize("shape", [(24, 40, 96, 160), (50, 66, 200, 264), (13, 21, 61, 85)])
def test_fused_kernel_gpu(task, dt, shape):
    from psalm_b200 import kernels
    H4, W4, H, W = shape
    logits, cls, seg, thing = _inputs(5, H4=H4, W4=W4, dtype=dt)
    ref = _reference_path(logits.float(), cls, seg, thing, H, W, task)
    got = PP.fused_postprocess(kernels, logits.cuda(), H, W, None if task == "referring" else cls.cuda(),
                               seg.cuda() if task == "referring" else None, thing, task in ("semantic", "panoptic"),
                               task in ("instance", "panoptic"), task == "panoptic", task == "referring")
    _compare(ref, got, task, 3e-3)
