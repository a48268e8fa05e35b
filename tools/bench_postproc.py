"""Time the fused post-processing kernels (generic vs tensor-core) on the bench workload shape.
usage: python tools/bench_postproc.py [--dtype bf16]"""
import argparse
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from psalm_b200 import _lib, kernels  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--ncu", action="store_true", help="one tensor-core panoptic call between cudaProfilerStart/Stop")
    a = ap.parse_args()
    dt = {"bf16": torch.bfloat16, "fp16": torch.float16}[a.dtype]
    Q, ncls, H4, W4, H, W = 100, 133, 256, 256, 1024, 1024
    g = torch.Generator().manual_seed(0)
    logits = (torch.randn(Q, H4, W4, generator=g) * 4).to(dt).cuda()
    probs = F.softmax(torch.randn(Q, ncls + 1, generator=g) * 3, -1)[:, :-1]
    probsT = torch.zeros(144, 112, dtype=torch.float16)
    probsT[:ncls, :Q] = probs.t().half()
    probsT = probsT.cuda()
    keep = torch.rand(Q, generator=g) > 0.5
    wq = torch.where(keep, torch.rand(Q, generator=g), torch.zeros(Q)).cuda()
    negq = (keep.float() - 1).cuda()
    slots = torch.arange(100, dtype=torch.int32).cuda()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    if a.ncu:
        _lib.check(_lib.lib().psalm_set_postproc_impl(2), "set")
        kernels.postproc_fused(logits, H, W, probsT=probsT, wq=wq, negq=negq, slot_query=slots, ncls=ncls)
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
        kernels.postproc_fused(logits, H, W, probsT=probsT, wq=wq, negq=negq, slot_query=slots, ncls=ncls)
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
        return
    for impl, name in ((1, "generic"), (2, "tensor-core")):
        _lib.check(_lib.lib().psalm_set_postproc_impl(impl), "set")
        for combo, kw in (("panoptic(sem+inst+pan)", dict(probsT=probsT, wq=wq, negq=negq, slot_query=slots, ncls=ncls)),
                          ("instance", dict(slot_query=slots)), ("semantic", dict(probsT=probsT, ncls=ncls))):
            ts = []
            for i in range(a.iters + 3):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                kernels.postproc_fused(logits, H, W, **kw)
                e1.record()
                torch.cuda.synchronize()
                if i >= 3:
                    ts.append(e0.elapsed_time(e1) * 1e3)
            ts.sort()
            print(f"{name:12s} {combo:24s} median {ts[len(ts) // 2]:8.1f} us  min {ts[0]:8.1f} us (incl. partial-sum reduce)")
    _lib.check(_lib.lib().psalm_set_postproc_impl(0), "set")


if __name__ == "__main__":
    main()
