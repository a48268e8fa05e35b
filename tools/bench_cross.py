"""Time masked cross-attention (100 queries x Lk keys, 8 heads x 32) for the split-K variants.
usage: python tools/bench_cross.py [--ncu]   (--ncu: one call per configuration, for a profiler run)"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from psalm_b200 import _lib, kernels  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ncu", action="store_true")
    ap.add_argument("--iters", type=int, default=30)
    a = ap.parse_args()
    dt = torch.bfloat16
    B, Lq, nh, C = 1, 100, 8, 256
    g = torch.Generator().manual_seed(0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for Lk in (1024, 4096, 16384):
        q, k, v = (torch.randn(B, n, C, generator=g).to(dt).cuda() for n in (Lq, Lk, Lk))
        bits = torch.randint(-2**31, 2**31 - 1, (B, Lq, (Lk + 31) // 32), generator=g, dtype=torch.int64).to(torch.int32).cuda()
        row_open = torch.zeros(B, Lq, dtype=torch.uint8).cuda()
        for impl, name in ((3, "cluster"), (2, "workspace")):
            _lib.check(_lib.lib().psalm_set_attention_impl(impl), "set")
            for splits in (None, 4, 8, 16):
                sp = splits or kernels.pick_splits(B, nh, Lq, Lk)
                if sp > (Lk + 63) // 64:
                    continue
                ws = torch.empty(_lib.lib().psalm_cross_attention_workspace_bytes(B, nh, 32, Lq, sp) // 4 + 16, dtype=torch.float32, device="cuda")
                if a.ncu:
                    torch.cuda.cudart().cudaProfilerStart()
                    kernels.cross_attention(q, k, v, bits, row_open, nh, splits=sp, workspace=ws)
                    torch.cuda.cudart().cudaProfilerStop()
                    continue
                ts = []
                for i in range(a.iters + 3):
                    flush.zero_()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    torch.cuda._sleep(2_000_000)
                    e0.record()
                    kernels.cross_attention(q, k, v, bits, row_open, nh, splits=sp, workspace=ws)
                    e1.record()
                    torch.cuda.synchronize()
                    if i >= 3:
                        ts.append(e0.elapsed_time(e1) * 1e3)
                ts.sort()
                print(f"Lk={Lk:6d} {name:10s} splits={sp:3d}{'*' if splits is None else ' '} median {ts[len(ts) // 2]:7.1f} us  min {ts[0]:7.1f} us")
    _lib.check(_lib.lib().psalm_set_attention_impl(0), "set")


if __name__ == "__main__":
    main()
