#!/usr/bin/env python
"""Build the REFERENCE's own MSDeformAttn CUDA op for sm_100 — the GPU kernel-to-beat (SURVEY.md section 8d,
BASELINE.md section 3) — into the git-ignored baseline/_ref/ (the built .so travels to the GPU box; no reference
source is copied into this repository).

    python baseline/build_ref_msda.py            # needs /root/reference (build container only)

Recipe: copy ops/src to a scratch directory under /tmp (the reference tree is read-only), apply the two-line
`value.type()` -> `value.scalar_type()` fix inside the AT_DISPATCH macros (ms_deform_attn_cuda.cu:69,139; torch >= 2
removed the deprecated overload), compile with torch.utils.cpp_extension for compute capability 10.0.
tools/bench_msda.py loads baseline/_ref/MultiScaleDeformableAttention*.so when present and times
`ms_deform_attn_forward` (ms_deformable_im2col_gpu_kernel, ms_deform_im2col_cuda.cuh:243-304) beside our kernels.
"""
import glob
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_OPS = "/root/reference/psalm/model/mask_decoder/Mask2Former_Simplify/modeling/pixel_decoder/ops/src"
OUT = os.path.join(ROOT, "baseline", "_ref")
SCRATCH = "/tmp/psalm_ref_msda_src"


def main():
    if not os.path.isdir(REF_OPS):
        print("reference sources not present (%s): nothing to build" % REF_OPS)
        return 1
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
    os.environ.setdefault("MAX_JOBS", "8")
    shutil.rmtree(SCRATCH, ignore_errors=True)
    shutil.copytree(REF_OPS, SCRATCH)
    cu = os.path.join(SCRATCH, "cuda", "ms_deform_attn_cuda.cu")
    src = open(cu).read()
    n = src.count("HALF(value.type(),")
    src = src.replace("HALF(value.type(),", "HALF(value.scalar_type(),")   # lines 69 and 139 only
    open(cu, "w").write(src)
    print("patched %d dispatch sites" % n)
    from torch.utils.cpp_extension import load
    os.makedirs(OUT, exist_ok=True)
    sources = [os.path.join(SCRATCH, "vision.cpp")] + glob.glob(os.path.join(SCRATCH, "cpu", "*.cpp")) + \
        glob.glob(os.path.join(SCRATCH, "cuda", "*.cu"))
    load(name="MultiScaleDeformableAttention", sources=sources, extra_include_paths=[SCRATCH],
         extra_cflags=["-DWITH_CUDA"], extra_cuda_cflags=["-DWITH_CUDA", "-DCUDA_HAS_FP16=1", "-D__CUDA_NO_HALF_OPERATORS__",
                                                          "-D__CUDA_NO_HALF_CONVERSIONS__", "-D__CUDA_NO_HALF2_OPERATORS__"],
         build_directory=OUT, is_python_module=False, verbose=True)
    print("built:", glob.glob(os.path.join(OUT, "*.so")))
    return 0


if __name__ == "__main__":
    sys.exit(main())
