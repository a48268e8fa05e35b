/*
 * TEST INFRASTRUCTURE — CPU restatement (plain C) of the reference MSDeformAttn forward.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this.
 *
 * Follows, loop for loop, the reference CUDA kernel
 *   ms_deformable_im2col_gpu_kernel   ops/src/cuda/ms_deform_im2col_cuda.cuh:243-304
 *   ms_deform_attn_im2col_bilinear    ops/src/cuda/ms_deform_im2col_cuda.cuh:39-89
 * (paths relative to /root/reference/psalm/model/mask_decoder/Mask2Former_Simplify/modeling/
 * pixel_decoder/).  Pinned against the reference's own CPU path ms_deform_attn_core_pytorch
 * (ops/functions/ms_deform_attn_func.py:52-78) on the vectors of ops/test.py:24-63 —
 * see tests/golden/msda_*.npz and oracle/gen_golden.py.
 *
 * Layouts: value [B,S,M,D], shapes [L,2] (H,W), starts [L], loc [B,Lq,M,L,P,2] (x,y in [0,1]),
 * w [B,Lq,M,L,P], out [B,Lq,M*D].
 */
#include <math.h>
#include <stdint.h>

#define DEFINE_MSDA(NAME, T)                                                                      \
  static T NAME##_bilinear(const T* data, int H, int W, int M, int D, T h, T w, int m, int c) {    \
    /* .cuh:44-88 */                                                                              \
    const int h_low = (int)floor((double)h), w_low = (int)floor((double)w);                       \
    const int h_high = h_low + 1, w_high = w_low + 1;                                              \
    const T lh = h - h_low, lw = w - w_low, hh = 1 - lh, hw = 1 - lw;                              \
    const long w_stride = (long)M * D, h_stride = (long)W * w_stride;                             \
    const long base = (long)m * D + c;                                                            \
    T v1 = 0, v2 = 0, v3 = 0, v4 = 0;                                                              \
    if (h_low >= 0 && w_low >= 0) v1 = data[h_low * h_stride + w_low * w_stride + base];          \
    if (h_low >= 0 && w_high <= W - 1) v2 = data[h_low * h_stride + w_high * w_stride + base];    \
    if (h_high <= H - 1 && w_low >= 0) v3 = data[h_high * h_stride + w_low * w_stride + base];    \
    if (h_high <= H - 1 && w_high <= W - 1) v4 = data[h_high * h_stride + w_high * w_stride + base]; \
    const T w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;                                \
    return w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4;                                                  \
  }                                                                                               \
  void NAME(const T* value, const int64_t* shapes, const int64_t* starts, const T* loc,           \
            const T* attn, T* out, int B, int S, int M, int D, int L, int Lq, int P) {            \
    for (int b = 0; b < B; ++b)                                                                    \
      for (int q = 0; q < Lq; ++q)                                                                 \
        for (int m = 0; m < M; ++m) {                                                              \
          const long si = ((long)b * Lq + q) * M + m; /* .cuh:263 sampling_index */               \
          for (int c = 0; c < D; ++c) {                                                            \
            long wptr = si * L * P, lptr = wptr << 1; /* .cuh:270-271 */                          \
            T col = 0;                                                                             \
            for (int l = 0; l < L; ++l) {                                                          \
              const int H = (int)shapes[2 * l], W = (int)shapes[2 * l + 1];                       \
              const T* vptr = value + ((long)b * S + starts[l]) * M * D; /* .cuh:273,282 */       \
              for (int p = 0; p < P; ++p) {                                                        \
                const T loc_w = loc[lptr], loc_h = loc[lptr + 1], weight = attn[wptr];            \
                const T h_im = loc_h * H - (T)0.5, w_im = loc_w * W - (T)0.5; /* .cuh:290-291 */  \
                if (h_im > -1 && w_im > -1 && h_im < H && w_im < W)           /* .cuh:293 */      \
                  col += NAME##_bilinear(vptr, H, W, M, D, h_im, w_im, m, c) * weight;            \
                wptr += 1;                                                                         \
                lptr += 2;                                                                         \
              }                                                                                    \
            }                                                                                      \
            out[si * D + c] = col;                                                                 \
          }                                                                                        \
        }                                                                                          \
  }

DEFINE_MSDA(msda_ref_f32, float)
DEFINE_MSDA(msda_ref_f64, double)
