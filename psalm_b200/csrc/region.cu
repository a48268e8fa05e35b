// Region prompts: point-sampled, mean-pooled region features.
//
// Replaces `region_pooling.forward` (visual_prompt_module/context_cluster.py:333-400): per region, the projector's
// feature map [h, w, C] is repeated once per region, 256 points are bilinearly sampled with
// F.grid_sample(align_corners=True) (context_cluster.py:43-68, :355-371) and averaged (AdaptiveAvgPool1d, :392).
// Here one CTA handles one region x one 16-byte channel vector per thread: the points are read once (broadcast), the
// four corners are 16-byte loads from the token-major map, sums stay in fp32 registers, no [K, C, h, w] copy exists.
// The sample POINTS are an input (the reference draws them on the host with torch.randperm / torch.randint,
// context_cluster.py:31-40; the host side here does the same, psalm_b200/region.py).
#include "common.cuh"

namespace psalm {

template <typename T>
__global__ void __launch_bounds__(256) region_pool_kernel(const T* __restrict__ tokens, const float* __restrict__ points,
                                                          const int* __restrict__ region_image, T* __restrict__ out, int h,
                                                          int w, int C, int P) {
  constexpr int CH = Vec16<T>::CH;
  const int r = blockIdx.x;
  const int c = (blockIdx.y * blockDim.x + threadIdx.x) * CH;
  if (c >= C) return;
  const T* map = tokens + (size_t)region_image[r] * h * w * C + c;
  const float* pt = points + (size_t)r * P * 2;
  float acc[CH];
#pragma unroll
  for (int e = 0; e < CH; ++e) acc[e] = 0.f;
  for (int p = 0; p < P; ++p) {
    // grid = 2 * (x, y) - 1, then grid_sampler_unnormalize(align_corners=True): ((g + 1) / 2) * (size - 1)
    const float gy = 2.0f * __ldg(pt + 2 * p) - 1.0f, gx = 2.0f * __ldg(pt + 2 * p + 1) - 1.0f;
    const float iy = ((gy + 1.f) / 2) * (float)(h - 1), ix = ((gx + 1.f) / 2) * (float)(w - 1);
    const float fy = floorf(iy), fx = floorf(ix);
    const int y0 = (int)fy, x0 = (int)fx;
    const float ly = iy - fy, lx = ix - fx;
    const float wgt[4] = {(1.f - ly) * (1.f - lx), (1.f - ly) * lx, ly * (1.f - lx), ly * lx};   // nw, ne, sw, se
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int yy = y0 + (k >> 1), xx = x0 + (k & 1);
      if (yy < 0 || yy >= h || xx < 0 || xx >= w) continue;   // padding_mode = zeros
      float f[CH];
      load16_as_f32<T>(map + ((size_t)yy * w + xx) * C, f);
#pragma unroll
      for (int e = 0; e < CH; ++e) acc[e] = fmaf(f[e], wgt[k], acc[e]);
    }
  }
  const float inv = 1.f / (float)P;
#pragma unroll
  for (int e = 0; e < CH; ++e) acc[e] *= inv;
  store16_from_f32<T>(out + (size_t)r * C + c, acc);
}

template <typename T>
static int launch_region_pool(const void* tokens, const float* points, const int* region_image, void* out, int h, int w, int C,
                              int R, int P, cudaStream_t st) {
  constexpr int CH = Vec16<T>::CH;
  const int vecs = C / CH;
  dim3 grid(R, (vecs + 255) / 256);
  region_pool_kernel<T><<<grid, 256, 0, st>>>((const T*)tokens, points, region_image, (T*)out, h, w, C, P);
  return check_launch("region_pool_kernel");
}

}  // namespace psalm

extern "C" int psalm_region_pool(const void* tokens, const float* points, const int* region_image, void* out, int B, int h,
                                 int w, int C, int R, int P, int dtype, void* stream) {
  using namespace psalm;
  PSALM_REQUIRE(tokens && points && region_image && out, "region_pool: null pointer");
  PSALM_REQUIRE(B > 0 && h > 0 && w > 0 && R > 0 && P > 0, "region_pool: bad shape");
  PSALM_REQUIRE(C > 0 && C % (dtype == PSALM_F32 ? 4 : 8) == 0, "region_pool: C=%d must be a multiple of the 16-byte vector", C);
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case PSALM_F32: return launch_region_pool<float>(tokens, points, region_image, out, h, w, C, R, P, st);
    case PSALM_F16: return launch_region_pool<__half>(tokens, points, region_image, out, h, w, C, R, P, st);
    case PSALM_BF16: return launch_region_pool<__nv_bfloat16>(tokens, points, region_image, out, h, w, C, R, P, st);
  }
  set_error("region_pool: unknown dtype %d", dtype);
  return PSALM_E_ARG;
}
