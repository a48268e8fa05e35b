// Fused post-processing of eval_seg (reference language_model/llava_phi.py:1399-1406, 325-447).
//
// The reference up-samples the 100 mask-logit maps to [100, H, W] fp32 (419 MB at 1024^2) and then walks
// that tensor ~a dozen times (sigmoid, semantic einsum, >0 masks, mask scores, panoptic arg-max, areas).
// Here ONE kernel reads the low-resolution logits [Q, H4, W4] (13-26 MB, L2 resident) and, per tile of
// 8 x 16 output pixels, produces everything the task heads need:
//   * x_q = bilinear(logits_q) (F.interpolate align_corners=False semantics), s_q = sigmoid(x_q)
//   * sem_seg[c, p] = sum_q softmax(cls)[q, c] * s_q          (mma.sync fp16 x fp16 -> fp32, 144x112 x 112x128)
//   * panoptic arg-max over kept queries of score_q * s_q, and whether s >= 0.5 at the winner
//   * instance masks [x_q > 0] for the selected (query, class) slots, written in slot order
//   * per-query partial sums for the mask scores / panoptic areas (deterministic per-CTA partials)
// The full-resolution [Q, H, W] logits / sigmoid tensors are never materialised.
#include <type_traits>

#include "common.cuh"

namespace psalm {

constexpr int PP_TH = 8, PP_TW = 16, PP_PIX = PP_TH * PP_TW;   // output tile
constexpr int PP_QP = 112;                                      // queries padded to a multiple of 16
constexpr int PP_CP = 144;                                      // classes padded to a multiple of 16
constexpr int PP_SRC_MAX = 64;                                  // source taps per query per tile
constexpr int PP_SLD = 105;                                     // source row stride (odd: conflict-free)
constexpr int PP_BLD = PP_PIX + 8;                              // B operand row stride (halfs)
constexpr int PP_ALD = PP_QP + 8;                               // A operand row stride (halfs)

__device__ __forceinline__ void pp_ldsm_x4(uint32_t (&r)[4], const void* p) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void pp_ldsm_x4_t(uint32_t (&r)[4], const void* p) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void pp_mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

struct PostprocArgs {
  const void* logits;      // [Q, H4, W4]
  const __half* probsT;    // [PP_CP, PP_QP] fp16, zero padded (class-major), or null
  const float* wq;         // [Q] keep ? score : 0        (panoptic), or null
  const float* negq;       // [Q] keep ? 0 : -1
  const int* slot_query;   // [K] query index of every instance slot (-1 = unused), or null
  float* sem_seg;          // [ncls, H, W]
  float* inst_masks;       // [K, H, W]
  int* ids;                // [H, W]
  unsigned char* in_mask;  // [H, W]
  float* partials;         // [n_cta, Q, 5]: pos_cnt, pos_sig, ge_cnt, area, inter
  int Q, H4, W4, H, W, ncls, K;
  // composed resampling (sem_seg_postprocess, llava_phi.py:1399-1430): up-sample to the padded input size (Hp, Wp), crop
  // to the un-padded box (oh, ow), resize to the output size (H, W).  composed == 0: (H, W) is the up-sampled size itself.
  int composed, Hp, Wp, oh, ow;
};

template <typename T, bool COMPOSED>
__global__ void __launch_bounds__(256) postproc_fused_kernel(PostprocArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* src = reinterpret_cast<float*>(smem_raw);                               // [PP_SRC_MAX][PP_SLD]
  __half* Bs = reinterpret_cast<__half*>(src + PP_SRC_MAX * PP_SLD);             // [PP_QP][PP_BLD] (64*105*4 B is a multiple of 16)
  __half* As = Bs + PP_QP * PP_BLD;                                             // [PP_CP][PP_ALD]
  uint32_t* posbits = reinterpret_cast<uint32_t*>(As + PP_CP * PP_ALD);         // [Q pad 112][4]  x > 0
  uint32_t* gebits = posbits + PP_QP * 4;                                       // [Q pad 112][4]  x >= 0
  float* stats = reinterpret_cast<float*>(gebits + PP_QP * 4);                  // [112][5]
  float* amax_v = stats + PP_QP * 5;                                            // [2][128]
  int* amax_q = reinterpret_cast<int*>(amax_v + 2 * PP_PIX);                    // [2][128]
  float* amax_x = reinterpret_cast<float*>(amax_q + 2 * PP_PIX);                // [2][128]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Q = a.Q;
  const int ty0 = blockIdx.y * PP_TH, tx0 = blockIdx.x * PP_TW;
  // low-resolution scale: to the output size, or (composed) to the padded input size; second stage: crop -> output
  const float sh = COMPOSED ? (float)a.H4 / (float)a.Hp : (float)a.H4 / (float)a.H;
  const float sw = COMPOSED ? (float)a.W4 / (float)a.Wp : (float)a.W4 / (float)a.W;
  const float sh2 = COMPOSED ? (float)a.oh / (float)a.H : 1.f, sw2 = COMPOSED ? (float)a.ow / (float)a.W : 1.f;
  // source window of this tile (ATen area_pixel_compute_source_index, align_corners = false)
  auto srcf = [](float scale, int d) { const float s = scale * ((float)d + 0.5f) - 0.5f; return s < 0.f ? 0.f : s; };
  const int ylast = min(ty0 + PP_TH, a.H) - 1, xlast = min(tx0 + PP_TW, a.W) - 1;
  int sy0, sx0, sy1, sx1;
  if (COMPOSED) {   // output rows -> rows of the cropped up-sampled image -> low-resolution rows
    const int Y0 = (int)srcf(sh2, ty0), Y1 = min((int)srcf(sh2, ylast) + 1, a.oh - 1);
    const int X0 = (int)srcf(sw2, tx0), X1 = min((int)srcf(sw2, xlast) + 1, a.ow - 1);
    sy0 = (int)srcf(sh, Y0); sx0 = (int)srcf(sw, X0);
    sy1 = min((int)srcf(sh, Y1) + 1, a.H4 - 1); sx1 = min((int)srcf(sw, X1) + 1, a.W4 - 1);
  } else {
    sy0 = (int)srcf(sh, ty0); sx0 = (int)srcf(sw, tx0);
    sy1 = min((int)srcf(sh, ylast) + 1, a.H4 - 1); sx1 = min((int)srcf(sw, xlast) + 1, a.W4 - 1);
  }
  const int SR = sy1 - sy0 + 1, SC = sx1 - sx0 + 1;

  for (int i = tid; i < PP_QP * 5; i += 256) stats[i] = 0.f;
  for (int i = tid; i < SR * SC * Q; i += 256) {
    const int q = i / (SR * SC), r = i - q * (SR * SC);
    const int yy = sy0 + r / SC, xx = sx0 + r % SC;
    src[r * PP_SLD + q] = to_f32<T>(reinterpret_cast<const T*>(a.logits)[((size_t)q * a.H4 + yy) * a.W4 + xx]);
  }
  if (a.probsT)
    for (int i = tid; i < PP_CP * PP_QP / 8; i += 256) {
      const int row = i / (PP_QP / 8), c8 = (i % (PP_QP / 8)) * 8;
      *reinterpret_cast<uint4*>(&As[row * PP_ALD + c8]) = __ldg(reinterpret_cast<const uint4*>(a.probsT + row * PP_QP + c8));
    }
  // zero the padded query rows of the B operand
  for (int i = tid; i < (PP_QP - Q) * PP_BLD; i += 256) Bs[Q * PP_BLD + i] = __float2half(0.f);
  __syncthreads();

  // ---- phase 1: thread = (pixel p, query parity); x, sigmoid, per-query warp statistics, running arg-max
  const int p = tid & (PP_PIX - 1), half = tid >> 7;
  const int py = ty0 + p / PP_TW, px = tx0 + p % PP_TW;
  const bool inb = py < a.H && px < a.W;
  const int pyc = py < a.H ? py : a.H - 1, pxc = px < a.W ? px : a.W - 1;
  // separable taps: NT rows x NT columns of the source window with weights wy[i] * wx[j]; one stage: 2 x 2, composed: the
  // two rows (columns) of the cropped up-sampled image each expand to two low-resolution rows (columns): 4 x 4
  constexpr int NT = COMPOSED ? 4 : 2;
  int oy[NT], ox[NT];
  float wy[NT], wx[NT];
  {
    auto taps = [&](float scale, int n_lo, int lo0, float coord, int& i0, int& i1, float& w0, float& w1) {
      const int c0 = (int)coord;
      const int c1 = c0 + (c0 < n_lo - 1 ? 1 : 0);
      const float l = coord - (float)c0;
      i0 = c0 - lo0; i1 = c1 - lo0; w0 = 1.f - l; w1 = l;
    };
    if (COMPOSED) {
      const float fY = srcf(sh2, pyc), fX = srcf(sw2, pxc);
      const int Y0 = (int)fY, X0 = (int)fX;
      const int Y1 = Y0 + (Y0 < a.oh - 1 ? 1 : 0), X1 = X0 + (X0 < a.ow - 1 ? 1 : 0);
      const float lY = fY - (float)Y0, lX = fX - (float)X0;
      float w0, w1;
      taps(sh, a.H4, sy0, srcf(sh, Y0), oy[0], oy[1], w0, w1); wy[0] = (1.f - lY) * w0; wy[1] = (1.f - lY) * w1;
      taps(sh, a.H4, sy0, srcf(sh, Y1), oy[2], oy[3], w0, w1); wy[2] = lY * w0; wy[3] = lY * w1;
      taps(sw, a.W4, sx0, srcf(sw, X0), ox[0], ox[1], w0, w1); wx[0] = (1.f - lX) * w0; wx[1] = (1.f - lX) * w1;
      taps(sw, a.W4, sx0, srcf(sw, X1), ox[2], ox[3], w0, w1); wx[2] = lX * w0; wx[3] = lX * w1;
    } else {
      taps(sh, a.H4, sy0, srcf(sh, pyc), oy[0], oy[1], wy[0], wy[1]);
      taps(sw, a.W4, sx0, srcf(sw, pxc), ox[0], ox[1], wx[0], wx[1]);
    }
  }
  int off[NT][NT];
#pragma unroll
  for (int i = 0; i < NT; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j) off[i][j] = (oy[i] * SC + ox[j]) * PP_SLD;
  float best_v = -2.f, best_x = 0.f;
  int best_q = 0;
  const int wsub = warp & 3;   // which 32-pixel group of the tile this warp covers
  const float* wqp = a.wq;
  const float* ngp = a.negq;
  for (int q = half; q < Q; q += 2) {
    float x;
    if (COMPOSED) {
      x = 0.f;
#pragma unroll
      for (int i = 0; i < NT; ++i) {
        float r = 0.f;
#pragma unroll
        for (int j = 0; j < NT; ++j) r = fmaf(wx[j], src[off[i][j] + q], r);
        x = fmaf(wy[i], r, x);
      }
    } else {
      x = wy[0] * (wx[0] * src[off[0][0] + q] + wx[1] * src[off[0][1] + q]) + wy[1] * (wx[0] * src[off[1][0] + q] + wx[1] * src[off[1][1] + q]);
    }
    const float s = __fdividef(1.f, 1.f + __expf(-x));
    Bs[q * PP_BLD + p] = __float2half(s);
    const uint32_t mpos = __ballot_sync(0xffffffffu, inb && x > 0.f);
    const uint32_t mge = __ballot_sync(0xffffffffu, inb && x >= 0.f);
    if (lane == 0) {
      posbits[q * 4 + wsub] = mpos;
      gebits[q * 4 + wsub] = mge;
    }
    if (wqp) {
      const float v = fmaf(__ldg(wqp + q), s, __ldg(ngp + q));
      if (v > best_v) { best_v = v; best_q = q; best_x = x; }   // strict >: first maximum wins, like argmax
    }
  }
  if (a.wq) {
    amax_v[half * PP_PIX + p] = best_v;
    amax_q[half * PP_PIX + p] = best_q;
    amax_x[half * PP_PIX + p] = best_x;
  }
  __syncthreads();
  if (tid >= 128 && tid - 128 < Q) {   // per-query statistics of this tile (second half of the CTA; first half does the arg-max)
    const int q = tid - 128;
    float cnt = 0.f, gcnt = 0.f, ps = 0.f;
#pragma unroll
    for (int wd = 0; wd < 4; ++wd) {
      const uint32_t m = posbits[q * 4 + wd];
      cnt += (float)__popc(m);
      gcnt += (float)__popc(gebits[q * 4 + wd]);
      const __half2* row = reinterpret_cast<const __half2*>(&Bs[q * PP_BLD + wd * 32]);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float2 f = __half22float2(row[i]);
        ps += ((m >> (2 * i)) & 1u) ? f.x : 0.f;
        ps += ((m >> (2 * i + 1)) & 1u) ? f.y : 0.f;
      }
    }
    stats[q * 5 + 0] = cnt;
    stats[q * 5 + 1] = ps;
    stats[q * 5 + 2] = gcnt;
  }
  if (a.wq && tid < PP_PIX) {
    float v0 = amax_v[p], v1 = amax_v[PP_PIX + p];
    int q0 = amax_q[p], q1 = amax_q[PP_PIX + p];
    float x0v = amax_x[p], x1v = amax_x[PP_PIX + p];
    // even queries (half 0) vs odd queries (half 1): lower index wins ties
    const bool take1 = (v1 > v0) || (v1 == v0 && q1 < q0);
    const int qb = take1 ? q1 : q0;
    const float xb = take1 ? x1v : x0v;
    if (inb) {
      a.ids[(size_t)py * a.W + px] = qb;
      const bool im = xb >= 0.f;
      a.in_mask[(size_t)py * a.W + px] = im ? 1 : 0;
      atomicAdd(&stats[qb * 5 + 3], 1.f);
      if (im) atomicAdd(&stats[qb * 5 + 4], 1.f);
    }
  }
  // ---- phase 2: semantic map, D[class, pixel] = probsT[class, q] * sig[q, pixel]
  if (a.probsT) {
    float acc[PP_CP / 16][2][4];
#pragma unroll
    for (int i = 0; i < PP_CP / 16; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) acc[i][j][0] = acc[i][j][1] = acc[i][j][2] = acc[i][j][3] = 0.f;
    const int mi = lane >> 3;
#pragma unroll
    for (int ks = 0; ks < PP_QP / 16; ++ks) {
      uint32_t b[4];   // two 8-pixel n-tiles of this warp's 16 pixels
      pp_ldsm_x4_t(b, &Bs[(ks * 16 + (lane & 7) + (mi & 1) * 8) * PP_BLD + warp * 16 + (mi >> 1) * 8]);
#pragma unroll
      for (int mt = 0; mt < PP_CP / 16; ++mt) {
        uint32_t af[4];
        pp_ldsm_x4(af, &As[(mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * PP_ALD + ks * 16 + (lane >> 4) * 8]);
        pp_mma(acc[mt][0], af, b[0], b[1]);
        pp_mma(acc[mt][1], af, b[2], b[3]);
      }
    }
    const int g = lane >> 2, t4 = lane & 3;
#pragma unroll
    for (int mt = 0; mt < PP_CP / 16; ++mt)
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const int c = mt * 16 + g + r * 8;
          const int pp = warp * 16 + j * 8 + 2 * t4;
          const int yy = ty0 + pp / PP_TW, xx = tx0 + pp % PP_TW;
          if (c < a.ncls && yy < a.H && xx < a.W) {
            float* dst = a.sem_seg + ((size_t)c * a.H + yy) * a.W + xx;
            if (xx + 1 < a.W && ((size_t)dst & 7) == 0) *reinterpret_cast<float2*>(dst) = make_float2(acc[mt][j][2 * r], acc[mt][j][2 * r + 1]);
            else { dst[0] = acc[mt][j][2 * r]; if (xx + 1 < a.W) dst[1] = acc[mt][j][2 * r + 1]; }
          }
        }
  }
  __syncthreads();
  // ---- phase 3: instance masks in slot order
  if (a.slot_query) {
    for (int k = half; k < a.K; k += 2) {
      const int q = a.slot_query[k];
      if (q < 0 || !inb) continue;
      const bool bit = (posbits[q * 4 + (p >> 5)] >> (p & 31)) & 1u;
      a.inst_masks[((size_t)k * a.H + py) * a.W + px] = bit ? 1.f : 0.f;
    }
  }
  // ---- per-CTA partial statistics (deterministic final reduction on the host side of the call)
  float* part = a.partials + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * Q * 5;
  for (int i = tid; i < Q * 5; i += 256) part[i] = stats[i];
}

constexpr size_t pp_smem_bytes() {
  return sizeof(float) * (PP_SRC_MAX * PP_SLD) + sizeof(__half) * (PP_QP * PP_BLD + PP_CP * PP_ALD) +
         sizeof(uint32_t) * PP_QP * 8 + sizeof(float) * PP_QP * 5 + (sizeof(float) * 2 + sizeof(int)) * 2 * PP_PIX;
}

}  // namespace psalm

namespace psalm {
// postproc_fast.cu
bool postproc_fast_ok(int Q, int H4, int W4, int H, int W, int ncls, int K, int dtype);
int postproc_fast_ctas(int H, int W);
int postproc_fast_launch(const void* logits, const void* probsT_f16, const float* wq, const float* negq,
                         const int* slot_query, float* sem_seg, float* inst_masks, int* ids, unsigned char* in_mask,
                         float* partials, int Q, int H4, int W4, int H, int W, int ncls, int K, int dtype,
                         cudaStream_t st);
}  // namespace psalm

using namespace psalm;

static int launch_generic(const PostprocArgs& a, bool composed, int dtype, cudaStream_t st) {
  dim3 grid((a.W + PP_TW - 1) / PP_TW, (a.H + PP_TH - 1) / PP_TH);
  const size_t smem = pp_smem_bytes();
  cudaError_t e = cudaSuccess;
#define PPL(T, C)                                                                                               \
  e = cudaFuncSetAttribute(postproc_fused_kernel<T, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
  if (e == cudaSuccess) postproc_fused_kernel<T, C><<<grid, 256, smem, st>>>(a)
#define PPD(T) do { if (composed) { PPL(T, true); } else { PPL(T, false); } } while (0)
  if (dtype == PSALM_F32) PPD(float);
  else if (dtype == PSALM_F16) PPD(__half);
  else if (dtype == PSALM_BF16) PPD(__nv_bfloat16);
  else { set_error("postproc_fused: unknown dtype %d", dtype); return PSALM_E_ARG; }
#undef PPD
#undef PPL
  if (e != cudaSuccess) { set_error("postproc_fused: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return PSALM_E_CUDA; }
  return check_launch("postproc_fused_kernel");
}

static int g_postproc_impl = 0;   // 0 auto, 1 generic SIMT-blend kernel, 2 tensor-core kernel

extern "C" int psalm_set_postproc_impl(int impl) {
  PSALM_REQUIRE(impl >= 0 && impl <= 2, "set_postproc_impl: 0 (auto), 1 (generic) or 2 (tensor-core)");
  g_postproc_impl = impl;
  return PSALM_OK;
}

static bool use_fast(int Q, int H4, int W4, int H, int W, int ncls, int K, int dtype) {
  return g_postproc_impl != 1 && postproc_fast_ok(Q, H4, W4, H, W, ncls, K, dtype);
}

extern "C" int psalm_postproc_partials(int Q, int H4, int W4, int H, int W, int ncls, int K, int dtype, int* rows) {
  PSALM_REQUIRE(rows && H > 0 && W > 0, "postproc_partials: bad arguments");
  if (g_postproc_impl == 2 && !postproc_fast_ok(Q, H4, W4, H, W, ncls, K, dtype)) {
    set_error("postproc_partials: tensor-core path forced but unsupported (needs 16-bit logits, x1..x8 power-of-two "
              "up-sampling, Q <= 112, ncls <= 144)");
    return PSALM_E_ARG;
  }
  *rows = use_fast(Q, H4, W4, H, W, ncls, K, dtype) ? postproc_fast_ctas(H, W)
                                                     : ((W + PP_TW - 1) / PP_TW) * ((H + PP_TH - 1) / PP_TH);
  return PSALM_OK;
}

extern "C" int psalm_postproc_fused(const void* logits, const void* probsT_f16, const float* wq, const float* negq,
                                    const int* slot_query, float* sem_seg, float* inst_masks, int* ids,
                                    unsigned char* in_mask, float* partials, int Q, int H4, int W4, int H, int W,
                                    int ncls, int K, int dtype, void* stream) {
  PSALM_REQUIRE(logits && partials, "postproc_fused: null pointer");
  PSALM_REQUIRE((probsT_f16 == nullptr) == (sem_seg == nullptr), "postproc_fused: probsT and sem_seg go together");
  PSALM_REQUIRE((wq == nullptr) == (ids == nullptr) && (wq == nullptr) == (negq == nullptr) && (wq == nullptr) == (in_mask == nullptr),
                "postproc_fused: wq / negq / ids / in_mask go together");
  PSALM_REQUIRE((slot_query == nullptr) == (inst_masks == nullptr), "postproc_fused: slot_query and inst_masks go together");
  if (g_postproc_impl == 2)
    PSALM_REQUIRE(postproc_fast_ok(Q, H4, W4, H, W, ncls, K, dtype), "postproc_fused: tensor-core path forced but unsupported");
  if (use_fast(Q, H4, W4, H, W, ncls, K, dtype))
    return postproc_fast_launch(logits, probsT_f16, wq, negq, slot_query, sem_seg, inst_masks, ids, in_mask, partials,
                                Q, H4, W4, H, W, ncls, K, dtype, (cudaStream_t)stream);
  PSALM_REQUIRE(Q > 0 && Q <= 104 && ncls <= PP_CP, "postproc_fused: Q=%d (max 104) / ncls=%d (max %d) unsupported", Q, ncls, PP_CP);
  // source taps per tile must fit the shared-memory window: (TH*scale + 2) x (TW*scale + 2)
  const float sh = (float)H4 / (float)H, sw = (float)W4 / (float)W;
  const int SR = (int)(PP_TH * sh) + 3, SC = (int)(PP_TW * sw) + 3;
  PSALM_REQUIRE(SR * SC <= PP_SRC_MAX, "postproc_fused: resize factor too small for the fused path (%dx%d source taps per tile)", SR, SC);
  PostprocArgs a{logits, (const __half*)probsT_f16, wq, negq, slot_query, sem_seg, inst_masks, ids, in_mask, partials,
                 Q, H4, W4, H, W, ncls, K, 0, H, W, H, W};
  return launch_generic(a, false, dtype, (cudaStream_t)stream);
}

// source taps of one output tile for the composed (up-sample -> crop -> resize) path
static void composed_window(int H4, int W4, int Hp, int Wp, int oh, int ow, int H, int W, int& SR, int& SC) {
  const float r2h = (float)oh / (float)H, r2w = (float)ow / (float)W, r1h = (float)H4 / (float)Hp, r1w = (float)W4 / (float)Wp;
  SR = (int)(((float)PP_TH * r2h + 2.f) * r1h) + 3;
  SC = (int)(((float)PP_TW * r2w + 2.f) * r1w) + 3;
}

extern "C" int psalm_postproc_crop_supported(int Q, int H4, int W4, int Hp, int Wp, int oh, int ow, int H, int W, int ncls) {
  if (Q <= 0 || Q > 104 || ncls > PP_CP || oh <= 0 || ow <= 0 || oh > Hp || ow > Wp || H <= 0 || W <= 0) return 0;
  int SR, SC;
  composed_window(H4, W4, Hp, Wp, oh, ow, H, W, SR, SC);
  return SR * SC <= PP_SRC_MAX ? 1 : 0;
}

extern "C" int psalm_postproc_crop_partials(int H, int W, int* rows) {
  PSALM_REQUIRE(rows && H > 0 && W > 0, "postproc_crop_partials: bad arguments");
  *rows = ((W + PP_TW - 1) / PP_TW) * ((H + PP_TH - 1) / PP_TH);
  return PSALM_OK;
}

extern "C" int psalm_postproc_fused_crop(const void* logits, const void* probsT_f16, const float* wq, const float* negq,
                                         const int* slot_query, float* sem_seg, float* inst_masks, int* ids,
                                         unsigned char* in_mask, float* partials, int Q, int H4, int W4, int Hp, int Wp,
                                         int oh, int ow, int H, int W, int ncls, int K, int dtype, void* stream) {
  PSALM_REQUIRE(logits && partials, "postproc_fused_crop: null pointer");
  PSALM_REQUIRE((probsT_f16 == nullptr) == (sem_seg == nullptr), "postproc_fused_crop: probsT and sem_seg go together");
  PSALM_REQUIRE((wq == nullptr) == (ids == nullptr) && (wq == nullptr) == (negq == nullptr) && (wq == nullptr) == (in_mask == nullptr),
                "postproc_fused_crop: wq / negq / ids / in_mask go together");
  PSALM_REQUIRE((slot_query == nullptr) == (inst_masks == nullptr), "postproc_fused_crop: slot_query and inst_masks go together");
  PSALM_REQUIRE(psalm_postproc_crop_supported(Q, H4, W4, Hp, Wp, oh, ow, H, W, ncls),
                "postproc_fused_crop: unsupported geometry (Q=%d, crop %dx%d of %dx%d -> %dx%d)", Q, oh, ow, Hp, Wp, H, W);
  PostprocArgs a{logits, (const __half*)probsT_f16, wq, negq, slot_query, sem_seg, inst_masks, ids, in_mask, partials,
                 Q, H4, W4, H, W, ncls, K, 1, Hp, Wp, oh, ow};
  return launch_generic(a, true, dtype, (cudaStream_t)stream);
}
