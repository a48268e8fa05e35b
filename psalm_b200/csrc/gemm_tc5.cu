// Linear layers with a fused epilogue on the 5th-generation tensor cores: out = epilogue(A · Wᵀ + bias), 16-bit storage,
// fp32 accumulation in tensor memory.
//
// Replaces, on the 16-bit path, the library GEMM + separate elementwise pass of
//   * Swin `Mlp.fc1` + exact-erf `nn.GELU` (swin_trans.py:37-44, 24 blocks): the stand-alone GELU pass re-reads and
//     re-writes the 4C-wide activation (268 MB each way per block at stage 0, B = 4) - 5.2 % of the step;
//   * MSDeformAttn `value_proj` + the [B,S,8,32] -> [B,8,S,32] head-major copy (ops/modules/ms_deform_attn.py:95-99
//     + our layout): the epilogue stores each 32-column chunk (= one head) where the sampling
//     kernel wants it.
//
// Persistent CTA per SM, 128 x 256 output tiles, BK = 64 (one SWIZZLE_128B atom per row):
//   warp 16     TMA producer: A box [128 x 64] and W box [256 x 64] per k-block into a 3-stage ring (48 KB per stage)
//   warp 17     one elected lane issues 4 x tcgen05.mma M128 N256 K16 per k-block, commits the stage back to the producer
//               and, after the last k-block, the accumulator to the epilogue
//   warps 0-15  epilogue: thread = output row (TMEM lane), warpgroup g takes columns 64 g .. 64 g + 63 in chunks of 32:
//               tcgen05.ld -> + bias -> GELU -> 16-bit pack -> swizzled staging box in shared memory -> one TMA store per
//               warp ([32 rows x 128 B], or two [32 x 64 B] head boxes) so that L2 sees whole lines, not the 16-byte pieces
//               of 32 different rows a thread-per-row st.global would send (2.2 us of LSU time per tile, measured as
//               the limiter of the first version)
// Two accumulators (2 x 256 TMEM columns) so that the MMAs of tile i + 1 run under the epilogue of tile i: at K = 128 the
// epilogue (32 K GELUs per tile) is the longest stage, not the tensor pipe and not HBM.
//
// GELU: 0.5 x (1 + erf(x / sqrt 2)) = max(x, 0) - |x| E / 2 with E = erfc(|x| / sqrt 2); E / 2 = 2^q(t), q a degree-7
// polynomial fitted to log2(erfc(z) / 2) on z in [0, 5], t = sat(z / 5) (one FMUL.SAT; beyond z = 5, erfc < 2e-12).
// Relative error of the result < 2.5e-5 for |x| < 7, absolute < 1.5e-6 everywhere (checked against float64 erf, fp32
// Horner emulated) - two orders below the 16-bit rounding of the output - at one MUFU and 3.5 packed FMAs per element
// (fma.rn.f32x2) instead of libm's erff.  The first version of this epilogue was issue-bound at Swin stage 0
// (profiles/r2d_gemm_tc5_*: 13.4 instructions per element, 68 % issue slots busy); this one needs ~9.
#include <cuda.h>

#include <type_traits>

#include "common.cuh"

namespace psalm {

namespace gt {
constexpr int BM = 128, BN = 256, BK = 64, NS = 3;
constexpr int A_BYTES = BM * BK * 2, W_BYTES = BN * BK * 2, STAGE = A_BYTES + W_BYTES;
constexpr int EPI_WARPS = 16, THREADS = (EPI_WARPS + 2) * 32;
constexpr int EPI_COLS = BN / (EPI_WARPS / 4);   // columns per epilogue warp (4 warps cover the 128 TMEM lanes)
constexpr int STG_WARP = 32 * EPI_COLS * 2;       // output staging of one epilogue warp: [32 rows x 64 columns] = 4 KB
constexpr int BIAS_WARP = EPI_COLS * 4;           // the warp's bias slice as fp32
constexpr size_t SMEM = 1024 + (size_t)NS * STAGE + (size_t)EPI_WARPS * (STG_WARP + BIAS_WARP) + 256;
}  // namespace gt

struct GtParams {
  const void* bias;   // [N] storage dtype, or null
  void* out;
  int M, N, K;
  int epilogue;       // 0 bias, 1 bias + GELU(erf), 2 bias + head-major store [M / S, N / 32, S, 32]
  int S;              // rows per image (epilogue 2)
  int tiles_m, tiles_n;
  int tn_shift;       // log2(tiles_n) when it is a power of two, else -1
};

__device__ __forceinline__ uint32_t gt_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t gt_desc(uint32_t smem_addr) {   // K-major SWIZZLE_128B, 1024 B per 8-row group
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fff);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
template <typename T>
__device__ __forceinline__ uint32_t gt_idesc() {
  const uint32_t fmt = std::is_same<T, __nv_bfloat16>::value ? 1u : 0u;
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(gt::BN >> 3) << 17) | ((uint32_t)(gt::BM >> 4) << 24);
}
__device__ __forceinline__ void gt_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(gt_u32(bar)), "r"(count));
}
#define gt_mbar_wait(bar, parity)                                                                       \
  do {                                                                                                  \
    uint32_t done_ = 0, spins_ = 0;                                                                     \
    while (!done_) {                                                                                    \
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"   \
                   "selp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done_) : "r"(gt_u32(bar)), "r"((uint32_t)(parity)) : "memory"); \
      if (++spins_ > (1u << 26)) __trap(); /* never hang the GPU on a protocol bug */                   \
    }                                                                                                   \
  } while (0)
__device__ __forceinline__ void gt_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(gt_u32(bar)) : "memory");
}
__device__ __forceinline__ void gt_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(gt_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void gt_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(gt_u32(bar)) : "memory");
}
__device__ __forceinline__ void gt_tma_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(
          gt_u32(dst)),
      "l"(map), "r"(c0), "r"(c1), "r"(gt_u32(bar))
      : "memory");
}
__device__ __forceinline__ void gt_tma_store_2d(const CUtensorMap* map, int c0, int c1, const void* src) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];\n" ::"l"(map), "r"(c0), "r"(c1),
               "r"(gt_u32(src))
               : "memory");
}
__device__ __forceinline__ void gt_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void gt_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,"
      "%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}

typedef unsigned long long gt_f2;   // two packed fp32 lanes (fma.rn.f32x2 operands)
__device__ __forceinline__ gt_f2 gt_pk(float a, float b) {
  gt_f2 r;
  asm("mov.b64 %0, {%1, %2};\n" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void gt_unpk(gt_f2 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;\n" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ gt_f2 gt_fma2(gt_f2 a, gt_f2 b, gt_f2 c) {
  gt_f2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;\n" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ void gt_sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ gt_f2 gt_add2(gt_f2 a, gt_f2 b) {
  gt_f2 r;
  asm("add.rn.f32x2 %0, %1, %2;\n" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
// GELU of two values; see the header comment.
__device__ __forceinline__ void gt_gelu2(float& x0, float& x1) {
  const float a0 = fabsf(x0), a1 = fabsf(x1);
  const gt_f2 t = gt_pk(__saturatef(a0 * (0.70710678f * 0.2f)), __saturatef(a1 * (0.70710678f * 0.2f)));
  gt_f2 q = gt_pk(-1.401790814e+00f, -1.401790814e+00f);
  q = gt_fma2(q, t, gt_pk(7.022554923e+00f, 7.022554923e+00f));
  q = gt_fma2(q, t, gt_pk(-1.563424726e+01f, -1.563424726e+01f));
  q = gt_fma2(q, t, gt_pk(2.078584664e+01f, 2.078584664e+01f));
  q = gt_fma2(q, t, gt_pk(-1.893136839e+01f, -1.893136839e+01f));
  q = gt_fma2(q, t, gt_pk(-2.294412836e+01f, -2.294412836e+01f));
  q = gt_fma2(q, t, gt_pk(-8.139466606e+00f, -8.139466606e+00f));
  q = gt_fma2(q, t, gt_pk(-1.000004739e+00f, -1.000004739e+00f));
  float q0, q1, h0, h1;
  gt_unpk(q, q0, q1);
  asm("ex2.approx.ftz.f32 %0, %1;\n" : "=f"(h0) : "f"(q0));
  asm("ex2.approx.ftz.f32 %0, %1;\n" : "=f"(h1) : "f"(q1));
  x0 = fmaf(-a0, h0, fmaxf(x0, 0.f));
  x1 = fmaf(-a1, h1, fmaxf(x1, 0.f));
}

template <typename T>
__device__ __forceinline__ float2 load_pair_f32(const T* p);
template <>
__device__ __forceinline__ float2 load_pair_f32<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p));
}
template <>
__device__ __forceinline__ float2 load_pair_f32<__half>(const __half* p) {
  return __half22float2(*reinterpret_cast<const __half2*>(p));
}
template <typename T>
__device__ __forceinline__ uint32_t gt_pack(float a, float b);
template <>
__device__ __forceinline__ uint32_t gt_pack<__nv_bfloat16>(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
template <>
__device__ __forceinline__ uint32_t gt_pack<__half>(float a, float b) {
  __half2 v = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

template <typename T, int EPI>
__global__ void __launch_bounds__(gt::THREADS, 1)
linear_tc5_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapW,
                  const __grid_constant__ CUtensorMap mapO, GtParams p) {
  using namespace gt;
  extern __shared__ unsigned char gt_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(gt_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* stg_all = base + (size_t)NS * STAGE;
  float* bias_all = reinterpret_cast<float*>(stg_all + (size_t)EPI_WARPS * STG_WARP);
  uint64_t* bars = reinterpret_cast<uint64_t*>(stg_all + (size_t)EPI_WARPS * (STG_WARP + BIAS_WARP));
  uint64_t* full = bars;               // [NS] TMA landed
  uint64_t* empty = bars + NS;         // [NS] the MMAs reading the stage have retired
  uint64_t* acc_full = bars + 2 * NS;  // [2]  accumulator complete
  uint64_t* acc_free = acc_full + 2;   // [2]  accumulator drained by the 8 epilogue warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_free + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_tiles = p.tiles_m * p.tiles_n;
  const int kblocks = p.K / BK;

  if (warp == EPI_WARPS + 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"(gt_u32(tmem_slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
  }
  if (tid == 0) {
    for (int s = 0; s < NS; ++s) {
      gt_mbar_init(&full[s], 1);
      gt_mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      gt_mbar_init(&acc_full[a], 1);
      gt_mbar_init(&acc_free[a], EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  if (warp == EPI_WARPS) {
    // ------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t it = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int tm = p.tn_shift >= 0 ? t >> p.tn_shift : t / p.tiles_n, tn = t - tm * p.tiles_n;
        for (int kb = 0; kb < kblocks; ++kb, ++it) {
          const int s = it % NS;
          if (it >= NS) gt_mbar_wait(&empty[s], ((it / NS) - 1) & 1);
          unsigned char* st = base + (size_t)s * STAGE;
          gt_mbar_expect_tx(&full[s], STAGE);
          gt_tma_2d(st, &mapA, kb * BK, tm * BM, &full[s]);
          gt_tma_2d(st + A_BYTES, &mapW, kb * BK, tn * BN, &full[s]);
        }
      }
    }
  } else if (warp == EPI_WARPS + 1) {
    // ------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = gt_idesc<T>();
      uint32_t it = 0, li = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++li) {
        const int a = li & 1;
        if (li >= 2) gt_mbar_wait(&acc_free[a], ((li >> 1) - 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        const uint32_t d = tmem + a * BN;
        for (int kb = 0; kb < kblocks; ++kb, ++it) {
          const int s = it % NS;
          gt_mbar_wait(&full[s], (it / NS) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
          const uint32_t sa = gt_u32(base + (size_t)s * STAGE), sw = sa + A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            gt_mma(d, gt_desc(sa + k * 32), gt_desc(sw + k * 32), idesc, (kb | k) ? 1u : 0u);
          gt_commit(&empty[s]);
        }
        gt_commit(&acc_full[a]);
      }
    }
  } else {
    // ------------------------------------------------ epilogue
    const int q4 = warp & 3, g = warp >> 2;   // TMEM lane quarter (hardware: warp % 4), column group
    unsigned char* stg = stg_all + (size_t)warp * STG_WARP;
    float* bias_s = bias_all + warp * EPI_COLS;
    uint32_t li = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++li) {
      const int tm = p.tn_shift >= 0 ? t >> p.tn_shift : t / p.tiles_n, tn = t - tm * p.tiles_n;
      const int a = li & 1;
      gt_mbar_wait(&acc_full[a], (li >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      const uint32_t taddr = tmem + ((uint32_t)(q4 * 32) << 16) + a * BN + g * EPI_COLS;
      const int col0 = tn * BN + g * EPI_COLS;
      const long long row0 = (long long)tm * BM + q4 * 32;      // first row of this warp's box
      if (li > 0) {                                             // the previous store has finished reading the staging box
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
      }
      {   // this warp's 64 bias values as fp32 (lane l converts columns 2 l, 2 l + 1)
        float2 b2 = make_float2(0.f, 0.f);
        if (p.bias) b2 = load_pair_f32<T>(reinterpret_cast<const T*>(p.bias) + col0 + 2 * lane);
        asm volatile("st.shared.v2.f32 [%0], {%1, %2};\n" ::"r"(gt_u32(bias_s) + lane * 8), "f"(b2.x), "f"(b2.y) : "memory");
      }
      __syncwarp();
#pragma unroll 1
      for (int c = 0; c < EPI_COLS / 32; ++c) {
        uint32_t v[32];
        gt_ld32(taddr + c * 32, v);
        if (c == EPI_COLS / 32 - 1) {   // the accumulator columns of this warp are in registers: hand it back to the MMA warp
          asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
          __syncwarp();
          if (lane == 0) gt_mbar_arrive(&acc_free[a]);
        }
        uint32_t o[16];
        const uint32_t b4 = gt_u32(bias_s + c * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 bb;                                // broadcast read: columns 4 j .. 4 j + 3 of the chunk
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];\n" : "=f"(bb.x), "=f"(bb.y), "=f"(bb.z), "=f"(bb.w) : "r"(b4 + j * 16));
          float x0, x1, x2, x3;
          gt_unpk(gt_add2(gt_pk(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1])), gt_pk(bb.x, bb.y)), x0, x1);
          gt_unpk(gt_add2(gt_pk(__uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3])), gt_pk(bb.z, bb.w)), x2, x3);
          if (EPI == 1) {
            gt_gelu2(x0, x1);
            gt_gelu2(x2, x3);
          }
          o[2 * j] = gt_pack<T>(x0, x1);
          o[2 * j + 1] = gt_pack<T>(x2, x3);
        }
        if (EPI == 2) {   // box c = one head: [32 rows x 64 B], SWIZZLE_64B (16-byte chunk ^ bits 1-2 of the row)
          const uint32_t bx = gt_u32(stg) + c * 2048 + lane * 64;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            gt_sts128(bx + ((j ^ ((lane >> 1) & 3)) << 4), o[j * 4], o[j * 4 + 1], o[j * 4 + 2], o[j * 4 + 3]);
        } else {          // one box [32 rows x 128 B], SWIZZLE_128B (16-byte chunk ^ row % 8)
          const uint32_t bx = gt_u32(stg) + lane * 128;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            gt_sts128(bx + (((c * 4 + j) ^ (lane & 7)) << 4), o[j * 4], o[j * 4 + 1], o[j * 4 + 2], o[j * 4 + 3]);
        }
      }
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
      __syncwarp();
      if (lane == 0 && row0 < p.M) {   // rows beyond M are clipped by the tensor map
        if (EPI == 2) {
          const long long img = row0 / p.S, srow = row0 - img * p.S;
#pragma unroll
          for (int c = 0; c < EPI_COLS / 32; ++c) {
            const long long orow = (img * (p.N / 32) + (col0 / 32 + c)) * p.S + srow;
            gt_tma_store_2d(&mapO, 0, (int)orow, stg + c * 2048);
          }
        } else {
          gt_tma_store_2d(&mapO, col0, (int)row0, stg);
        }
        asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == EPI_WARPS + 1) {
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tmem));
  }
}

// ---- host ------------------------------------------------------------------------------------------------------
typedef CUresult (*GtEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static GtEncodeFn gt_encode_fn() {
  static GtEncodeFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<GtEncodeFn>(p);
  }
  return fn;
}
static bool gt_make_map(CUtensorMap* map, const void* base, long long rows, long long cols, long long row_stride_elems,
                        int box_cols, int box_rows, CUtensorMapSwizzle swz, int dtype) {
  GtEncodeFn fn = gt_encode_fn();
  if (!fn) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)row_stride_elems * 2};
  const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  return fn(map, dtype == PSALM_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
            const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
static int gt_sm_count() {
  static PerDevice cache;
  const int d = PerDevice::dev();
  if (cache.first() || cache.v[d] == 0) cudaDeviceGetAttribute(&cache.v[d], cudaDevAttrMultiProcessorCount, d);
  return cache.v[d] > 0 ? cache.v[d] : 148;
}

template <typename T, int EPI>
static cudaError_t gt_launch(const CUtensorMap& ma, const CUtensorMap& mw, const CUtensorMap& mo, const GtParams& p, int grid,
                             cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(linear_tc5_kernel<T, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gt::SMEM);
  if (e != cudaSuccess) return e;
  linear_tc5_kernel<T, EPI><<<grid, gt::THREADS, gt::SMEM, st>>>(ma, mw, mo, p);
  return cudaSuccess;
}

}  // namespace psalm

using namespace psalm;

extern "C" int psalm_linear_fused_supported(long long M, int N, int K, int epilogue, long long rows_per_image, int dtype) {
  if (dtype != PSALM_BF16 && dtype != PSALM_F16) return 0;
  if (M <= 0 || N <= 0 || K <= 0 || N % gt::BN || K % gt::BK || M > (1ll << 31) - gt::BM) return 0;
  if (epilogue < 0 || epilogue > 2) return 0;
  if (epilogue == 2 && (rows_per_image <= 0 || M % rows_per_image || rows_per_image % 32)) return 0;   // a warp's 32 rows: one image
  return 1;
}

extern "C" int psalm_linear_fused(const void* a, long long a_row_stride, const void* w, const void* bias, void* out, long long M,
                                  int N, int K, int epilogue, long long rows_per_image, int dtype, void* stream) {
  PSALM_REQUIRE(a && w && out, "linear_fused: null pointer");
  PSALM_REQUIRE(psalm_linear_fused_supported(M, N, K, epilogue, rows_per_image, dtype),
                "linear_fused: unsupported shape M=%lld N=%d K=%d epilogue=%d (16-bit storage, N %% 256 == 0, K %% 64 == 0)", M, N, K,
                epilogue);
  PSALM_REQUIRE(a_row_stride >= K && a_row_stride % 8 == 0, "linear_fused: a_row_stride must be >= K and a multiple of 8 elements");
  PSALM_REQUIRE((reinterpret_cast<uintptr_t>(a) & 15) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(out) & 15) == 0 && (!bias || (reinterpret_cast<uintptr_t>(bias) & 15) == 0),
                "linear_fused: pointers must be 16-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  GtParams p;
  p.bias = bias; p.out = out; p.M = (int)M; p.N = N; p.K = K; p.epilogue = epilogue; p.S = (int)(epilogue == 2 ? rows_per_image : 1);
  p.tiles_m = (int)((M + gt::BM - 1) / gt::BM);
  p.tiles_n = N / gt::BN;
  p.tn_shift = -1;
  for (int sft = 0; sft < 16; ++sft)
    if ((1 << sft) == p.tiles_n) p.tn_shift = sft;
  CUtensorMap ma, mw, mo;
  const bool maps_ok =
      gt_make_map(&ma, a, M, K, a_row_stride, gt::BK, gt::BM, CU_TENSOR_MAP_SWIZZLE_128B, dtype) &&
      gt_make_map(&mw, w, N, K, K, gt::BK, gt::BN, CU_TENSOR_MAP_SWIZZLE_128B, dtype) &&
      (epilogue == 2 ? gt_make_map(&mo, out, M * (N / 32), 32, 32, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B, dtype)
                     : gt_make_map(&mo, out, M, N, N, gt::EPI_COLS, 32, CU_TENSOR_MAP_SWIZZLE_128B, dtype));
  if (!maps_ok) {
    set_error("linear_fused: cuTensorMapEncodeTiled failed");
    return PSALM_E_CUDA;
  }
  const long long tiles = (long long)p.tiles_m * p.tiles_n;
  const int grid = (int)(tiles < gt_sm_count() ? tiles : gt_sm_count());
  cudaError_t e;
  const bool bf = dtype == PSALM_BF16;
  switch (epilogue) {
    case 0: e = bf ? gt_launch<__nv_bfloat16, 0>(ma, mw, mo, p, grid, st) : gt_launch<__half, 0>(ma, mw, mo, p, grid, st); break;
    case 1: e = bf ? gt_launch<__nv_bfloat16, 1>(ma, mw, mo, p, grid, st) : gt_launch<__half, 1>(ma, mw, mo, p, grid, st); break;
    default: e = bf ? gt_launch<__nv_bfloat16, 2>(ma, mw, mo, p, grid, st) : gt_launch<__half, 2>(ma, mw, mo, p, grid, st); break;
  }
  if (e != cudaSuccess) {
    set_error("linear_fused: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    return PSALM_E_CUDA;
  }
  return check_launch("linear_fused");
}
