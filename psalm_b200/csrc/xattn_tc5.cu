// Masked cross-attention of the Mask2Former decoder on the 5th-generation tensor cores: tcgen05.mma with TMEM
// accumulators, K/V through TMA, softmax with one thread per query row.
//
// Replaces nn.MultiheadAttention with a float -inf mask [B*8, 100, HW] and materialised probabilities
// (transformer_decoder/mask2former_transformer_decoder.py:93-105, CrossAttentionLayer.forward_post), 100 queries x
// HW keys, 8 heads x 32, 16-bit storage.
//
// Why tcgen05 here: the warp-level mma.sync formulation (xattn_tma.cu, flash_mma_kernel<CrossMma>) spends ~14
// instructions per score element per thread - fragment shuffles for the row maxima, mask bit tests on a scattered
// fragment layout, accumulator rescaling, ldmatrix, HMMA issue, register moves under 128-255 registers of fragment
// state - and ran at 62 us (B = 4, HW = 16384; profiles/r2b_xattn_mma_ncu.txt: 25.2 M warp instructions at 0.43 IPC
// per scheduler).  With the accumulators in tensor memory a thread owns a whole query row: the mask word of 32 keys is
// ONE register tested bit by bit at compile-time positions, the row maximum and sum are thread-local (no shuffles, no
// votes), and the tensor work is issued by one thread.  What is left per element is the mask select, the exp2 and the
// running sum: the kernel is bound by the MUFU pipe (16 exp2 / clk / SM).
//
// CTA = 4 heads (two head pairs) of one key range of one image, 18 warps, 1 CTA per SM (all 512 TMEM columns):
//   warps 0-15  softmax: warpgroup h = head h, thread = query row (128 rows = the UMMA M; rows >= Lq are blocked)
//   warp 16     one elected lane issues every tcgen05.mma and tcgen05.commit
//   warp 17     one elected lane issues the TMA loads (K and V boxes [64 keys x 64 channels] = two heads, SWIZZLE_128B)
// Per 64-key tile and head:
//   S_h = Q_h K_h^T   2 x tcgen05.mma M128 N64 K16; A = the pre-scaled Q pair tile, B = the K box, both K-major: head
//                     (h & 1) is K-chunks 2(h & 1), 2(h & 1) + 1 of the 64-wide rows (descriptor start + 32 B per chunk)
//   softmax           tcgen05.ld of the S row, mask word -> selects, p = 2^(s - m) against a lazily raised reference
//                     m (raised when a score exceeds it by 2^8; then the O row is rescaled in TMEM), P written to
//                     shared memory as the K-major A operand of the next MMA
//   O_pair += P_h V   4 x tcgen05.mma M128 N64 K16 against the WHOLE V box (both heads of the pair, MN-major, the
//                     layout TMA delivers): half of that product is discarded - the tensor pipe has the room, and
//                     every descriptor stays a full SWIZZLE_128B atom; head h keeps columns 32 (h & 1) .. + 31
// The S MMAs of tile t+1 are issued as soon as the softmax threads have pulled S(t) into registers, so they run under
// the softmax of tile t.  Split-K over key ranges (one wave of CTAs); (m, l, O) partials are merged by a small kernel.
#include <cuda.h>

#include <type_traits>

#include "common.cuh"

namespace psalm {

namespace xb {
constexpr int NH = 8, HD = 32, C = 256, HPC = 4, KT = 64, NS = 2;
constexpr int TILE = 128 * 128;                 // bytes of a [128 rows x 64] 16-bit tile (Q pair, P)
constexpr int BOX = KT * 128;                   // bytes of a [64 keys x 64 ch] box
constexpr int HALF = 2 * BOX;                   // the K (or V) boxes of a tile: two head pairs = 16 KB
constexpr int SM_WARPS = HPC * 4;               // 16 softmax warps
constexpr int THREADS = (SM_WARPS + 2) * 32;    // + MMA warp + TMA warp
constexpr size_t SMEM = 1024 + (size_t)NS * 2 * HALF + 2 * TILE + 2 * HPC * TILE + 256;   // P double-buffered
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kRaise = 8.f;
}  // namespace xb

struct XbParams {
  const void* q;             // [B, Lq, 256]
  const uint32_t* bits;      // [B, Lq, W32] or null
  const uint8_t* row_open;   // [B, Lq] or null
  void* out;                 // [B, Lq, 256]
  float* part_o;             // [B, splits, 8, Lq, 32]
  float* part_ml;            // [B, splits, 8, Lq, 2]
  int B, Lq, Lk, W32, splits, tiles_per_split;
  float qscale;
};

__device__ __forceinline__ uint32_t xb_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t xb_swz(int r, int c) {   // 16-byte chunk c of row r in a [rows x 128 B] SWIZZLE_128B tile
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4));
}
// UMMA shared-memory descriptor, SWIZZLE_128B, 1024 B between 8-row groups (same encoding as attn_tc5.cu)
__device__ __forceinline__ uint64_t xb_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fff);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
template <typename T>
__device__ __forceinline__ uint32_t xb_idesc(int N, bool b_mn_major) {
  const uint32_t fmt = std::is_same<T, __nv_bfloat16>::value ? 1u : 0u;
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((b_mn_major ? 1u : 0u) << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void xb_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(xb_u32(bar)), "r"(count));
}
// a macro, not a function: profiler samples of a wait are then attributed to the call site (which barrier stalls)
#define xb_mbar_wait(bar, parity)                                                                       \
  do {                                                                                                  \
    uint32_t done_ = 0, spins_ = 0;                                                                     \
    while (!done_) {                                                                                    \
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"   \
                   "selp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done_) : "r"(xb_u32(bar)), "r"((uint32_t)(parity)) : "memory"); \
      if (++spins_ > (1u << 26)) __trap(); /* never hang the GPU on a protocol bug */                   \
    }                                                                                                   \
  } while (0)

__device__ __forceinline__ void xb_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(xb_u32(bar)) : "memory");
}
__device__ __forceinline__ void xb_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(xb_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void xb_commit(uint64_t* bar) {   // arrives on `bar` when every MMA issued so far has retired
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(xb_u32(bar)) : "memory");
}
__device__ __forceinline__ void xb_tma_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(
          xb_u32(dst)),
      "l"(map), "r"(c0), "r"(c1), "r"(xb_u32(bar))
      : "memory");
}
__device__ __forceinline__ void xb_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void xb_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,"
      "%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ void xb_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }
__device__ __forceinline__ void xb_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,"
      "%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};\n" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
      "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]),
      "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]),
      "r"(v[31])
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ float xb_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;\n" : "=f"(y) : "f"(x));
  return y;
}

// ------------------------------------------------------------------------------------------------------------
// grid = (splits, NH / HPC, B), block = 576, 1 CTA per SM
// ------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(xb::THREADS, 1)
xattn_tc5_kernel(const __grid_constant__ CUtensorMap mapK, const __grid_constant__ CUtensorMap mapV, XbParams p) {
  using namespace xb;
  extern __shared__ unsigned char xb_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(xb_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* sK = base;                                // [NS][K box 0, K box 1]   (K and V have separate rings: a K slot
  unsigned char* sV = base + (size_t)NS * HALF;            // [NS][V box 0, V box 1]    is free as soon as S retired, long before V)
  unsigned char* sQ = base + (size_t)NS * 2 * HALF;        // [2 head pairs][128 x 64]
  unsigned char* sP = sQ + 2 * TILE;                       // [2 buffers][4 heads][128 x 64 keys]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * HPC * TILE);
  uint64_t* k_full = bars;             // [NS]  TMA landed
  uint64_t* k_free = bars + NS;        // [NS]  the S MMAs reading the slot have retired
  uint64_t* v_full = bars + 2 * NS;    // [NS]
  uint64_t* v_free = bars + 3 * NS;    // [NS]  the PV MMAs reading the slot have retired
  uint64_t* s_full = bars + 4 * NS;    // [4]   S_h of the current tile is in TMEM
  uint64_t* s_free = s_full + HPC;     // [4]   the 128 rows of S_h are in registers
  uint64_t* p_full = s_free + HPC;     // [2][4] P_h (buffer t & 1) is in shared memory
  uint64_t* p_free = p_full + 2 * HPC; // [2][4] the PV MMA that read P_h (buffer t & 1) has retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(p_free + 2 * HPC);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int sp = blockIdx.x, hg = blockIdx.y, b = blockIdx.z;
  const int total_tiles = (p.Lk + KT - 1) / KT;
  const int base_t = total_tiles / p.splits, rem_t = total_tiles % p.splits;   // even distribution of the key tiles
  const int tile0 = sp * base_t + (sp < rem_t ? sp : rem_t);
  const int nt = base_t + (sp < rem_t ? 1 : 0);

  if (warp == SM_WARPS) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"(xb_u32(tmem_slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
  }
  if (tid == 0) {
    for (int s = 0; s < NS; ++s) {
      xb_mbar_init(&k_full[s], 1);
      xb_mbar_init(&k_free[s], 1);
      xb_mbar_init(&v_full[s], 1);
      xb_mbar_init(&v_free[s], 1);
    }
    for (int h = 0; h < HPC; ++h) {
      xb_mbar_init(&s_full[h], 1);
      xb_mbar_init(&s_free[h], 4);     // one arrival per softmax warp of the head
      for (int u = 0; u < 2; ++u) {
        xb_mbar_init(&p_full[u * HPC + h], 4);
        xb_mbar_init(&p_free[u * HPC + h], 1);
      }
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  // ---- Q of this head group -> two [128 x 64] SWIZZLE_128B pair tiles, pre-scaled by scale * log2e (rows >= Lq zero)
  if (warp < SM_WARPS) {
    const T* qb = reinterpret_cast<const T*>(p.q) + (size_t)b * p.Lq * C + hg * HPC * HD;
    for (int i = tid; i < 128 * 16; i += SM_WARPS * 32) {     // 16 chunks of 16 B per row (4 heads x 32 ch)
      const int row = i >> 4, c = i & 15;
      float f[8];
      if (row < p.Lq) {
        load16_as_f32<T>(qb + (size_t)row * C + c * 8, f);
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] *= p.qscale;
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = 0.f;
      }
      store16_from_f32<T>(reinterpret_cast<T*>(sQ + (c >> 3) * TILE + xb_swz(row, c & 7)), f);
    }
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // Q was written through the generic proxy, the MMA reads it through the async one
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::);
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
  const uint32_t tmem = *tmem_slot;

  if (warp == SM_WARPS + 1) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      for (int t = 0; t < nt; ++t) {
        const int st = t % NS;
        const int row0 = b * p.Lk + (tile0 + t) * KT;
        if (t >= NS) xb_mbar_wait(&k_free[st], ((t / NS) - 1) & 1);
        xb_mbar_expect_tx(&k_full[st], (uint32_t)HALF);
#pragma unroll
        for (int bx = 0; bx < 2; ++bx) xb_tma_2d(sK + (size_t)st * HALF + bx * BOX, &mapK, (hg * 2 + bx) * 64, row0, &k_full[st]);
        if (t >= NS) xb_mbar_wait(&v_free[st], ((t / NS) - 1) & 1);
        xb_mbar_expect_tx(&v_full[st], (uint32_t)HALF);
#pragma unroll
        for (int bx = 0; bx < 2; ++bx) xb_tma_2d(sV + (size_t)st * HALF + bx * BOX, &mapV, (hg * 2 + bx) * 64, row0, &v_full[st]);
      }
    }
  } else if (warp == SM_WARPS) {
    // =============================== MMA issuer ===============================
    if (lane == 0) {
      const uint32_t idesc_s = xb_idesc<T>(KT, false), idesc_o = xb_idesc<T>(64, true);
      for (int t = 0; t <= nt; ++t) {
        if (t < nt) {   // ---- S_h(t) = Q_h K_h(t)^T for the four heads
          const int st = t % NS;
          xb_mbar_wait(&k_full[st], (t / NS) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
          for (int h = 0; h < HPC; ++h) {
            if (t > 0) {
              xb_mbar_wait(&s_free[h], (t - 1) & 1);     // S_h(t-1) has been pulled into registers
              asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
            }
            const uint32_t a0 = xb_u32(sQ + (h >> 1) * TILE) + (h & 1) * 64;
            const uint32_t b0 = xb_u32(sK + (size_t)st * HALF + (h >> 1) * BOX) + (h & 1) * 64;
#pragma unroll
            for (int k = 0; k < 2; ++k) xb_mma(tmem + h * 64, xb_desc(a0 + k * 32), xb_desc(b0 + k * 32), idesc_s, k);
            xb_commit(&s_full[h]);
          }
          xb_commit(&k_free[st]);    // the K slot is free once these S MMAs have retired
        }
        if (t > 0) {    // ---- O_pair(h) += P_h(t-1) V(t-1)
          const int st = (t - 1) % NS;
          const int u = (t - 1) & 1;            // P buffer of tile t-1; its barriers complete once per two tiles
          xb_mbar_wait(&v_full[st], ((t - 1) / NS) & 1);
          for (int h = 0; h < HPC; ++h) {
            xb_mbar_wait(&p_full[u * HPC + h], ((t - 1) >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
            const uint32_t a0 = xb_u32(sP + (u * HPC + h) * TILE);
            const uint32_t b0 = xb_u32(sV + (size_t)st * HALF + (h >> 1) * BOX);
#pragma unroll
            for (int k = 0; k < KT / 16; ++k)
              xb_mma(tmem + 256 + h * 64, xb_desc(a0 + k * 32), xb_desc(b0 + k * 2048), idesc_o, (t > 1 || k > 0) ? 1u : 0u);
            xb_commit(&p_free[u * HPC + h]);
          }
          xb_commit(&v_free[st]);    // the V slot is free once these PV MMAs have retired
        }
      }
    }
  } else {
    // =============================== softmax: warpgroup = head, thread = query row ===============================
    const int hl = warp >> 2, row = (warp & 3) * 32 + lane;
    const int h = hg * HPC + hl;
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t tS = tmem + lane_off + hl * 64;
    const uint32_t tO = tmem + lane_off + 256 + hl * 64 + (hl & 1) * 32;   // this head's half of the pair product
    unsigned char* pbuf = sP + hl * TILE;
    const bool valid = row < p.Lq;
    const size_t grow = (size_t)b * p.Lq + (valid ? row : 0);
    const bool use_bits = valid && p.bits != nullptr && !(p.row_open && p.row_open[grow]);
    const uint32_t* brow = p.bits + grow * p.W32;
    float m_ref = -INFINITY, l_run = 0.f;
    // blocked bits of keys [32 kt32, 32 kt32 + 32): the raw word is only LOADED here (nothing touches it until the next
    // tile, so the load latency hides under a whole tile of math); `mask_fix` applies validity / tail at the use
    const int W32 = p.W32;
    auto mask_load = [&](int kt32) -> uint32_t { return (use_bits && kt32 < W32) ? __ldg(brow + kt32) : 0u; };
    auto mask_fix = [&](uint32_t raw, int kt32) -> uint32_t {
      const int left = p.Lk - kt32 * 32;
      if (!valid || left <= 0) return 0xffffffffu;
      return left < 32 ? (raw | (0xffffffffu << left)) : raw;
    };
    uint32_t w_next[2] = {0u, 0u};
    if (nt > 0) {
      w_next[0] = mask_load(2 * tile0);
      w_next[1] = mask_load(2 * tile0 + 1);
    }
    for (int t = 0; t < nt; ++t) {
      const uint32_t w[2] = {mask_fix(w_next[0], 2 * (tile0 + t)), mask_fix(w_next[1], 2 * (tile0 + t) + 1)};
      if (t + 1 < nt) {   // next tile's mask words travel under this tile's math
        w_next[0] = mask_load(2 * (tile0 + t + 1));
        w_next[1] = mask_load(2 * (tile0 + t + 1) + 1);
      }
      unsigned char* prow = pbuf + (t & 1) * HPC * TILE;
      xb_mbar_wait(&s_full[hl], t & 1);
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
      float ps[4] = {0.f, 0.f, 0.f, 0.f};
      // the tile is processed as two 32-key halves (32 live score registers instead of 64); the reference m is checked
      // per half
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t v[32];
        xb_ld32(tS + half * 32, v);
        xb_ld_wait();
        if (half == 1) {                                   // the S accumulator of this head may be overwritten
          asm volatile("tcgen05.fence::before_thread_sync;\n" ::);
          __syncwarp();
          if (lane == 0) xb_mbar_arrive(&s_free[hl]);
        }
        const uint32_t bm = w[half];
        // ---- blocked keys -> -inf (ONE select per element: the maximum and the exponential both see the masked score)
        float sm[32];
        float tm[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          sm[j] = ((bm >> j) & 1u) ? -INFINITY : __uint_as_float(v[j]);
          tm[j & 3] = fmaxf(tm[j & 3], sm[j]);
        }
        const float tmax = fmaxf(fmaxf(tm[0], tm[1]), fmaxf(tm[2], tm[3]));
        // ---- lazily raised reference
        float factor = 1.f;
        bool raise = false;
        if (m_ref == -INFINITY) {
          m_ref = tmax;                                     // nothing accumulated for this row yet (O row and l are 0)
        } else if (tmax > m_ref + kRaise) {
          factor = xb_exp2(m_ref - tmax);
          m_ref = tmax;
          raise = true;
        }
        const float msub = (m_ref == -INFINITY) ? 0.f : m_ref;
        // ---- this P buffer was last read by the PV MMA of tile t-2
        if (half == 0 && t > 1) {
          xb_mbar_wait(&p_free[(t & 1) * HPC + hl], ((t >> 1) - 1) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
        }
        if (__any_sync(0xffffffffu, raise)) {               // rescale the O rows of this warp in TMEM (rare after tile 0)
          if (t > 0) {                                      // O_h must be quiescent: the PV MMA of tile t-1 has retired
            xb_mbar_wait(&p_free[((t - 1) & 1) * HPC + hl], ((t - 1) >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
          }
          // half 1: the P chunks of half 0 were written against the old reference; fold the factor into them through
          // the row sum and the accumulator only - they are re-scaled below
          l_run = (l_run + (ps[0] + ps[1]) + (ps[2] + ps[3])) * factor;
          ps[0] = ps[1] = ps[2] = ps[3] = 0.f;
          uint32_t ov[32];
          xb_ld32(tO, ov);
          xb_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) ov[j] = __float_as_uint(__uint_as_float(ov[j]) * factor);
          xb_st32(tO, ov);
          if (half == 1) {                                  // re-scale the already written P chunks of half 0 (this row)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              uint4* pc = reinterpret_cast<uint4*>(prow + xb_swz(row, c));
              uint4 q4 = *pc;
              uint32_t* qw = reinterpret_cast<uint32_t*>(&q4);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                float lo, hi;
                unpack2<T>(qw[e], lo, hi);
                qw[e] = pack2<T>(lo * factor, hi * factor);
              }
              *pc = q4;
            }
          }
        }
        // ---- p = 2^(s - m) (blocked: 2^-inf = 0), packed to 16 bits into the K-major A operand of the PV MMA
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const float p0 = xb_exp2(sm[j] - msub), p1 = xb_exp2(sm[j + 1] - msub);
          ps[(j >> 1) & 3] += p0 + p1;
          pk[j >> 1] = pack2<T>(p0, p1);
        }
#pragma unroll
        for (int c = 0; c < 4; ++c)
          *reinterpret_cast<uint4*>(prow + xb_swz(row, half * 4 + c)) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
      }
      l_run += (ps[0] + ps[1]) + (ps[2] + ps[3]);
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;\n" ::);
      __syncwarp();
      if (lane == 0) xb_mbar_arrive(&p_full[(t & 1) * HPC + hl]);
    }
    // ---- epilogue: O row of this head out of TMEM
    if (nt > 0) {
      xb_mbar_wait(&p_free[((nt - 1) & 1) * HPC + hl], ((nt - 1) >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
    }
    uint32_t ov[32];
    if (nt > 0) {
      xb_ld32(tO, ov);
      xb_ld_wait();
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) ov[j] = 0u;
    }
    if (valid) {
      if (p.splits == 1) {
        const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
        T* dst = reinterpret_cast<T*>(p.out) + ((size_t)b * p.Lq + row) * C + h * HD;
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          uint4 o4;
          o4.x = pack2<T>(__uint_as_float(ov[j]) * inv, __uint_as_float(ov[j + 1]) * inv);
          o4.y = pack2<T>(__uint_as_float(ov[j + 2]) * inv, __uint_as_float(ov[j + 3]) * inv);
          o4.z = pack2<T>(__uint_as_float(ov[j + 4]) * inv, __uint_as_float(ov[j + 5]) * inv);
          o4.w = pack2<T>(__uint_as_float(ov[j + 6]) * inv, __uint_as_float(ov[j + 7]) * inv);
          *reinterpret_cast<uint4*>(dst + j) = o4;
        }
      } else {
        const size_t pr = (((size_t)b * p.splits + sp) * NH + h) * p.Lq + row;
        float4* po = reinterpret_cast<float4*>(p.part_o + pr * HD);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          po[j] = make_float4(__uint_as_float(ov[4 * j]), __uint_as_float(ov[4 * j + 1]), __uint_as_float(ov[4 * j + 2]),
                              __uint_as_float(ov[4 * j + 3]));
        *reinterpret_cast<float2*>(p.part_ml + pr * 2) = make_float2(m_ref, l_run);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::);
  __syncthreads();
  if (warp == SM_WARPS) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tmem));
}

// combine: one warp per (b, h, row), lane = channel; the (m, l) pairs of all splits are fetched with one coalesced
// load (lane = split) and broadcast by shuffle, the partial rows (128 B each) with independent loads
template <typename T>
__global__ void __launch_bounds__(256) xattn_tc5_combine_kernel(XbParams p) {
  using namespace xb;
  const int lane = threadIdx.x & 31;
  const int nw = p.B * NH * p.Lq;
  const int wi = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (wi >= nw) return;
  const int row = wi % p.Lq, h = (wi / p.Lq) % NH, b = wi / (p.Lq * NH);
  const size_t stride = (size_t)NH * p.Lq;                  // partial rows between consecutive splits
  const size_t pr0 = ((size_t)b * p.splits * NH + h) * p.Lq + row;
  float L = 0.f, O = 0.f, M = -INFINITY;
  for (int s0 = 0; s0 < p.splits; s0 += 32) {              // chunks of 32 splits (one per lane)
    const int n = p.splits - s0 < 32 ? p.splits - s0 : 32;
    float2 ml = make_float2(-INFINITY, 0.f);
    if (lane < n) ml = *reinterpret_cast<const float2*>(p.part_ml + (pr0 + (size_t)(s0 + lane) * stride) * 2);
    float cm = ml.x;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cm = fmaxf(cm, __shfl_xor_sync(0xffffffffu, cm, o));
    const float Mn = fmaxf(M, cm);
    if (Mn == -INFINITY) continue;
    const float resc = (M == -INFINITY) ? 0.f : xb_exp2(M - Mn);
    const float e_mine = (ml.x == -INFINITY) ? 0.f : xb_exp2(ml.x - Mn);
    float lsum = ml.y * e_mine;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
    L = L * resc + lsum;
    O *= resc;
    const float* po = p.part_o + (pr0 + (size_t)s0 * stride) * HD + lane;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    int s = 0;
    for (; s + 4 <= n; s += 4) {
      float v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = po[(size_t)(s + u) * stride * HD];
#pragma unroll
      for (int u = 0; u < 4; ++u) acc[u] = fmaf(v[u], __shfl_sync(0xffffffffu, e_mine, s + u), acc[u]);
    }
    for (; s < n; ++s) acc[0] = fmaf(po[(size_t)s * stride * HD], __shfl_sync(0xffffffffu, e_mine, s), acc[0]);
    O += (acc[0] + acc[1]) + (acc[2] + acc[3]);
    M = Mn;
  }
  const float val = L > 0.f ? O / L : 0.f;
  reinterpret_cast<T*>(p.out)[((size_t)b * p.Lq + row) * C + h * HD + lane] = from_f32<T>(val);
}

// ---- host ------------------------------------------------------------------------------------------------------
typedef CUresult (*XbEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static XbEncodeFn xb_encode_fn() {
  static XbEncodeFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<XbEncodeFn>(p);
  }
  return fn;
}
static bool xb_make_map(CUtensorMap* map, const void* base, long long rows, long long row_stride_elems, int dtype) {
  XbEncodeFn fn = xb_encode_fn();
  if (!fn) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)xb::C, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)row_stride_elems * 2};
  const cuuint32_t box[2] = {64, (cuuint32_t)xb::KT};
  const cuuint32_t estr[2] = {1, 1};
  return fn(map, dtype == PSALM_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
            const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
static int xb_sm_count() {
  static PerDevice cache;
  const int d = PerDevice::dev();
  if (cache.first() || cache.v[d] == 0) cudaDeviceGetAttribute(&cache.v[d], cudaDevAttrMultiProcessorCount, d);
  return cache.v[d] > 0 ? cache.v[d] : 148;
}
static int xb_splits(int B, int Lk) {
  const int tiles = (Lk + xb::KT - 1) / xb::KT;
  int s = xb_sm_count() / ((xb::NH / xb::HPC) * (B > 0 ? B : 1));   // one wave, 1 CTA per SM
  const int cap = (tiles + 1) / 2;                                   // at least 2 key tiles (128 keys) per CTA
  if (s > cap) s = cap;
  if (s < 1) s = 1;
  return s;
}

int tc5_cross_workspace_bytes(int B, int Lq, int Lk, size_t* bytes) {
  const int s = xb_splits(B, Lk);
  *bytes = s > 1 ? (size_t)B * s * xb::NH * Lq * (xb::HD + 2) * sizeof(float) : 0;
  return s;
}

int tc5_cross_attention(const void* q, const void* k, const void* v, long long kv_row_stride, const uint32_t* mask_bits,
                        const uint8_t* row_open, void* out, float* workspace, size_t workspace_bytes, int B, int Lq, int Lk,
                        int dtype, cudaStream_t st) {
  XbParams p;
  p.q = q; p.bits = mask_bits; p.row_open = row_open; p.out = out;
  p.B = B; p.Lq = Lq; p.Lk = Lk; p.W32 = (Lk + 31) / 32;
  size_t need = 0;
  p.splits = tc5_cross_workspace_bytes(B, Lq, Lk, &need);
  const int tiles = (Lk + xb::KT - 1) / xb::KT;
  p.tiles_per_split = (tiles + p.splits - 1) / p.splits;
  p.qscale = 1.0f / sqrtf((float)xb::HD) * xb::kLog2e;
  p.part_o = p.part_ml = nullptr;
  if (p.splits > 1) {
    PSALM_REQUIRE(workspace && workspace_bytes >= need, "masked_cross_attention(tcgen05): workspace of %zu bytes needed", need);
    p.part_o = workspace;
    p.part_ml = workspace + (size_t)B * p.splits * xb::NH * Lq * xb::HD;
  }
  CUtensorMap mk, mv;
  if (!xb_make_map(&mk, k, (long long)B * Lk, kv_row_stride, dtype) || !xb_make_map(&mv, v, (long long)B * Lk, kv_row_stride, dtype)) {
    set_error("masked_cross_attention(tcgen05): cuTensorMapEncodeTiled failed");
    return PSALM_E_CUDA;
  }
  dim3 grid(p.splits, xb::NH / xb::HPC, B);
  cudaError_t e;
  if (dtype == PSALM_BF16) {
    e = cudaFuncSetAttribute(xattn_tc5_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)xb::SMEM);
    if (e == cudaSuccess) xattn_tc5_kernel<__nv_bfloat16><<<grid, xb::THREADS, xb::SMEM, st>>>(mk, mv, p);
  } else {
    e = cudaFuncSetAttribute(xattn_tc5_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)xb::SMEM);
    if (e == cudaSuccess) xattn_tc5_kernel<__half><<<grid, xb::THREADS, xb::SMEM, st>>>(mk, mv, p);
  }
  if (e != cudaSuccess) {
    set_error("masked_cross_attention(tcgen05): cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    return PSALM_E_CUDA;
  }
  if (p.splits > 1) {
    const long long nw = (long long)B * xb::NH * Lq;
    const int blocks = (int)((nw + 7) / 8);
    if (dtype == PSALM_BF16) xattn_tc5_combine_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(p);
    else xattn_tc5_combine_kernel<__half><<<blocks, 256, 0, st>>>(p);
  }
  return check_launch("masked_cross_attention(tcgen05)");
}

}  // namespace psalm
