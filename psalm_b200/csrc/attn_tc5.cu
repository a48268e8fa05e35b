// Causal prefill attention of the LLM on the 5th-generation tensor cores (tcgen05 + TMEM), head_dim 64,
// 16-bit storage (third-party PhiAttention eager path; call site language_model/llava_phi.py:1354-1363).
//
// One CTA (128 threads) per (128-query tile, head, batch); thread r owns query row r = TMEM lane r, so the whole
// online softmax is thread-local (no shuffles):
//   S = Q K^T      tcgen05.mma M128 N128 K16 x 4, A = Q, B = K both K-major SWIZZLE_128B in shared memory
//                  (qkv rows are 128 B of head_dim: a 16-byte-chunk XOR swizzle of the plain cp.async image)
//   softmax        the S row (128 fp32) is pulled into registers with tcgen05.ld, which frees the S accumulator:
//                  the S MMA of tile t+1 is issued right away and runs under the softmax of tile t
//   O += P V       tcgen05.mma M128 N64 K16 x 8, A = P (written back to shared memory as bf16/fp16), B = V in its
//                  natural [key][head_dim] layout = MN-major SWIZZLE_128B (same shared-memory image as K, only the
//                  descriptor / b_major bit differ).  O stays in TMEM for the whole kernel: P is computed against
//                  a reference maximum that is only raised when the row maximum grows by more than 2^8 (exact:
//                  numerator and denominator share the reference), and only then is the O row rescaled in TMEM
//                  (tcgen05.ld / st) - after the first tile that is rare, so there is no per-tile O traffic.
// K is double-buffered with cp.async one full tile ahead, V is refilled as soon as the previous PV MMA has
// retired; the S and PV MMAs are issued from two different warps so that no single softmax warp carries both issue
// latencies; the second CTA on the SM covers what is left of the tensor-core / softmax bubbles.  fp32 accumulation.
#include <type_traits>

#include "common.cuh"

namespace psalm {

constexpr int FA_BQ = 128, FA_BK = 128, FA_HD = 64;
constexpr int FA_TILE = 128 * 128;   // bytes of a [128 rows x 64] 16-bit operand tile

__device__ __forceinline__ uint32_t fa_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// byte offset of 16-byte chunk c (0..7) of row r inside a [128 x 64] SWIZZLE_128B tile (8-row x 128 B atoms)
__device__ __forceinline__ uint32_t fa_swz(int r, int c) {
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4));
}

// UMMA shared-memory descriptor, SWIZZLE_128B, 1024 B between 8-row groups (cute::UMMA::SmemDescriptor bit layout).
// K-major operands: rows = M/N index; MN-major operand (V): rows = K index, 64 MN elements = one atom wide.
__device__ __forceinline__ uint64_t fa_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fff);          // start address, bits [0,14)
  d |= (uint64_t)1 << 16;                              // leading byte offset (single atom in that direction)
  d |= (uint64_t)(1024 >> 4) << 32;                    // stride byte offset, bits [32,46)
  d |= (uint64_t)1 << 46;                              // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                              // layout type SWIZZLE_128B
  return d;
}

// cute::UMMA::InstrDescriptor: c_format F32 (1) [4,6), a/b format [7,10)/[10,13), a_major bit 15, b_major bit 16
// (0 = K-major, 1 = MN-major), n_dim = N >> 3 at [17,23), m_dim = M >> 4 at [24,29)
template <typename T>
__device__ __forceinline__ uint32_t fa_idesc(int N, bool b_mn_major) {
  const uint32_t fmt = std::is_same<T, __nv_bfloat16>::value ? 1u : 0u;
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((b_mn_major ? 1u : 0u) << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(128 >> 4) << 24);
}

__device__ __forceinline__ void fa_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(fa_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fa_mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  long long spins = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done) : "r"(fa_smem_u32(bar)), "r"(parity) : "memory");
    if (++spins > (1ll << 28)) __trap();   // never hang the GPU on a protocol bug
  }
}

__device__ __forceinline__ void fa_tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
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

__device__ __forceinline__ void fa_tmem_ld32_nowait(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,"
      "%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ void fa_tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }

__device__ __forceinline__ void fa_tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,"
      "%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};\n"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
        "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]),
        "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]),
        "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
}

__device__ __forceinline__ float fa_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;\n" : "=f"(y) : "f"(x));
  return y;
}

// qkv [B,T,3,nh,64] (rotary applied) -> out [B,T,nh*64]; key_valid [B,T] uint8 or null
template <typename T>
__global__ void __launch_bounds__(128, 2) causal_tc5_kernel(const T* __restrict__ qkv, const uint8_t* __restrict__ key_valid,
                                                           T* __restrict__ out, int Tn, int nh, float scale_log2e) {
  extern __shared__ unsigned char fa_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(fa_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* sQ = base;                    // 16 KB
  unsigned char* sK = base + FA_TILE;          // 2 x 16 KB
  unsigned char* sV = base + 3 * FA_TILE;      // 16 KB (refilled as soon as the previous PV MMA has retired)
  unsigned char* sP = base + 4 * FA_TILE;      // 32 KB: two 64-key k-blocks of [128 x 64]
  __shared__ __align__(8) uint64_t bar_s, bar_o;
  __shared__ uint32_t tmem_base_smem;
  __shared__ uint32_t kvbits[16][4];           // invalid-key bits of up to 16 key tiles (T <= 2048)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nqt = (Tn + FA_BQ - 1) / FA_BQ;
  const int qt = nqt - 1 - (int)blockIdx.x;    // heavy (late) query tiles first
  const int h = blockIdx.y, b = blockIdx.z;
  const int q0 = qt * FA_BQ;
  const int ntiles = qt + 1;                   // causal, BQ == BK: key tiles 0..qt
  const size_t row_stride = (size_t)3 * nh * FA_HD;
  const T* qbase = qkv + (size_t)b * Tn * row_stride + (size_t)h * FA_HD;
  const T* kbase = qbase + (size_t)nh * FA_HD;
  const T* vbase = qbase + (size_t)2 * nh * FA_HD;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;\n" ::"r"(fa_smem_u32(&tmem_base_smem)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
  }
  if (tid == 0) {
    fa_mbar_init(&bar_s, 1);
    fa_mbar_init(&bar_o, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::);
  }
  // invalid-key bits: one warp per tile, 4 keys per lane
  for (int t = warp; t < ntiles; t += 4) {
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const int kj = t * FA_BK + w * 32 + lane;
      const bool bad = kj >= Tn || (key_valid && !key_valid[(size_t)b * Tn + kj]);
      const uint32_t m = __ballot_sync(0xffffffffu, bad);
      if (lane == 0) kvbits[t][w] = m;
    }
  }

  // rows [r0, r0 + 128) of a [T, 64] operand (row stride = row_stride) -> swizzled tile; rows >= T zero-filled.
  // The caller commits the cp.async group.
  auto load_tile = [&](const T* src, int r0, unsigned char* dst) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int i = tid + 128 * j;
      const int r = i >> 3, c = i & 7;
      const bool ok = r0 + r < Tn;
      const T* g = src + (size_t)(ok ? r0 + r : 0) * row_stride + c * 8;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(fa_smem_u32(dst + fa_swz(r, c))), "l"(g), "r"(ok ? 16 : 0));
    }
  };
  auto commit = [&]() { asm volatile("cp.async.commit_group;\n" ::); };
  // cp.async groups, in commit order: {Q, K0} {K1} {V0} | per tile t: {K(t+2)} {V(t)}.  Everything is drained
  // before PV(t) is issued, so at the S(t+1) issue only the two groups of tile t may still be in flight.
  load_tile(qbase, q0, sQ);
  load_tile(kbase, 0, sK);
  commit();
  if (ntiles > 1) load_tile(kbase, FA_BK, sK + FA_TILE);
  commit();
  load_tile(vbase, 0, sV);
  commit();

  asm volatile("tcgen05.fence::before_thread_sync;\n" ::);
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
  const uint32_t tmem_base = tmem_base_smem;
  const uint32_t tS = tmem_base, tO = tmem_base + 128;          // column offsets: S 128 columns, O 64 columns
  const uint32_t lane_off = (uint32_t)(warp * 32) << 16;        // this warp's TMEM lanes
  const uint32_t idesc_s = fa_idesc<T>(128, false), idesc_o = fa_idesc<T>(64, true);

  auto issue_s = [&](int t) {   // S = Q K(t)^T (single thread)
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
    const uint32_t a0 = fa_smem_u32(sQ), b0 = fa_smem_u32(sK + (t & 1) * FA_TILE);
#pragma unroll
    for (int k = 0; k < FA_HD / 16; ++k) {
      const uint64_t da = fa_desc(a0 + k * 32), db = fa_desc(b0 + k * 32);   // 16 elements = 32 B inside the atom
      const uint32_t acc = k > 0 ? 1u : 0u;
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
          ::"r"(tS), "l"(da), "l"(db), "r"(idesc_s), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(fa_smem_u32(&bar_s)) : "memory");
  };

  // Q and K(0) landed: S(0)
  asm volatile("cp.async.wait_group 2;\n" ::);
  asm volatile("fence.proxy.async.shared::cta;\n" ::);
  __syncthreads();
  if (tid == 32) issue_s(0);

  const int row = tid, qi = q0 + row;
  float m_ref = -INFINITY, l_run = 0.f;   // reference maximum (log2 domain) of P and of the row sum

  for (int t = 0; t < ntiles; ++t) {
    // ---- S(t) ready (so K(t) is consumed): refill its buffer with K(t+2), pull the S row into registers
    fa_mbar_wait(&bar_s, (uint32_t)(t & 1));
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
    if (t + 2 < ntiles) load_tile(kbase, (t + 2) * FA_BK, sK + (t & 1) * FA_TILE);
    commit();
    uint32_t v[128];
#pragma unroll
    for (int w = 0; w < 4; ++w) fa_tmem_ld32_nowait(tS + lane_off + (uint32_t)(w * 32), &v[w * 32]);
    // ---- PV(t-1) retired (it was issued after S(t), so this may spin briefly while the S row streams in):
    //      the V and P buffers are free and the O accumulator is stable
    if (t > 0) {
      fa_mbar_wait(&bar_o, (uint32_t)((t - 1) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
      load_tile(vbase, t * FA_BK, sV);
    }
    commit();
    fa_tmem_ld_wait();
    // ---- the S accumulator is free again once every warp has its rows: issue S(t+1) under this softmax
    asm volatile("cp.async.wait_group 2;\n" ::);          // K(t+1) landed
    asm volatile("fence.proxy.async.shared::cta;\n" ::);
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::);
    __syncthreads();
    if (tid == 32 && t + 1 < ntiles) issue_s(t + 1);   // warp 1 pays the S issue latency, warp 2 the PV one

    // ---- softmax of row `row` over the 128 keys of this tile (scores in the log2 domain)
    const bool diag = (t == qt);
    uint32_t blk[4] = {kvbits[t][0], kvbits[t][1], kvbits[t][2], kvbits[t][3]};
    const bool masked = diag || ((blk[0] | blk[1] | blk[2] | blk[3]) != 0u);
    if (diag) {
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        const int d = row - w * 32;   // keys with (w*32 + j) > row are in the future
        blk[w] |= d >= 31 ? 0u : (d < 0 ? 0xffffffffu : (0xffffffffu << (d + 1)));
      }
    }
    float tm[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};   // four independent chains
    if (!masked) {
#pragma unroll
      for (int j = 0; j < 128; ++j) tm[j & 3] = fmaxf(tm[j & 3], __uint_as_float(v[j]));
    } else {
#pragma unroll
      for (int j = 0; j < 128; ++j)
        tm[j & 3] = ((blk[j >> 5] >> (j & 31)) & 1u) ? tm[j & 3] : fmaxf(tm[j & 3], __uint_as_float(v[j]));
    }
    const float tmax = fmaxf(fmaxf(tm[0], tm[1]), fmaxf(tm[2], tm[3])) * scale_log2e;   // scale > 0: max commutes
    // raise the reference only when the maximum grew by more than 2^8 (P <= 256 otherwise: harmless)
    float factor = 1.f;
    bool raise = false;
    if (m_ref == -INFINITY) {
      m_ref = tmax;                                        // nothing accumulated for this row yet (O row and l are 0)
    } else if (tmax > m_ref + 8.f) {
      factor = fa_exp2(m_ref - tmax);
      m_ref = tmax;
      raise = true;
    }
    const float msub = (m_ref == -INFINITY) ? 0.f : m_ref;

    if (__any_sync(0xffffffffu, raise)) {                  // rescale the O rows of this warp in TMEM
      l_run *= factor;
#pragma unroll
      for (int w = 0; w < 2; ++w) {
        uint32_t ov[32];
        fa_tmem_ld32(tO + lane_off + (uint32_t)(w * 32), ov);
#pragma unroll
        for (int j = 0; j < 32; ++j) ov[j] = __float_as_uint(__uint_as_float(ov[j]) * factor);
        fa_tmem_st32(tO + lane_off + (uint32_t)(w * 32), ov);
      }
    }
    float ps[4] = {0.f, 0.f, 0.f, 0.f};   // four independent chains
    // p = 2^(s * scale - m_ref), packed to 16 bits and written as the A operand of the PV MMA: 32 keys = 4 chunks
    // of 16 B in key block (w >> 1), chunks (w & 1) * 4 .. + 3 of this row.  Two real code paths (the mask
    // tests would otherwise be executed as selects on every tile: +25 % instructions).
    auto emit = [&](auto masked_tag) {
      constexpr bool kMasked = decltype(masked_tag)::value;
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        const uint32_t bm = kMasked ? blk[w] : 0u;
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          float p0 = fa_exp2(fmaf(__uint_as_float(v[w * 32 + j]), scale_log2e, -msub));
          float p1 = fa_exp2(fmaf(__uint_as_float(v[w * 32 + j + 1]), scale_log2e, -msub));
          if (kMasked) {
            p0 = ((bm >> j) & 1u) ? 0.f : p0;
            p1 = ((bm >> (j + 1)) & 1u) ? 0.f : p1;
          }
          ps[(j >> 1) & 3] += p0 + p1;
          pk[j >> 1] = pack2<T>(p0, p1);
        }
        unsigned char* prow = sP + (w >> 1) * FA_TILE;
#pragma unroll
        for (int c = 0; c < 4; ++c)
          *reinterpret_cast<uint4*>(prow + fa_swz(row, (w & 1) * 4 + c)) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
      }
    };
    if (masked) emit(std::true_type{});
    else emit(std::false_type{});
    l_run += (ps[0] + ps[1]) + (ps[2] + ps[3]);

    // ---- V(t) landed, P written: O += P V
    asm volatile("cp.async.wait_group 0;\n" ::);
    asm volatile("fence.proxy.async.shared::cta;\n" ::);
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::);
    __syncthreads();
    if (tid == 64) {
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
      const uint32_t a0 = fa_smem_u32(sP), b0 = fa_smem_u32(sV);
#pragma unroll
      for (int k = 0; k < FA_BK / 16; ++k) {
        const uint64_t da = fa_desc(a0 + (k >> 2) * FA_TILE + (k & 3) * 32);   // P: K-major, two 64-key blocks
        const uint64_t db = fa_desc(b0 + k * 2048);                            // V: MN-major, 16 keys = two 8-row atoms
        const uint32_t acc = (t > 0 || k > 0) ? 1u : 0u;
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
            ::"r"(tO), "l"(da), "l"(db), "r"(idesc_o), "r"(acc) : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(fa_smem_u32(&bar_o)) : "memory");
    }
  }
  asm volatile("cp.async.wait_group 0;\n" ::);
  fa_mbar_wait(&bar_o, (uint32_t)((ntiles - 1) & 1));
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::);

  {
    const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
    T* dst = out + ((size_t)b * Tn + (qi < Tn ? qi : 0)) * nh * FA_HD + (size_t)h * FA_HD;
#pragma unroll
    for (int w = 0; w < 2; ++w) {
      uint32_t ov[32];
      fa_tmem_ld32(tO + lane_off + (uint32_t)(w * 32), ov);   // warp-collective: every lane takes part
      if (qi < Tn) {
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          uint4 o4;
          o4.x = pack2<T>(__uint_as_float(ov[j]) * inv, __uint_as_float(ov[j + 1]) * inv);
          o4.y = pack2<T>(__uint_as_float(ov[j + 2]) * inv, __uint_as_float(ov[j + 3]) * inv);
          o4.z = pack2<T>(__uint_as_float(ov[j + 4]) * inv, __uint_as_float(ov[j + 5]) * inv);
          o4.w = pack2<T>(__uint_as_float(ov[j + 6]) * inv, __uint_as_float(ov[j + 7]) * inv);
          *reinterpret_cast<uint4*>(dst + w * 32 + j) = o4;
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::);
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;\n" ::"r"(tmem_base));
}

bool tc5_causal_ok(int B, int T_, int nh, int hd, int dtype) {
  return hd == FA_HD && (dtype == PSALM_BF16 || dtype == PSALM_F16) && T_ >= 1 && T_ <= 16 * FA_BK && nh <= 65535 && B <= 65535;
}

int tc5_causal_attention(const void* qkv, const uint8_t* key_valid, void* out, int B, int T_, int nh, int hd, int dtype,
                         cudaStream_t st) {
  if (!tc5_causal_ok(B, T_, nh, hd, dtype)) {
    set_error("causal_attention(tcgen05): needs head_dim 64, 16-bit storage, T <= %d", 16 * FA_BK);
    return PSALM_E_UNSUPPORTED;
  }
  const size_t smem = 6 * (size_t)FA_TILE + 1024;
  const float sc = (1.0f / sqrtf((float)hd)) * 1.4426950408889634f;
  dim3 grid((T_ + FA_BQ - 1) / FA_BQ, nh, B);
  cudaError_t e;
  if (dtype == PSALM_BF16) {
    using T = __nv_bfloat16;
    e = cudaFuncSetAttribute(causal_tc5_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) causal_tc5_kernel<T><<<grid, 128, smem, st>>>((const T*)qkv, key_valid, (T*)out, T_, nh, sc);
  } else {
    using T = __half;
    e = cudaFuncSetAttribute(causal_tc5_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) causal_tc5_kernel<T><<<grid, 128, smem, st>>>((const T*)qkv, key_valid, (T*)out, T_, nh, sc);
  }
  if (e != cudaSuccess) {
    set_error("causal_attention(tcgen05): cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    return PSALM_E_CUDA;
  }
  return check_launch("causal_tc5_kernel");
}

}  // namespace psalm
