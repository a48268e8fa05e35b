// Final mask projection on the 5th-generation tensor cores (tcgen05 + TMEM):
//   out[q, p] = sum_c mask_embed[q, c] * feats[p, c]      (einsum "bqc,bchw->bqhw",
//   mask2former_transformer_decoder.py:750; Q <= 128, C = 256, P = H4*W4 = 65536 at 1024^2)
//
// Both operands are K-major in HBM (token-major feature map), which is exactly the canonical UMMA
// K-major SWIZZLE_128B shared-memory layout after a 16-byte-chunk XOR swizzle, so tiles are staged with
// plain 16-byte cp.async copies (no transposition).  One persistent CTA per SM:
//   * A = mask_embed padded to 128 rows, resident in shared memory for the whole kernel (64 KB);
//   * B = 128-pixel feature tiles (64 KB), double buffered;
//   * D = 128 x 128 fp32 accumulators in TMEM, double buffered (2 x 128 columns): a single elected
//     thread issues 16 tcgen05.mma (M = 128, N = 128, K = 16) per tile and commits to an mbarrier;
//   * the four warps drain TMEM with tcgen05.ld (32 lanes x 32 columns per instruction), convert and
//     store tile t-1 while the tensor core works on tile t and cp.async fetches tile t+1.
// fp32 accumulation; 16-bit storage types only (fp32 storage keeps the exact SIMT kernel).
#include <type_traits>

#include "common.cuh"

namespace psalm {

constexpr int TC_M = 128, TC_N = 128, TC_K = 256, TC_UK = 16;
constexpr int TC_TILE_BYTES = TC_M * TC_K * 2;   // 64 KB per operand tile (A and one B stage)

__device__ __forceinline__ uint32_t tc_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// byte offset of element (row r, k) inside a [128 x 256] K-major SWIZZLE_128B operand tile
__device__ __forceinline__ uint32_t tc_swz_offset(int r, int k8 /* 16-byte chunk index 0..31 */) {
  const int kb = k8 >> 3, c = k8 & 7;          // 64-element k-block, chunk inside the 128-byte row
  return (uint32_t)(kb * (TC_M * 128) + (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4));
}

// UMMA shared-memory descriptor, K-major, SWIZZLE_128B (cute::UMMA::SmemDescriptor bit layout)
__device__ __forceinline__ uint64_t tc_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fff);          // start address, bits [0,14)
  d |= (uint64_t)1 << 16;                              // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                    // stride byte offset: 8 rows x 128 B, bits [32,46)
  d |= (uint64_t)1 << 46;                              // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                              // layout type SWIZZLE_128B
  return d;
}

template <typename T>
__device__ __forceinline__ uint32_t tc_idesc() {
  // cute::UMMA::InstrDescriptor: c_format F32 (1) [4,6), a/b format [7,10)/[10,13), K-major A and B,
  // n_dim = N >> 3 at [17,23), m_dim = M >> 4 at [24,29)
  const uint32_t fmt = std::is_same<T, __nv_bfloat16>::value ? 1u : 0u;
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
}

__device__ __forceinline__ void tc_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(tc_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void tc_mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  long long spins = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done) : "r"(tc_smem_u32(bar)), "r"(parity) : "memory");
    if (++spins > (1ll << 28)) __trap();   // never hang the GPU on a protocol bug
  }
}

template <typename T>
__global__ void __launch_bounds__(128, 1) mask_proj_tc5_kernel(const T* __restrict__ me, const T* __restrict__ feats,
                                                               T* __restrict__ out, int Q, int P, int tiles_per_batch,
                                                               int total_tiles) {
  extern __shared__ unsigned char tc_raw[];
  // 1024-byte alignment required by the 128-byte swizzle atoms
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(tc_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* sA = base;                        // 64 KB, (re)loaded when the batch index changes
  unsigned char* sB = base + TC_TILE_BYTES;        // 2 x 64 KB
  __shared__ __align__(8) uint64_t bar_mma[2];
  __shared__ uint32_t tmem_base_smem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;\n" ::"r"(tc_smem_u32(&tmem_base_smem)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
  }
  if (tid == 0) {
    tc_mbar_init(&bar_mma[0], 1);
    tc_mbar_init(&bar_mma[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::);
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
  const uint32_t tmem_base = tmem_base_smem;
  const uint32_t idesc = tc_idesc<T>();

  auto load_tile = [&](const T* src, int rows_valid, unsigned char* dst) {
    // 128 rows x 32 chunks of 16 bytes, rows >= rows_valid are zero filled
    for (int i = tid; i < TC_M * 32; i += 128) {
      const int r = i >> 5, k8 = i & 31;
      const bool ok = r < rows_valid;
      const uint32_t d = tc_smem_u32(dst + tc_swz_offset(r, k8));
      const T* g = src + (size_t)(ok ? r : 0) * TC_K + k8 * 8;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(g), "r"(ok ? 16 : 0));
    }
    asm volatile("cp.async.commit_group;\n" ::);
  };

  // drain TMEM accumulator stage `ps` of tile `tile` to global memory (row = query = TMEM lane)
  auto epilogue = [&](int tile, int ps) {
    const int pb = tile / tiles_per_batch, pp0 = (tile % tiles_per_batch) * TC_N;
    const int q = warp * 32 + lane;
#pragma unroll 1
    for (int c0 = 0; c0 < TC_N; c0 += 32) {
      uint32_t v[32];
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(ps * TC_N + c0);
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,"
          "%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
            "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
            "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
            "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
      if (q < Q) {
        T* dst = out + ((size_t)pb * Q + q) * P + pp0 + c0;
        if (pp0 + c0 + 32 <= P && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            uint4 w;
            w.x = pack2<T>(__uint_as_float(v[j]), __uint_as_float(v[j + 1]));
            w.y = pack2<T>(__uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
            w.z = pack2<T>(__uint_as_float(v[j + 4]), __uint_as_float(v[j + 5]));
            w.w = pack2<T>(__uint_as_float(v[j + 6]), __uint_as_float(v[j + 7]));
            *reinterpret_cast<uint4*>(dst + j) = w;
          }
        } else {
          for (int j = 0; j < 32; ++j)
            if (pp0 + c0 + j < P) dst[j] = from_f32<T>(__uint_as_float(v[j]));
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::);
  };

  int cur_batch = -1;
  int it = 0;
  int prev_tile = -1;
  // iteration `it` (tile t, stage s): wait MMA(it-1) | load(t+1) | MMA(t) | epilogue(t-1)
  const int first = blockIdx.x, step = gridDim.x;
  if (first < total_tiles) {
    const int b = first / tiles_per_batch, p0 = (first % tiles_per_batch) * TC_N;
    load_tile(feats + ((size_t)b * P + p0) * TC_K, min(TC_N, P - p0), sB);
  }
  for (int t = first; t < total_tiles; t += step, ++it) {
    const int s = it & 1;
    const int b = t / tiles_per_batch;
    if (prev_tile >= 0) {
      // MMA(it-1) finished: its B stage (s^1) and the A tile may be overwritten, its accumulators are ready
      tc_mbar_wait(&bar_mma[s ^ 1], (uint32_t)(((it - 1) >> 1) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
    }
    if (b != cur_batch) {   // A operand of this batch element (once per CTA at batch 1)
      load_tile(me + (size_t)b * Q * TC_K, Q, sA);
      cur_batch = b;
    }
    const int tn = t + step;
    if (tn < total_tiles) {
      const int bn = tn / tiles_per_batch, pn = (tn % tiles_per_batch) * TC_N;
      load_tile(feats + ((size_t)bn * P + pn) * TC_K, min(TC_N, P - pn), sB + (s ^ 1) * TC_TILE_BYTES);
      asm volatile("cp.async.wait_group 1;\n" ::);
    } else {
      asm volatile("cp.async.wait_group 0;\n" ::);
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::);   // cp.async (generic proxy) -> tensor core (async proxy)
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
      const uint32_t a0 = tc_smem_u32(sA), b0 = tc_smem_u32(sB + s * TC_TILE_BYTES);
      const uint32_t d_tmem = tmem_base + (uint32_t)(s * TC_N);
#pragma unroll
      for (int k = 0; k < TC_K / TC_UK; ++k) {
        // k-th 16-element slice: k-block (k >> 2) of 64 elements, 32 bytes per slice inside the swizzle atom
        const uint32_t off = (uint32_t)((k >> 2) * (TC_M * 128) + (k & 3) * 32);
        const uint64_t da = tc_desc(a0 + off), db = tc_desc(b0 + off);
        const uint32_t acc = k > 0 ? 1u : 0u;
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
            ::"r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(tc_smem_u32(&bar_mma[s])) : "memory");
    }
    if (prev_tile >= 0) epilogue(prev_tile, s ^ 1);   // overlaps with the MMAs of tile t
    prev_tile = t;
    __syncthreads();
  }
  if (prev_tile >= 0) {   // drain the last tile
    const int ps = (it - 1) & 1;
    tc_mbar_wait(&bar_mma[ps], (uint32_t)(((it - 1) >> 1) & 1));
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
    epilogue(prev_tile, ps);
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::);
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;\n" ::"r"(tmem_base));
}

int tc5_mask_proj(const void* me, const void* feats, void* out, int B, int Q, int P, int dtype, cudaStream_t st) {
  const int tiles_per_batch = (P + TC_N - 1) / TC_N;
  const int total = tiles_per_batch * B;
  const size_t smem = 3 * (size_t)TC_TILE_BYTES + 1024;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = total < sms ? total : sms;
  cudaError_t e;
  if (dtype == PSALM_BF16) {
    using T = __nv_bfloat16;
    e = cudaFuncSetAttribute(mask_proj_tc5_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("mask_proj_tc5: %s", cudaGetErrorString(e)); return PSALM_E_CUDA; }
    mask_proj_tc5_kernel<T><<<grid, 128, smem, st>>>((const T*)me, (const T*)feats, (T*)out, Q, P, tiles_per_batch, total);
  } else {
    using T = __half;
    e = cudaFuncSetAttribute(mask_proj_tc5_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("mask_proj_tc5: %s", cudaGetErrorString(e)); return PSALM_E_CUDA; }
    mask_proj_tc5_kernel<T><<<grid, 128, smem, st>>>((const T*)me, (const T*)feats, (T*)out, Q, P, tiles_per_batch, total);
  }
  return check_launch("mask_proj_tc5_kernel");
}

}  // namespace psalm
