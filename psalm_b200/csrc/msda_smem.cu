// Multi-scale deformable attention for the pixel-decoder encoder: value tiles staged in shared memory by TMA, the
// bilinear gather + weighted sum of a (query, head) as a tensor-core contraction over gathered rows.
//
// Replaces ms_deformable_im2col_gpu_kernel + softmax + location arithmetic (reference
// ops/src/cuda/ms_deform_im2col_cuda.cuh:243-304, ops/modules/ms_deform_attn.py:103-110) like
// msda_encoder_fused_kernel (msda.cu), whose limit is the L1 gather path: 8.26 M corner fetches of 64 B per
// layer-image, one L1 wavefront each, plus ~30 instructions per fetched 16 bytes (two-byte -> fp32 conversions, FMAs,
// shuffles) - 45 us at 1024^2 (profiles/r1k_msda_fused_bf16_ncu_details.txt: L1/TEX 72 %, issue slots 65 %).  Here:
//
//   * CTA = (16 x 16 cell of the finest level, head, image).  The queries of ALL levels whose reference point lies in
//     the cell (256 + 64 + 16 at 1024^2) sample the same neighbourhood, so one (cell + halo) tile per level is staged
//     ONCE: three cp.async.bulk.tensor.4d loads (TMA) of boxes [th x tw x 32 channels] with SWIZZLE_64B; coordinates
//     outside the map are zero-filled by the TMA unit = the op's zero padding, no predicates in the hot loop.
//   * out[q, h, :] = sum over 48 corners of w_c * V[corner_c, :] is, per k-step of 16 corners (= the 4 samples of one
//     level), ONE ldmatrix.x4.trans whose 32 lanes supply the 32 gathered row addresses (16 corners x 2 channel blocks)
//     and two mma.sync.m16n8k16: A = the gathered value rows (exact 16-bit data), B = the corner weights in columns
//     0 / 1 as a hi + lo pair of 16-bit floats (w = hi + lo to 2^-17), fp32 accumulation.  6 HMMA + 6 LDSM per (query,
//     head) instead of ~400 lane instructions: no conversions, no FMAs, no per-sample shuffles.
//   * sample parameters (softmax over the 12 logits, reference point, location, bilinear weights, tile index) are
//     computed once per (query, head) by 12 lanes of a half-warp and handed to the contraction through 1 KB of
//     per-warp shared memory.
//   * samples that leave the staged tile (offset beyond the halo) take a predicated global-memory path for that sample
//     only (its MMA weights are zeroed): results do not depend on the halo.
// 16-bit storage, M = 8, D = 32, L = 3, P = 4, levels ordered coarse -> fine (the pixel decoder's res5, res4, res3).
#include <cuda.h>

#include <type_traits>

#include "common.cuh"

namespace psalm {

namespace ms {
constexpr int M = 8, D = 32, L = 3, P = 4, LP = 12, CELL = 16, WARPS = 16, THREADS = WARPS * 32;
constexpr int OW_ROW = M * LP * 3;   // 288 values per query: offsets (m, l, p, xy) then logits (m, l, p)
}

struct MsParams {
  const void* value;   // [B, M, S, D] (for the out-of-tile path)
  const void* ow;      // [B, S, 288]
  void* out;           // [B, S, 256]
  int B, S;
  int H[3], W[3], start[3];
  int tw[3], th[3], toff[3];   // tile (TMA box) width / height in pixels, byte offset in shared memory
  int halo, cells_x, cells_y;
  int stage_off;               // byte offset of the per-warp staging area
};

__device__ __forceinline__ int ms_floordiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }
__device__ __forceinline__ int ms_ceildiv(int a, int b) { return (a >= 0) ? (a + b - 1) / b : -((-a) / b); }
__device__ __forceinline__ uint32_t ms_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <typename T>
__device__ __forceinline__ void ms_mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  if constexpr (std::is_same<T, __nv_bfloat16>::value) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  } else {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
}

// w -> (hi, lo) with hi + lo == w to 16+ mantissa bits
template <typename T>
__device__ __forceinline__ void ms_split(float w, float& hi, float& lo) {
  hi = to_f32<T>(from_f32<T>(w));
  lo = w - hi;
}

// value of level l out of three scalars (no dynamically indexed arrays: those would live in local memory, and with
// 2 x 109 KB of shared memory per SM the L1 that backs local memory is ~10 KB: 86 % of such loads missed)
template <typename V>
__device__ __forceinline__ V ms_pick(int l, V a, V b, V c) { return l == 0 ? a : (l == 1 ? b : c); }

// first query index of level `l` along one axis whose reference point lies in cell >= c:
// smallest q with (2q + 1) * Wf >= 2 * CELL * c * Wl
__device__ __forceinline__ int ms_q_lo(int c, int Wl, int Wf) {
  int q = ms_ceildiv(2 * ms::CELL * c * Wl - Wf, 2 * Wf);
  q = q < 0 ? 0 : q;
  return q > Wl ? Wl : q;
}

// grid = (cells, M, B), block = 256
template <typename T, typename TO>
__global__ void __launch_bounds__(ms::THREADS, 2)
msda_smem_kernel(const __grid_constant__ CUtensorMap map0, const __grid_constant__ CUtensorMap map1,
                 const __grid_constant__ CUtensorMap map2, MsParams p) {
  using namespace ms;
  extern __shared__ unsigned char ms_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(ms_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cell = blockIdx.x, m = blockIdx.y, b = blockIdx.z;
  const int cy = cell / p.cells_x, cx = cell - cy * p.cells_x;
  const int Wf = p.W[L - 1], Hf = p.H[L - 1];

  // ---- per-level constants in scalars (the level loops below are fully unrolled)
  const int W0 = p.W[0], W1 = p.W[1], W2 = p.W[2], H0 = p.H[0], H1 = p.H[1], H2 = p.H[2];
  const int tw0 = p.tw[0], tw1 = p.tw[1], tw2 = p.tw[2], th0 = p.th[0], th1 = p.th[1], th2 = p.th[2];
  // tile origins (level pixel coordinates, may be negative: zero fill)
  const int tx00 = ms_floordiv(2 * CELL * cx * W0 - Wf, 2 * Wf) - p.halo, ty00 = ms_floordiv(2 * CELL * cy * H0 - Hf, 2 * Hf) - p.halo;
  const int tx01 = ms_floordiv(2 * CELL * cx * W1 - Wf, 2 * Wf) - p.halo, ty01 = ms_floordiv(2 * CELL * cy * H1 - Hf, 2 * Hf) - p.halo;
  const int tx02 = ms_floordiv(2 * CELL * cx * W2 - Wf, 2 * Wf) - p.halo, ty02 = ms_floordiv(2 * CELL * cy * H2 - Hf, 2 * Hf) - p.halo;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(ms_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    const uint32_t bytes = (uint32_t)((tw0 * th0 + tw1 * th1 + tw2 * th2) * D * (int)sizeof(T));
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(ms_u32(&bar)), "r"(bytes) : "memory");
#define MS_TMA(MAP, OFF, X, Y)                                                                                            \
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];\n" \
               ::"r"(ms_u32(smem + (OFF))), "l"(&(MAP)), "r"(0), "r"(X), "r"(Y), "r"(b * M + m), "r"(ms_u32(&bar)) : "memory")
    MS_TMA(map0, p.toff[0], tx00, ty00);
    MS_TMA(map1, p.toff[1], tx01, ty01);
    MS_TMA(map2, p.toff[2], tx02, ty02);
#undef MS_TMA
  }
  // ---- the queries of this cell: per level a rectangle [qx0, qx0 + nx) x [qy0, ...), units numbered level by level
  const int qx00 = ms_q_lo(cx, W0, Wf), qx01 = ms_q_lo(cx, W1, Wf), qx02 = ms_q_lo(cx, W2, Wf);
  const int qy00 = ms_q_lo(cy, H0, Hf), qy01 = ms_q_lo(cy, H1, Hf), qy02 = ms_q_lo(cy, H2, Hf);
  const int nx0 = ms_q_lo(cx + 1, W0, Wf) - qx00, nx1 = ms_q_lo(cx + 1, W1, Wf) - qx01, nx2 = ms_q_lo(cx + 1, W2, Wf) - qx02;
  const int cnt1 = nx0 * (ms_q_lo(cy + 1, H0, Hf) - qy00);
  const int cnt2 = cnt1 + nx1 * (ms_q_lo(cy + 1, H1, Hf) - qy01);
  const int units = cnt2 + nx2 * (ms_q_lo(cy + 1, H2, Hf) - qy02);
  // row = ul / nx by multiply-shift (ul < 4096, nx <= 64: exact)
  const int inv0 = (65536 + (nx0 > 0 ? nx0 : 1) - 1) / (nx0 > 0 ? nx0 : 1);
  const int inv1 = (65536 + (nx1 > 0 ? nx1 : 1) - 1) / (nx1 > 0 ? nx1 : 1);
  const int inv2 = (65536 + (nx2 > 0 ? nx2 : 1) - 1) / (nx2 > 0 ? nx2 : 1);
  const float rW0 = 1.f / (float)W0, rW1 = 1.f / (float)W1, rW2 = 1.f / (float)W2;
  const float rH0 = 1.f / (float)H0, rH1 = 1.f / (float)H1, rH2 = 1.f / (float)H2;
  // per-warp staging: weights [2 units][12 samples] uint4 (hi01, hi23, lo01, lo23), tile pixel index [2][12] (+ 4 pad),
  // output row [2][32] fp32
  unsigned char* stg = smem + p.stage_off + warp * 1024;
  uint4* stg_w = reinterpret_cast<uint4*>(stg);                    // 384 B
  int* stg_p = reinterpret_cast<int*>(stg + 384);                  // 128 B
  float* stg_o = reinterpret_cast<float*>(stg + 512);              // 256 B
  const T* vplane = reinterpret_cast<const T*>(p.value) + ((size_t)b * M + m) * (size_t)p.S * D;

  __syncthreads();   // mbarrier initialised before anybody polls it
  {
    uint32_t done = 0, spins = 0;
    while (!done) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                   : "=r"(done) : "r"(ms_u32(&bar)) : "memory");
      if (++spins > (1u << 26)) __trap();
    }
  }
  const uint32_t tb0 = ms_u32(smem + p.toff[0]), tb1 = ms_u32(smem + p.toff[1]), tb2 = ms_u32(smem + p.toff[2]);
  const int half = lane >> 4, hl = lane & 15;
  const int g = lane >> 2, t4 = lane & 3;
  const int st0 = p.start[0], st1 = p.start[1], st2 = p.start[2];

  // unit index -> (level, qy, qx)
  auto unit_query = [&](int u, int& ql, int& qxx, int& qyy) {
    ql = (u >= cnt1) + (u >= cnt2);
    const int ul = u - ms_pick(ql, 0, cnt1, cnt2);
    const int r = (ul * ms_pick(ql, inv0, inv1, inv2)) >> 16;
    qyy = ms_pick(ql, qy00, qy01, qy02) + r;
    qxx = ms_pick(ql, qx00, qx01, qx02) + (ul - r * ms_pick(ql, nx0, nx1, nx2));
  };
  // raw Linear outputs (offset x, offset y, logit) of sample hl of unit u0 + half; fetched one iteration ahead so that
  // their latency hides under the previous pair of units
  auto fetch = [&](int u0, float& ox, float& oy, float& lg) {
    const int u = u0 + half;
    ox = oy = 0.f;
    lg = -INFINITY;
    if (u < units && hl < LP) {
      int ql, qxx, qyy;
      unit_query(u, ql, qxx, qyy);
      const int q = ms_pick(ql, st0, st1, st2) + qyy * ms_pick(ql, W0, W1, W2) + qxx;
      const TO* owrow = reinterpret_cast<const TO*>(p.ow) + ((size_t)b * p.S + q) * OW_ROW;
      if constexpr (sizeof(TO) == 2) {
        float a, c;
        unpack2<TO>(__ldg(reinterpret_cast<const uint32_t*>(owrow + m * LP * 2) + hl), a, c);
        ox = a;
        oy = c;
      } else {
        const float2 o2 = __ldg(reinterpret_cast<const float2*>(owrow + m * LP * 2) + hl);
        ox = o2.x;
        oy = o2.y;
      }
      lg = to_f32<TO>(owrow[M * LP * 2 + m * LP + hl]);
    }
  };
  float nox, noy, nlg;
  fetch(warp * 2, nox, noy, nlg);
  for (int u0 = warp * 2; u0 < units; u0 += WARPS * 2) {
    // =============== parameters of unit u0 + half, one sample per lane (hl < 12) ===============
    const int u = u0 + half;
    const bool uvalid = u < units;
    const float ox = nox, oy = noy, lg = nlg;
    fetch(u0 + WARPS * 2, nox, noy, nlg);
    int ql, qxx, qyy;
    unit_query(uvalid ? u : 0, ql, qxx, qyy);
    const int q = ms_pick(ql, st0, st1, st2) + qyy * ms_pick(ql, W0, W1, W2) + qxx;
    // reference point = pixel centre of the query, normalised (msdeformattn.py:76-87)
    const float rx = ((float)qxx + 0.5f) * ms_pick(ql, rW0, rW1, rW2), ry = ((float)qyy + 0.5f) * ms_pick(ql, rH0, rH1, rH2);
    const bool own = uvalid && hl < LP;
    const int sl = own ? (hl >> 2) : 0;         // level of the sample
    // softmax over the 12 logits of the half-warp (F.softmax, ms_deform_attn.py:105), exp2 on the SFU
    float mx = lg;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float e = 0.f;
    if (own) asm("ex2.approx.ftz.f32 %0, %1;\n" : "=f"(e) : "f"((lg - mx) * 1.4426950408889634f));
    float sum = e;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float aw = own ? __fdividef(e, sum) : 0.f;
    // sampling_locations = ref + off / (W_l, H_l) (ms_deform_attn.py:109-110), x_im = loc * W - 0.5 (im2col :290-291):
    // (rx + ox / W) * W - 0.5 = rx * W + ox - 0.5 (one rounding fewer than the reference's division; 16-bit storage)
    const int Wl = ms_pick(sl, W0, W1, W2), Hl = ms_pick(sl, H0, H1, H2);
    const float x = fmaf(rx, (float)Wl, ox - 0.5f);
    const float y = fmaf(ry, (float)Hl, oy - 0.5f);
    float w00 = 0.f, w01 = 0.f, w10 = 0.f, w11 = 0.f;
    int x0 = 0, y0 = 0, p00 = 0;
    bool slow = false;
    if (own && y > -1.f && x > -1.f && y < (float)Hl && x < (float)Wl) {
      const float yf = floorf(y), xf = floorf(x);
      y0 = (int)yf;
      x0 = (int)xf;
      const float ly = y - yf, lx = x - xf;
      const float wy1 = ly * aw, wy0 = aw - wy1;
      w01 = wy0 * lx; w00 = wy0 - w01; w11 = wy1 * lx; w10 = wy1 - w11;
      const int px = x0 - ms_pick(sl, tx00, tx01, tx02), py = y0 - ms_pick(sl, ty00, ty01, ty02);
      const int twl = ms_pick(sl, tw0, tw1, tw2);
      if (px >= 0 && py >= 0 && px + 1 < twl && py + 1 < ms_pick(sl, th0, th1, th2)) p00 = py * twl + px;
      else slow = true;          // this sample reads global memory below; it contributes nothing to the MMA
    }
    if (hl < LP) {
      uint4 pk = make_uint4(0, 0, 0, 0);
      if (!slow) {
        pk.x = pack2<T>(w00, w01);
        pk.y = pack2<T>(w10, w11);
        float h0, h1, h2, h3;
        unpack2<T>(pk.x, h0, h1);
        unpack2<T>(pk.y, h2, h3);
        pk.z = pack2<T>(w00 - h0, w01 - h1);
        pk.w = pack2<T>(w10 - h2, w11 - h3);
      }
      stg_w[half * LP + hl] = pk;
      stg_p[half * 16 + hl] = p00;
    }
    const uint32_t slow_mask = __ballot_sync(0xffffffffu, slow);
    __syncwarp();

    // =============== contraction: unit A (half 0), then unit B (half 1); every lane takes part ===============
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      if (u0 + hh >= units) break;
      float d[2][4];
#pragma unroll
      for (int i = 0; i < 4; ++i) d[0][i] = d[1][i] = 0.f;
      const uint32_t* w32 = reinterpret_cast<const uint32_t*>(stg_w + hh * LP);
#pragma unroll
      for (int ks = 0; ks < L; ++ks) {
        // A: 16 corners x 16 channels per m-tile, gathered: lane i addresses corner (i & 7) + 8 (i >> 4), channel
        // block (i >> 3) & 1 (+ 2 for the second m-tile)
        const int c = (lane & 7) + 8 * (lane >> 4);
        const int pidx = stg_p[hh * 16 + ks * 4 + (c >> 2)] + (c & 1) + ((c >> 1) & 1) * (ks == 0 ? tw0 : (ks == 1 ? tw1 : tw2));
        const uint32_t sw = (uint32_t)(pidx >> 1) & 3u, cb = (uint32_t)(lane >> 3) & 1u;
        const uint32_t rowaddr = (ks == 0 ? tb0 : (ks == 1 ? tb1 : tb2)) + (uint32_t)pidx * 64u;
        uint32_t a0[4], a1[4];
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                     : "=r"(a0[0]), "=r"(a0[1]), "=r"(a0[2]), "=r"(a0[3]) : "r"(rowaddr + ((cb ^ sw) << 4)));
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                     : "=r"(a1[0]), "=r"(a1[1]), "=r"(a1[2]), "=r"(a1[3]) : "r"(rowaddr + (((2u + cb) ^ sw) << 4)));
        // B: column 0 = hi weights, column 1 = lo weights of corners (2 t4, 2 t4 + 1) and (2 t4 + 8, 2 t4 + 9)
        uint32_t b0 = 0u, b1 = 0u;
        if (g < 2) {
          const int wi = g * 2 + (t4 & 1);
          b0 = w32[(ks * 4 + (t4 >> 1)) * 4 + wi];
          b1 = w32[(ks * 4 + 2 + (t4 >> 1)) * 4 + wi];
        }
        ms_mma<T>(d[0], a0, b0, b1);
        ms_mma<T>(d[1], a1, b0, b1);
      }
      if (t4 == 0) {   // columns 0 (hi) + 1 (lo): channels mt*16 + g and mt*16 + g + 8
        stg_o[hh * 32 + g] = d[0][0] + d[0][1];
        stg_o[hh * 32 + g + 8] = d[0][2] + d[0][3];
        stg_o[hh * 32 + 16 + g] = d[1][0] + d[1][1];
        stg_o[hh * 32 + 24 + g] = d[1][2] + d[1][3];
      }
    }
    __syncwarp();
    // =============== samples outside the staged tile: predicated global loads, lane = channel ===============
    if (slow_mask) {
      uint32_t mk = slow_mask;
      while (mk) {
        const int src = __ffs(mk) - 1;
        mk &= mk - 1;
        const int sx0 = __shfl_sync(0xffffffffu, x0, src), sy0 = __shfl_sync(0xffffffffu, y0, src);
        const int sll = __shfl_sync(0xffffffffu, sl, src);
        const float a00 = __shfl_sync(0xffffffffu, w00, src), a01 = __shfl_sync(0xffffffffu, w01, src);
        const float a10 = __shfl_sync(0xffffffffu, w10, src), a11 = __shfl_sync(0xffffffffu, w11, src);
        const int Ws = ms_pick(sll, W0, W1, W2), Hs = ms_pick(sll, H0, H1, H2);
        const T* vl = vplane + (size_t)ms_pick(sll, st0, st1, st2) * D + lane;
        float acc = 0.f;
        const bool vy0 = sy0 >= 0, vy1 = sy0 + 1 <= Hs - 1, vx0 = sx0 >= 0, vx1 = sx0 + 1 <= Ws - 1;
        if (vy0 && vx0) acc = fmaf(a00, to_f32<T>(vl[((size_t)sy0 * Ws + sx0) * D]), acc);
        if (vy0 && vx1) acc = fmaf(a01, to_f32<T>(vl[((size_t)sy0 * Ws + sx0 + 1) * D]), acc);
        if (vy1 && vx0) acc = fmaf(a10, to_f32<T>(vl[((size_t)(sy0 + 1) * Ws + sx0) * D]), acc);
        if (vy1 && vx1) acc = fmaf(a11, to_f32<T>(vl[((size_t)(sy0 + 1) * Ws + sx0 + 1) * D]), acc);
        stg_o[(src >> 4) * 32 + lane] += acc;
      }
      __syncwarp();
    }
    // =============== store: 4 lanes per unit, 16 B each ===============
    if (hl < 4 && uvalid) {
      float f[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = stg_o[half * 32 + hl * 8 + i];
      store16_from_f32<T>(reinterpret_cast<T*>(p.out) + ((size_t)b * p.S + q) * (M * D) + m * D + hl * 8, f);
    }
    __syncwarp();
  }
}

// ---- host ------------------------------------------------------------------------------------------------------
typedef CUresult (*MsEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static MsEncodeFn ms_encode_fn() {
  static MsEncodeFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<MsEncodeFn>(p);
  }
  return fn;
}

static int ms_floordiv_h(long long a, long long b) { return (int)((a >= 0) ? a / b : -((-a + b - 1) / b)); }

int g_msda_halo = 5;

bool msda_smem_ok(int M, int D, int L, int P, int value_dtype, const int64_t* shapes_host) {
  if (M != ms::M || D != ms::D || L != ms::L || P != ms::P) return false;
  if (value_dtype != PSALM_BF16 && value_dtype != PSALM_F16) return false;
  for (int l = 0; l + 1 < L; ++l)
    if (shapes_host[2 * l] > shapes_host[2 * l + 2] || shapes_host[2 * l + 1] > shapes_host[2 * l + 3]) return false;   // coarse -> fine
  return shapes_host[2 * (L - 1)] * 2 * ms::CELL < (1 << 20) && shapes_host[2 * (L - 1) + 1] * 2 * ms::CELL < (1 << 20);
}

template <typename T, typename TO>
static int launch_smem(const void* value, const void* ow, void* out, const int64_t* shapes_host, const int64_t* starts_host,
                       int B, int S, cudaStream_t st) {
  using namespace ms;
  MsParams p;
  p.value = value; p.ow = ow; p.out = out; p.B = B; p.S = S; p.halo = g_msda_halo;
  for (int l = 0; l < L; ++l) {
    p.H[l] = (int)shapes_host[2 * l];
    p.W[l] = (int)shapes_host[2 * l + 1];
    p.start[l] = (int)starts_host[l];
  }
  const int Wf = p.W[L - 1], Hf = p.H[L - 1];
  p.cells_x = (Wf + CELL - 1) / CELL;
  p.cells_y = (Hf + CELL - 1) / CELL;
  int off = 0;
  for (int l = 0; l < L; ++l) {
    int tw = 0, th = 0;
    for (int c = 0; c < p.cells_x; ++c) {
      const int lo = ms_floordiv_h(2LL * CELL * c * p.W[l] - Wf, 2LL * Wf) - p.halo;
      const int hi = ms_floordiv_h(2LL * CELL * (c + 1) * p.W[l] - Wf, 2LL * Wf) + 1 + p.halo;
      if (hi - lo + 1 > tw) tw = hi - lo + 1;
    }
    for (int c = 0; c < p.cells_y; ++c) {
      const int lo = ms_floordiv_h(2LL * CELL * c * p.H[l] - Hf, 2LL * Hf) - p.halo;
      const int hi = ms_floordiv_h(2LL * CELL * (c + 1) * p.H[l] - Hf, 2LL * Hf) + 1 + p.halo;
      if (hi - lo + 1 > th) th = hi - lo + 1;
    }
    while ((tw & 7) < 2 || (tw & 7) > 6) ++tw;   // the 4 corners of a sample in 4 different 16-byte bank groups
    PSALM_REQUIRE(tw <= 256 && th <= 256, "msda(smem): tile %d x %d exceeds the TMA box limit", tw, th);
    p.tw[l] = tw; p.th[l] = th; p.toff[l] = off;
    off += (tw * th * D * (int)sizeof(T) + 1023) / 1024 * 1024;
  }
  p.stage_off = off;
  const size_t smem = 1024 + (size_t)off + WARPS * 1024;
  PSALM_REQUIRE(smem <= 113 * 1024, "msda(smem): %zu bytes of shared memory per CTA (halo %d) - lower the halo", smem, p.halo);
  MsEncodeFn enc = ms_encode_fn();
  PSALM_REQUIRE(enc != nullptr, "msda(smem): cuTensorMapEncodeTiled unavailable");
  CUtensorMap maps[L];
  for (int l = 0; l < L; ++l) {
    const cuuint64_t dims[4] = {(cuuint64_t)D, (cuuint64_t)p.W[l], (cuuint64_t)p.H[l], (cuuint64_t)B * M};
    const cuuint64_t strides[3] = {(cuuint64_t)D * sizeof(T), (cuuint64_t)p.W[l] * D * sizeof(T), (cuuint64_t)S * D * sizeof(T)};
    const cuuint32_t box[4] = {(cuuint32_t)D, (cuuint32_t)p.tw[l], (cuuint32_t)p.th[l], 1};
    const cuuint32_t es[4] = {1, 1, 1, 1};
    void* base = const_cast<char*>(reinterpret_cast<const char*>(value)) + (size_t)p.start[l] * D * sizeof(T);
    const CUresult r = enc(&maps[l], std::is_same<T, __nv_bfloat16>::value ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16,
                           4, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    PSALM_REQUIRE(r == CUDA_SUCCESS, "msda(smem): cuTensorMapEncodeTiled failed for level %d (code %d)", l, (int)r);
  }
  auto kern = msda_smem_kernel<T, TO>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    set_error("msda(smem): cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    return PSALM_E_CUDA;
  }
  dim3 grid(p.cells_x * p.cells_y, M, B);
  kern<<<grid, THREADS, smem, st>>>(maps[0], maps[1], maps[2], p);
  return check_launch("msda_smem_kernel");
}

int msda_smem_fused(const void* value, const void* ow, void* out, const int64_t* shapes_host, const int64_t* starts_host, int B,
                    int S, int value_dtype, int ow_dtype, cudaStream_t st) {
  if (value_dtype == PSALM_BF16 && ow_dtype == PSALM_BF16) return launch_smem<__nv_bfloat16, __nv_bfloat16>(value, ow, out, shapes_host, starts_host, B, S, st);
  if (value_dtype == PSALM_BF16 && ow_dtype == PSALM_F32) return launch_smem<__nv_bfloat16, float>(value, ow, out, shapes_host, starts_host, B, S, st);
  if (value_dtype == PSALM_F16 && ow_dtype == PSALM_F16) return launch_smem<__half, __half>(value, ow, out, shapes_host, starts_host, B, S, st);
  if (value_dtype == PSALM_F16 && ow_dtype == PSALM_F32) return launch_smem<__half, float>(value, ow, out, shapes_host, starts_host, B, S, st);
  set_error("msda(smem): unsupported dtype combination value=%d ow=%d", value_dtype, ow_dtype);
  return PSALM_E_UNSUPPORTED;
}

}  // namespace psalm

extern "C" int psalm_set_msda_halo(int halo) {
  if (halo < 0 || halo > 16) {
    psalm::set_error("psalm_set_msda_halo: 0..16");
    return PSALM_E_ARG;
  }
  psalm::g_msda_halo = halo;
  return PSALM_OK;
}
