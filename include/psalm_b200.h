/*
 * psalm_b200 — C ABI of the B200-native PSALM inference hot path.
 *
 * Plain pointers and sizes only (no torch / ATen types).  Every entry point
 *   - takes DEVICE pointers unless the parameter name ends in `_host`,
 *   - enqueues work on `stream` (a cudaStream_t passed as void*) and never synchronises,
 *   - returns 0 on success or a negative PSALM_E_* code; psalm_last_error() holds the message.
 *     (The reference only printf()s launch failures — ms_deform_im2col_cuda.cuh:953-957 — and its
 *      Python caller swallows every exception — ops/modules/ms_deform_attn.py:117; we never do.)
 *
 * Reference interfaces each entry point replaces are cited per function as
 * (path relative to /root/reference/psalm/model/…:line).
 */
#ifndef PSALM_B200_H_
#define PSALM_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PSALM_ABI_VERSION 1

/* element types */
enum { PSALM_F32 = 0, PSALM_F16 = 1, PSALM_BF16 = 2 };

/* error codes */
enum {
  PSALM_OK = 0,
  PSALM_E_ARG = -1,      /* bad argument (shape, dtype, alignment, null pointer) */
  PSALM_E_UNSUPPORTED = -2,
  PSALM_E_CUDA = -3      /* CUDA runtime / launch error */
};

int psalm_abi_version(void);
const char* psalm_last_error(void);
/* compute capability the library was compiled for (100 => sm_100a) */
int psalm_compiled_arch(void);

/* ------------------------------------------------------------------------------------------
 * Multi-scale deformable attention, forward (sampling + aggregation).
 *
 * Replaces: ms_deform_attn_forward (mask_decoder/Mask2Former_Simplify/modeling/pixel_decoder/
 *   ops/src/ms_deform_attn.h:25-44 -> ops/src/cuda/ms_deform_attn_cuda.cu:25-85 ->
 *   ms_deformable_im2col_gpu_kernel, ops/src/cuda/ms_deform_im2col_cuda.cuh:243-304).
 *
 *   out[b,q,m,:] = sum_{l,p} w[b,q,m,l,p] * bilinear(value_l[b,:,m,:], loc[b,q,m,l,p,:])
 *   zero padding outside the map, pixel-centre convention (x_im = loc_x * W_l - 0.5).
 *
 *   value   [B,S,M,D]      (value_layout 0, the reference layout)   dtype value_dtype
 *           [B,M,S,D]      (value_layout 1, head-major, used by the fused pipeline)
 *   shapes  [L,2] int64 (H_l, W_l), starts [L] int64 — DEVICE pointers as in the reference
 *           (spatial_shapes.data<int64_t>()), or HOST pointers when shapes_on_host != 0.
 *   loc     [B,Lq,M,L,P,2], w [B,Lq,M,L,P]                          dtype loc_dtype
 *           (the reference requires loc_dtype == value_dtype; PSALM_F32 is always accepted)
 *   out     [B,Lq,M*D]                                              dtype value_dtype
 *   All tensors contiguous; out is fully overwritten (no pre-zeroing needed).
 * ------------------------------------------------------------------------------------------ */
int psalm_msda_forward(const void* value, const int64_t* shapes, const int64_t* starts,
                       const void* loc, const void* w, void* out,
                       int B, int S, int M, int D, int L, int Lq, int P,
                       int value_dtype, int loc_dtype, int value_layout, int shapes_on_host,
                       void* stream);

/* Fused sampling for the encoder layer: takes the raw outputs of the sampling_offsets /
 * attention_weights Linear layers and does softmax(L*P) + reference-point generation
 * (get_reference_points, pixel_decoder/msdeformattn.py:76-87, valid_ratios == 1) +
 * location arithmetic (ops/modules/ms_deform_attn.py:104-110) + sampling in one kernel, so
 * `sampling_locations` and `attention_weights` never exist in HBM.
 *
 *   value  [B,M,S,D] head-major, dtype value_dtype
 *   ow     [B,Lq,M*L*P*3]: per query, first M*L*P*2 offsets (m,l,p,xy order), then M*L*P logits
 *          dtype ow_dtype (F32 / F16 / BF16).  Lq must equal S (encoder self-attention).
 *   out    [B,Lq,M*D] dtype value_dtype
 *   shapes_host/starts_host: HOST int64 arrays. */
int psalm_msda_encoder_fused(const void* value, const void* ow, void* out,
                             const int64_t* shapes_host, const int64_t* starts_host,
                             int B, int S, int M, int D, int L, int P,
                             int value_dtype, int ow_dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PSALM_B200_H_ */
