/*
 * versband.h — C ABI of libversband_sm100.so (B200 / sm_100a only).
 *
 * Drop-in boundary for the AccompBand inference hot path of AaronZ345/VersBand.
 * The reference is pure PyTorch: every entry point below replaces the ATen / cuBLAS /
 * cuDNN / SDPA call that the cited reference line reaches.  All pointers are DEVICE
 * pointers borrowed for the duration of the call (the caller — PyTorch's caching
 * allocator in our host layer — owns every buffer), tensors are dense/strided as
 * documented per field, every call takes an explicit cudaStream_t (passed as void*),
 * returns 0 on success or a negative vb_status, and never allocates device memory.
 * vb_last_error() returns a thread-local message for the last failing call.
 *
 * There is no CPU fallback anywhere in this library.
 */
#ifndef VERSBAND_H_
#define VERSBAND_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VB_ABI_VERSION 1

typedef enum vb_status {
  VB_OK = 0,
  VB_ERR_INVALID = -1, /* bad argument (shape / alignment / null) */
  VB_ERR_CUDA = -2,    /* CUDA runtime / driver error, see vb_last_error() */
  VB_ERR_ARCH = -3     /* device is not sm_100 */
} vb_status;

/* activation codes used by vb_gemm / elementwise ops */
enum {
  VB_ACT_NONE = 0,
  VB_ACT_SILU = 1,     /* x*sigmoid(x): adaLN SiLU, VAE swish (autoencoder1d.py:168-170) */
  VB_ACT_LRELU = 2,    /* leaky relu with slope (hifigan.py:8,53; vocal2music_moe.py:342) */
  VB_ACT_GELU_ERF = 3, /* exact GELU (flag_large_dit_moe.py:154) */
  VB_ACT_TANH = 4      /* hifigan.py:141 */
};

/* vb_gemm epilogue modes */
enum {
  VB_EPI_GENERIC = 0,
  VB_EPI_QKV_ROPE = 1 /* fused q/k rotary + V transpose (flag_large_dit_moe.py:340-352) */
};

int vb_version(void);
const char* vb_last_error(void);
/* 0 if the current device can run this library (compute capability 10.x), else VB_ERR_ARCH */
int vb_check_device(void);

/* ------------------------------------------------------------------------------------------
 * vb_gemm — the tensor-core workhorse (tcgen05.mma, TMEM accumulator, TMA-fed smem pipeline).
 *
 *   D[g][m, n] = epilogue( sum_k A[g][row(m,k), k] * B[g][n, k] )      bf16 x bf16 -> fp32
 *
 * Replaces every nn.Linear / F.linear on the path (flag_large_dit_moe.py:340,389-390,406,
 * 484-485; vocal2music_moe.py:119,132,141-142,242,290), the bmm's of attention
 * (flag_large_dit_moe.py:381-386; autoencoder1d.py:262-270) and — through the conv
 * addressing below — every Conv1d / ConvTranspose1d (vocal2music_moe.py:327,340-350,358;
 * autoencoder1d.py:180-231,276-313,432-478; vocoder/hifigan/modules/hifigan.py:27-58,108-141).
 *
 * A is a rank-4 strided bf16 tensor (dim0 = reduction index, contiguous). For k-block kb
 * (64 elements) of tile row m0 the TMA box is fetched at coordinates
 *     c0 = (kb % kpt) * 64,  c1 = m0 + (kb / kpt) * dil - pad,
 *     c2 = g1*a_c2_g1 + g2*a_c2_g2,  c3 = g1*a_c3_g1 + g2*a_c3_g2
 * Out-of-range coordinates read as zero: this gives "same" conv padding, ragged M/K tails
 * and channel padding for free. A plain GEMM uses kpt = K/64, dil = pad = 0.
 * B is a rank-3 strided bf16 tensor (k, n, group), i.e. PyTorch's [out, in] weight layout,
 * fetched at (kb*64, n0, g1*b_c2_g1 + g2*b_c2_g2 + tile_group[m_tile]*b_c2_tile).
 * ------------------------------------------------------------------------------------------ */
typedef struct vb_gemm_args {
  const void* a;
  int64_t a_dims[4];
  int64_t a_strides[4]; /* elements; a_strides[0] == 1; others multiples of 8 */
  const void* b;
  int64_t b_dims[3];
  int64_t b_strides[3]; /* elements; b_strides[0] == 1; others multiples of 8 */

  int32_t M, N, K; /* rows / cols per group; K = padded reduction length, multiple of 64 */
  int32_t G1, G2;  /* grid.z = G1*G2, g1 = z % G1, g2 = z / G1 */
  int32_t kpt, dil, pad;
  int32_t a_c2_g1, a_c2_g2, a_c3_g1, a_c3_g2;
  int32_t b_c2_g1, b_c2_g2, b_c2_tile;
  int32_t bn; /* N-tile override (32/64/96/128), 0 = auto */

  /* MoE grouped mode (optional): per-M-tile group id and device-side tile count */
  const int32_t* tile_group;
  const int32_t* m_tiles_used;

  /* ---- epilogue --------------------------------------------------------------------- */
  int32_t mode;  /* VB_EPI_* */
  float alpha;   /* v = acc * alpha */
  int32_t swiglu; /* 1: columns come in (w1,w3) pairs; v[n/2] = silu(v[2i]) * v[2i+1] */
  const float* bias; /* v += bias[n + g1*bias_g1 + tile_group*bias_tile] */
  int64_t bias_g1, bias_tile;
  int32_t act;
  float act_slope; /* v = act(v) */
  const float* row_scale; /* v *= row_scale[m]  (indexed by tile row m, before row mapping) */
  const float* gate;      /* v *= gate[(m / gate_rows) * gate_ld + n] */
  int32_t gate_rows;
  int64_t gate_ld;
  /* output row: orow = row_map ? row_map[m] : m*row_mul + row_off + g1*row_off_g1;
   * rows with orow < 0 or orow >= rows_out are dropped */
  const int32_t* row_map;
  int32_t row_mul, row_off, row_off_g1, rows_out;
  const float* residual; /* v += residual[...] same addressing as out_f32 */
  int32_t accumulate;    /* v += out_f32[...] (previous contents) */
  float* out_f32;        /* optional fp32 store */
  int64_t out_ld, out_g1, out_g2; /* elements */
  /* optional bf16 store of act_out(v * out_scale): the next GEMM's A operand */
  void* out_bf16;
  int64_t outb_ld, outb_g1, outb_g2;
  int32_t act_out;
  float act_out_slope, out_scale;

  /* ---- VB_EPI_QKV_ROPE: N = 3*heads*head_dim, rows m = b*T + t ----------------------- */
  const float* rope_cs; /* [T][head_dim/2][2] (cos, sin), fp32 */
  int32_t T, heads, head_dim, Tp;
  void* q_out;  /* bf16 [B, heads, T, head_dim] (rotated) */
  void* k_out;  /* bf16 [B, heads, T, head_dim] (rotated) */
  void* vt_out; /* bf16 [B, heads, head_dim, Tp]  (V transposed: PV GEMM's K-major B) */
} vb_gemm_args;

int vb_gemm(const vb_gemm_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VERSBAND_H_ */
