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
  VB_ACT_TANH = 4,     /* hifigan.py:141 */
  VB_ACT_LOG10 = 5     /* log10(max(x, act_slope)): mel dynamic-range compression, preprocess/NAT_mel.py:25-26 */
};

/* vb_gemm epilogue modes */
enum {
  VB_EPI_GENERIC = 0,
  VB_EPI_QKV_ROPE = 1, /* fused q/k rotary + V transpose (flag_large_dit_moe.py:340-352) */
  VB_EPI_ROUTE = 2     /* short-KV attention folded into a gate: see the route_* fields */
};

int vb_version(void);
const char* vb_last_error(void);
/* 0 if the current device can run this library (compute capability 10.x), else VB_ERR_ARCH */
int vb_check_device(void);

/* ------------------------------------------------------------------------------------------
 * vb_gemm — the tensor-core workhorse (tcgen05.mma, TMEM accumulator, TMA-fed smem pipeline).
 *
 *   D[g][m, n] = epilogue( sum_k A[g][row(m,k), k] * B[g][n, k] )      fp16 x fp16 -> fp32
 *
 * Replaces every nn.Linear / F.linear on the path (flag_large_dit_moe.py:340,389-390,406,
 * 484-485; vocal2music_moe.py:119,132,141-142,242,290), the bmm's of attention
 * (flag_large_dit_moe.py:381-386; autoencoder1d.py:262-270) and — through the conv
 * addressing below — every Conv1d / ConvTranspose1d (vocal2music_moe.py:327,340-350,358;
 * autoencoder1d.py:180-231,276-313,432-478; vocoder/hifigan/modules/hifigan.py:27-58,108-141).
 *
 * A is a rank-4 strided fp16 tensor (dim0 = reduction index, contiguous). For k-block kb
 * (64 elements) of tile row m0 the TMA box is fetched at coordinates
 *     c0 = (kb % kpt) * 64,  c1 = m0 + (kb / kpt) * dil - pad,
 *     c2 = g1*a_c2_g1 + g2*a_c2_g2,  c3 = g1*a_c3_g1 + g2*a_c3_g2
 * Out-of-range coordinates read as zero: this gives "same" conv padding, ragged M/K tails
 * and channel padding for free. A plain GEMM uses kpt = K/64, dil = pad = 0.
 * B is a rank-3 strided fp16 tensor (k, n, group), i.e. PyTorch's [out, in] weight layout,
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
  int32_t bn; /* N-tile override (32/64/96/128/192/256; forces the 1-CTA kernel), 0 = auto (cost model, may pick the 2-CTA kernel) */
  int32_t operand_f32; /* 1: A and B are fp32 tensors multiplied as TF32 (kind::tf32); K multiple of 32,
                          strides in fp32 elements; the accuracy mode of DESIGN.md section 2 */

  /* MoE grouped mode (optional): per-M-tile group id and device-side tile count */
  const int32_t* tile_group;
  const int32_t* m_tiles_used;
  const int32_t* tile_row0; /* optional: first A/output row of M tile i (sparse layouts: fixed-capacity expert
                               buckets, vb_route_dispatch); NULL = i * 128 */

  /* ---- epilogue --------------------------------------------------------------------- */
  int32_t mode;  /* VB_EPI_* */
  float alpha;   /* v = acc * alpha */
  int32_t swiglu; /* gated unit over (w1,w3) column pairs: 1 = v[n/2] = silu(v[2i]) * v[2i+1] (SwiGLU FeedForward);
                     2 = gelu_new(v[2i]) * v[2i+1] (T5 v1.1 / FLAN-T5 DenseGatedActDense);
                     3 = sqrt(v[2i]^2 + v[2i+1]^2 + 1e-9): magnitude of interleaved (re, im) STFT columns */
  const float* bias; /* v += bias[n + g1*bias_g1 + tile_group*bias_tile] */
  int64_t bias_g1, bias_tile;
  int32_t act;
  float act_slope; /* v = act(v) */
  const float* row_scale; /* v *= row_scale[m]  (indexed by tile row m, before row mapping) */
  const float* gate;      /* v *= gate[(m / gate_rows) * gate_ld + g1 * gate_g1 + n] */
  int32_t gate_rows;
  int64_t gate_ld, gate_g1;
  /* output row: orow = row_map ? row_map[m] : m*row_mul + row_off + g1*row_off_g1;
   * rows with orow < 0 or orow >= rows_out are dropped */
  const int32_t* row_map;
  int32_t row_mul, row_off, row_off_g1, rows_out;
  const float* residual; /* v += residual[...] same addressing as out_f32 */
  const void* res16;     /* fp16 residual stored in the ACTIVATED representation y = leaky_relu(x): v += (y >= 0 ? y :
                            y * res16_neg), same addressing as out_fp16 (HiFi-GAN residual stream, vocoder.py) */
  float res16_neg;       /* 1 / negative_slope (1 = plain fp16 residual) */
  int32_t accumulate;    /* v += out_f32[...] (previous contents) */
  float* out_f32;        /* optional fp32 store */
  int64_t out_ld, out_g1, out_g2; /* elements */
  /* optional fp16 store of act_out(v * out_scale): the next GEMM's A operand */
  void* out_fp16;
  int32_t outb_f32; /* 1: this slot stores fp32 (the next GEMM runs in operand_f32 mode) */
  int64_t outb_ld, outb_g1, outb_g2;
  int32_t act_out;
  float act_out_slope, out_scale;

  /* ---- VB_EPI_QKV_ROPE: N = 3*heads*head_dim, rows m = b*T + t; v = acc*alpha + bias ---- */
  const float* rope_cs; /* [T][head_dim/2][2] (cos, sin), fp32; NULL = no rotation (VAE attn) */
  int32_t T, heads, head_dim, Tp;
  int32_t qkv_sec0; /* first section present: 0 = q|k|v (N = 3*H*hd), 1 = k|v only (N = 2*H*hd) */
  void* q_out;  /* fp16 [B, heads, T, head_dim] (rotated) */
  void* k_out;  /* fp16 [B, heads, T, head_dim] (rotated) */
  void* vt_out; /* fp16 [B, heads, head_dim, Tp]  (V transposed: PV GEMM's K-major B) */
  /* ---- VB_EPI_ROUTE: the caption cross-attention that feeds the Band-MoE caption gate
   * (vocal2music_moe.py:117-131: MultiheadAttention -> gating Linear) collapsed into this GEMM.
   * B holds, per batch g1 and head h, the 96 rows (scale * K_h W_q,h) (keys beyond the text length are
   * zero rows whose bias is -1e30), so that acc[m, h*96 + j] + bias is the attention score of token m
   * and key j; the epilogue takes the softmax over the 96 columns of a head and contracts it with
   * route_vg[g1][h][j][0..3] = V_h[j,:] . (W_gate W_o)[e, head h] — i.e. it emits, per token and head,
   * the head's contribution to the 4 gate logits: out_f32[(g1*M + m) * out_ld + h*4 + e].
   * Requires N % 96 == 0 (N = heads * 96), batches in G1 (M rows each), bias with bias_g1 = N. */
  const float* route_vg;
  int64_t route_vg_g1; /* elements between batches of route_vg (heads * 96 * 4) */
} vb_gemm_args;

int vb_gemm(const vb_gemm_args* args, void* stream);

/* Acoustic stem of the DiT (Embedding -> Conv1d k -> LeakyReLU -> AvgPool1d(2), midi + beats summed;
 * vocal2music_moe.py:336-350,388-393) as row gathers: tab_x [k][Vx][D] = W[:, :, tap] . emb[v] precomputed at
 * pack time (the vocabularies have 130 / 3 entries), out[b*Tpool + t] = pooled sum written as the split GEMM
 * operand [hi | hi | lo] (3*D columns, fp16 or tf32-rounded fp32) of the following 1x1 final_proj. */
int vb_stem_gather(const int64_t* midi, const int64_t* beats, const float* tab_m, const float* tab_b,
                   const float* bias_m, const float* bias_b, int Vm, int Vb, int k, int B, int Tm, int Tpool, int D,
                   float slope, void* out, int out_is_f32, void* stream);

/* ------------------------------------------------------------------------------------------
 * Formats either side of the path
 * vb_wave_prepare: waveform fp32 [B, n] -> clamp to [-1, 1], reflect-pad `pad` samples each side
 *   (preprocess/NAT_mel.py:62-66) and split into fp16 planes out[b] = [hi | lo | hi], each `plane` samples
 *   (hi + lo = the sample to ~3e-8): the A operand of the windowed-DFT GEMM, whose three "taps" (plane apart)
 *   against [W_hi | W_hi | W_lo] give fp32-class STFT accuracy from fp16 tensor-core operands.
 * vb_loudness_pcm16: per clip rms -> gain to `target_db` dBFS (scripts/test_final.py:342-347) -> int16 PCM
 *   (x * 32767, round to nearest even, saturating), scratch = B * 64 floats.
 * ------------------------------------------------------------------------------------------ */
/* AdamW update of one fp32 parameter tensor (training step): torch.optim.AdamW semantics
 * with bias corrections c1 = 1 - b1^t, c2 = 1 - b2^t computed by the caller. */
/* out[g][c][r] = in[g][r][c], fp32, `batch` matrices of rows x cols; ld_out >= rows, the padding columns are zeroed
 * (operand transposes of the training backward pass: dW = dY^T X reduces over the tokens, vb_gemm operands are K-major) */
int vb_transpose_f32(const float* in, int64_t rows, int64_t cols, int64_t ld_in, int64_t in_batch_stride, float* out,
                     int64_t ld_out, int64_t out_batch_stride, int batch, void* stream);
int vb_adamw(float* p, const float* g, float* m, float* v, int64_t n, float lr, float b1, float b2, float eps, float wd,
             float c1, float c2, void* stream);
int vb_wave_prepare(const float* wav, int B, int64_t n, int pad, int64_t plane, void* out_fp16, void* stream);
int vb_loudness_pcm16(const float* wav, int B, int64_t n, float target_db, float* scratch, int16_t* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * vb_resblock_pair — one HiFi-GAN ResBlock1 conv pair, x' = x + c2(lrelu(c1_{dil}(lrelu(x)) + b1)) + b2
 * (vocoder/hifigan/modules/hifigan.py:51-58), in ONE persistent tcgen05 kernel. Activations are fp16 in the
 * ACTIVATED representation y = leaky_relu_slope(x), channels-last [B, T, C]; y_in feeds conv1 directly, the
 * residual is recovered from the same tile, the intermediate stays in shared memory, y_out = lrelu(x').
 * w1 / w2: packed conv weights [C, k*64] fp16 (tap-major, C_in zero-padded to 64) as produced for vb_gemm;
 * b1 / b2 fp32 [C]. C in {32, 64}, odd k; both convs' weights must fit shared memory next to the tiles
 * (vb_resblock_pair_fits: e.g. C=64 up to k=11, C=32 any HiFi-GAN k). y_out must not alias y_in.
 * vb_mrf_combine: out = lrelu_{slope_out}((x0 + x1 + x2) * scale) over up to three activated-representation
 * tensors (the MRF mean of hifigan.py:132-138), n elements (multiple of 8).
 * ------------------------------------------------------------------------------------------ */
int vb_resblock_pair_fits(int C, int k, int dil);
int vb_resblock_set_trace(void* device_buf); /* debug: 12 tiles x 16 clock64 stamps of CTA 0 (tools/trace_pair.py) */
int vb_resblock_pair(const void* y_in, void* y_out, const void* w1, const void* w2, const float* b1, const float* b2,
                     int B, int T, int C, int k, int dil, float slope, void* stream);
int vb_mrf_combine(const void* r0, const void* r1, const void* r2, float neg, float scale, float slope_out, void* out,
                   int64_t n, void* stream);

/* Router + expert dispatch of one Band-MoE block in ONE launch (MoE.forward, vocal2music_moe.py:132-167):
 * per token the three Gumbel-softmaxes of vb_router (same arithmetic), then the token's operand row `x[n]`
 * (row_bytes each) is copied into its (group, expert) bucket. Bucket k owns rows [k*cap_rows, (k+1)*cap_rows)
 * of xs / row_map / row_scale (cap_rows = N rounded up to 128), so no prefix over bucket counts is needed. The
 * last CTA to finish writes the compact tile table of the grouped GEMMs — tile_group[i], tile_row0[i] for every
 * 128-row tile holding a token, *m_tiles_used —, invalidates the padding rows of partial tiles, reduces the
 * load-balancing loss (:95-115) and re-zeroes bucket_count / usage / done_ticket (all must be zero before the
 * first launch). row_map[r] = n (caption group) or N + n (acoustic group); row_scale[r] = hard weight x mask. */
int vb_route_dispatch(const float* cap_partial, int n_heads, const float* bc, const float* ac_logits, int ac_ld,
                      const float* hl_logits, int hl_ld, const float* e0, const float* e1, const float* e2, float tau,
                      int N, int T, int E, int row_bytes, const void* x, void* xs, int cap_rows, int32_t* idx_c,
                      int32_t* idx_a, float* cap_logits_out, int32_t* bucket_count, float* usage,
                      int32_t* done_ticket, int32_t* next_count, float* next_usage, int32_t* row_map,
                      float* row_scale, int32_t* tile_group, int32_t* tile_row0, int32_t* m_tiles_used,
                      float* lb_out, void* stream);

/* ------------------------------------------------------------------------------------------
 * vb_gemm_chain — up to 4 DEPENDENT vb_gemm problems in ONE persistent launch (1-CTA kernel, one N tile
 * width `bn` in {64, 128} for all of them, epilogues limited to the LINEAR / SWIGLU fusions). The flat tile
 * list runs problem 0, 1, ... in order; dep_mode[o] says what a tile of problem o waits for:
 *   0  nothing (mandatory for problem 0)
 *   1  the same (group, 128-row block) of problem o-1 has stored all its N tiles (A rows of o = output
 *      rows of o-1: expert w1|w3 -> w2, frequency-expert first -> second layer)
 *   2  all tiles of problem o-1 (producers that scatter rows: routed experts' w2 -> frequency experts)
 * The smem ring and the TMEM accumulators run through problem boundaries, so one problem's epilogue tail
 * overlaps the next one's mainloop. `counters` = device int32 scratch the CALLER has zeroed before the launch
 * (counters[0] is set to 1 if a dependency wait ever timed out); needs 2 + sum(groups*m_tiles + 1) ints.
 * Replaces the per-layer launches of FeedForward / the MoE expert loops (vocal2music_moe.py:154-178).
 * ------------------------------------------------------------------------------------------ */
int vb_gemm_chain(const vb_gemm_args* ops, const int32_t* dep_mode, int n_ops, int bn, int32_t* counters,
                  int64_t counters_len, void* stream);

/* ------------------------------------------------------------------------------------------
 * vb_attn_fused — flash-style attention on tcgen05 (S and O_j in TMEM, P through shared memory).
 * Replaces Attention.forward's two SDPA calls and the gated add (flag_large_dit_moe.py:381-402)
 * in one launch, and nn.MultiheadAttention's attention core (vocal2music_moe.py:119):
 *   out[b*T+t, h*hd:(h+1)*hd] = softmax(q k_self^T * scale) v_self
 *                              + tanh(gate[h]) * softmax(q k_text^T * scale) v_text
 * Either segment may be absent (k_self == NULL or k_text == NULL). All operands fp16:
 *   q        element strides (q_st token, q_sh head, q_sb batch), head_dim contiguous
 *   k_self   [B, H, T, hd]      vt_self  [B, H, hd, Tp]   (V transposed, Tp % 8 == 0)
 *   k_text   [B, H, L, hd]      vt_text  [B, H, hd, Lp]   (L <= 128, Lp % 8 == 0)
 * gate: fp32 [H] (raw, tanh applied inside) or NULL (= 1). head_dim in {64, 96}.
 * ------------------------------------------------------------------------------------------ */
typedef struct vb_attn_args {
  const void* q;
  int64_t q_st, q_sh, q_sb;
  const void* k_self;
  const void* vt_self;
  const void* k_text;
  const void* vt_text;
  const float* gate;
  void* out_fp16; /* [B*T, H*hd] */
  float* out_f32; /* optional, same layout */
  int32_t B, T, H, hd, Tp, L, Lp;
  float scale;
} vb_attn_args;

int vb_attn_fused(const vb_attn_args* args, void* stream);
int vb_attn_set_trace(void* device_buf); /* debug: clock64 stamps of CTA (0,0,0) (tools/profile_attn.py) */

/* ------------------------------------------------------------------------------------------
 * Model-level entry point: the whole HiFi-GAN generator (vocoder/hifigan/modules/hifigan.py:126-143) as ONE call
 * that issues the launch sequence from C++ (vb_model.cu). The caller supplies packed weights as device pointers
 * (layouts of versband_b200/vocoder.py:_pack — conv weights [Cout, k * ceil64(Cin)] fp16 tap-major with
 * weight-norm folded, transposed convs as u polyphase filters [u, Cout, Q * ceil64(Cin)], biases fp32), the input
 * mel fp32 [B, n_mel, T], the output wav fp32 [B, T * prod(rates)] and a workspace of vb_voc_workspace_bytes().
 * ------------------------------------------------------------------------------------------ */
#define VB_VOC_MAX_STAGES 8
typedef struct vb_voc_desc {
  int32_t n_mel, c0, pre_k, n_stages, n_kernels, post_k;
  float lrelu, post_b;
  int32_t rates[VB_VOC_MAX_STAGES], up_k[VB_VOC_MAX_STAGES], up_cout[VB_VOC_MAX_STAGES];
  int32_t rb_k[3], n_dil[3], rb_d[3][4]; /* resblock kernel sizes and their dilation lists (ResBlock1) */
  const void* pre_w;
  const float* pre_b;
  const void* up_w[VB_VOC_MAX_STAGES];
  const float* up_b[VB_VOC_MAX_STAGES];
  const void* rb_w1[VB_VOC_MAX_STAGES][3][4]; /* [stage][kernel][pair]: convs1 (dilated) */
  const void* rb_w2[VB_VOC_MAX_STAGES][3][4]; /*                        convs2 (dilation 1) */
  const float* rb_b1[VB_VOC_MAX_STAGES][3][4];
  const float* rb_b2[VB_VOC_MAX_STAGES][3][4];
  const float* post_w; /* [post_k, C_last] fp32 */
} vb_voc_desc;
int64_t vb_voc_workspace_bytes(const vb_voc_desc* d, int B, int T);
int vb_voc_forward(const vb_voc_desc* d, const float* mel, int B, int T, float* wav, void* workspace,
                   int64_t workspace_bytes, void* stream);

/* AutoencoderKL 1-D decoder (reference ldm/models/autoencoder1d.py:480-513, ldm/models/autoencoder1d.py decode =
 * post_quant_conv -> Decoder1D): z fp32 [B, zc, T] (multiplied by `scale`, i.e. 1/scale_factor of
 * ddpm.py decode_first_stage) -> mel fp32 [B, out_ch, T * 2^n_up]. The decoder is described as conv_in, a flat
 * list of layers in execution order (mid.block_1, mid.attn_1, mid.block_2, then per level: resblocks (+ attention)
 * and the upsample), norm_out and conv_out. Packed weights as in versband_b200/vae.py:_pack: conv weights
 * [Cout, k * ceil64(Cin)] fp16 tap-major, q|k|v 1x1 convs concatenated to [3C, C], biases and GroupNorm affine
 * fp32. */
#define VB_VAE_MAX_OPS 48
#define VB_VAE_RES 0  /* ResnetBlock1D: n1 -> w1(k1) -> n2 -> w2(k2) + (ws ? ws(ks) x : x) */
#define VB_VAE_ATTN 1 /* AttnBlock1D: n1 -> w1 = q|k|v [3C, C], b1 [3C] -> softmax(q k^T / sqrt(C)) v -> w2 = proj_out + x */
#define VB_VAE_UP 2   /* Upsample1D: nearest x2 -> w1(k1) */
#define VB_VAE_DOWN 3 /* Downsample1D (encoder): pad (0,1) + conv k3 stride 2 as a 2-tap conv over row pairs, w1 [C, 2 * 2C] */
typedef struct vb_vae_op {
  int32_t kind, cin, cout, k1, k2, ks;
  const float *n1_w, *n1_b, *n2_w, *n2_b;
  const void *w1, *w2, *ws;
  const float *b1, *b2, *bs;
} vb_vae_op;
typedef struct vb_vae_desc {
  int32_t zc, zq_c, c0, out_ch, groups, n_ops, pq_k, in_k, out_k;
  float eps;
  const void* pq_w; /* post_quant_conv [zq_c, pq_k * ceil64(zc)] */
  const float* pq_b;
  const void* in_w; /* decoder.conv_in [c0, in_k * ceil64(zq_c)] */
  const float* in_b;
  const float *no_w, *no_b; /* decoder.norm_out */
  const void* out_w;        /* decoder.conv_out [out_ch, out_k * ceil64(C_last)] */
  const float* out_b;
  vb_vae_op ops[VB_VAE_MAX_OPS];
} vb_vae_desc;
int64_t vb_vae_workspace_bytes(const vb_vae_desc* d, int B, int T);
int vb_vae_decode(const vb_vae_desc* d, const float* z, float scale, int B, int T, float* mel, void* workspace,
                  int64_t workspace_bytes, void* stream);
/* The encoder through the same descriptor type (reference autoencoder1d.py:49-53,384-409): zc = in_ch,
 * in_w = encoder.conv_in, ops = resblocks / attention / VB_VAE_DOWN, no_* = encoder.norm_out, out_w = encoder.conv_out
 * (out_ch = 2 * z_channels), pq_w = quant_conv applied LAST (zq_c = 2 * embed_dim).
 * x fp32 [B, in_ch, T] -> moments fp32 [B, zq_c, T / 2^n_down] (mean | logvar; DiagonalGaussianDistribution.mode() is
 * the first half). */
int64_t vb_vae_encode_workspace_bytes(const vb_vae_desc* d, int B, int T);
int vb_vae_encode(const vb_vae_desc* d, const float* x, int B, int T, float* moments, void* workspace,
                  int64_t workspace_bytes, void* stream);

/* One evaluation of the AccompBand DiT (reference TxtFlagLargeDiT.forward, vocal2music_moe.py:375-434) as ONE call:
 * the launch sequence of versband_b200/dit.py:DiTEngine.forward issued from C++ (vb_model.cu). Weights are device
 * pointers in the packed layouts of DiTEngine.pack (operand dtype = fp16, or fp32 with operand_f32 = 1); the
 * step-invariant conditioning (DiTEngine.prepare: acoustic stem, caption K/V of every block, folded caption gate)
 * is a vb_dit_cond; per-call inputs and outputs are a vb_dit_io. B = rows of the conditioning batch (the CFG pair
 * of one clip is B = 2), N = B * T tokens.
 * Workspace: vb_dit_workspace_bytes(d, B, T) bytes, ZERO-FILLED by the caller before its first use and then left to
 * this (d, B, T) — it carries the MoE bucket counters from one call to the next. */
#define VB_DIT_MAX_DEPTH 16
typedef struct vb_dit_block {
  const void *wqkv, *wo;                      /* [3D, D], [D, D] */
  const float *gate, *attn_norm, *ffn_norm;   /* [H], [D], [D] */
  const float* capg_b;                        /* [E]: W_gate b_out_proj + b_gate */
  const void *w13, *w2;                       /* routed experts [2E, 2Fh, D] ((w1,w3) row pairs), [2E, D, Fh] */
  const void *f13, *f2;                       /* frequency experts [E, 2Fh, 2*band], [E, band, Fh] */
} vb_dit_block;
typedef struct vb_dit_desc {
  int32_t C, D, depth, H, E, Fh, max_len, proj_in_k, operand_f32;
  float eps, temperature;
  int64_t mod_ld;                             /* depth*6D + 2D */
  const void* proj_in_w;                      /* [D, proj_in_k * kq] */
  const float* proj_in_b;
  const void *t_w0, *t_w2, *ada_w;            /* [D,256], [D,D], [mod_ld, D] */
  const float *t_b0, *t_b2, *ada_b, *t_freqs; /* t_freqs [128] */
  const float *hl_w, *hl_b;                   /* group gates of all blocks [2*depth, D], [2*depth] */
  const void* final_w;                        /* [32, D] (rows >= C zero) */
  const float* final_b;                       /* [32] */
  const float* rope;                          /* [max_len, hd/2, 2] */
  vb_dit_block blocks[VB_DIT_MAX_DEPTH];
} vb_dit_desc;
typedef struct vb_dit_cond {
  int32_t B, T, L, Lp;
  const float *acoustic, *ac_logits, *cap_emb; /* [N, D], [N, depth*E], [B, D] */
  const void* k_y[VB_DIT_MAX_DEPTH];           /* fp16 [B, H, L, hd] */
  const void* vt_y[VB_DIT_MAX_DEPTH];          /* fp16 [B, H, hd, Lp] */
  const void* wk_m[VB_DIT_MAX_DEPTH];          /* operand dtype [B, H*96, D] */
  const float* bs_m[VB_DIT_MAX_DEPTH];         /* [B, H, 96] */
  const float* vg_m[VB_DIT_MAX_DEPTH];         /* [B, H, 96, E] */
} vb_dit_cond;
typedef struct vb_dit_io {
  const float* x;     /* latent fp32 [Bx, C, T]; it feeds every group of Bx conditioning rows (CFG batching) */
  int32_t Bx;
  int32_t x_cl_ready; /* 1: x_cl already holds the channels-last latent of all B rows (x is not read) */
  void* x_cl;         /* optional caller-owned [B, T, 32] operand-dtype buffer (NULL: workspace) */
  const int64_t* t;   /* [B] integer timesteps; read only when mod / hl are NULL */
  const float* mod;   /* optional precomputed adaLN vectors [B, mod_ld] ... */
  const float* hl;    /* ... and group-gate logits [B, 2*depth] (the sampler precomputes all steps at once) */
  const float* e0[VB_DIT_MAX_DEPTH]; /* Exp(1) draws of the Gumbel router per block: [N, 2], */
  const float* e1[VB_DIT_MAX_DEPTH]; /* [N, E] (caption group), */
  const float* e2[VB_DIT_MAX_DEPTH]; /* [N, E] (acoustic group) */
  float* v;           /* out: velocity, channels-last fp32 [N, 32] (first C columns valid) */
  float* out_v;       /* optional out: [B, C, T] */
  float* lb;          /* out: load-balance loss per block [depth] */
} vb_dit_io;
int64_t vb_dit_workspace_bytes(const vb_dit_desc* d, int B, int T);
int vb_dit_forward(const vb_dit_desc* d, const vb_dit_cond* cond, const vb_dit_io* io, void* workspace,
                   int64_t workspace_bytes, void* stream);

/* Step-invariant conditioning of the DiT as ONE call (versband_b200/dit.py:DiTEngine.prepare / prepare_pair):
 * acoustic stem of the Bm clips (midi / beats int64 [Bm, 1, Tm]) replicated to every group of Bm conditioning rows,
 * caption embedder over the B = groups * Bm captions (fp32 [B, L, ori_dim], L <= 96), and per block the text
 * K / V^T and the folded caption gate operands. `out` describes caller-allocated OUTPUT buffers (B, T, L,
 * Lp = ceil8(L) set; every pointer of vb_dit_cond set and written here); caption_out: optional fp32 [B, L, D]. */
typedef struct vb_dit_prep_block {
  const float* y_norm;       /* [D] attention_y_norm */
  const void* wkvy;          /* [2D, D]: wk_y | wv_y */
  const void* mkv_w;         /* [2D, D]: k | v rows of the MoE cross-attention in_proj */
  const float* mkv_b;        /* [2D] */
  const void* mq_wT;         /* [H, D, ceil_kq(hd)]: (scale W_q,h)^T */
  const float *mq_bs, *capg_w; /* [D] scale * b_q; [E, D] W_gate W_out_proj */
} vb_dit_prep_block;
typedef struct vb_dit_prep_desc {
  int32_t ori_dim, stem_k, Vm, Vb;
  const float *midi_tab, *beats_tab;       /* [stem_k, V, D]: conv tap x embedding row, folded in fp64 */
  const float *midi_proj_b, *beats_proj_b; /* [D] */
  const void* final_proj_w3;               /* [D, 3D] = [W_hi | W_lo | W_hi] */
  const float* final_proj_b;
  const float *acg_w, *acg_b;              /* acoustic gates of all blocks [depth*E, D], [depth*E] */
  const void *c_w0, *c_w2, *cap_w;         /* [D, ori_dim], [D, D], [D, D] */
  const float *c_b0, *c_b2, *c_ln_w, *c_ln_b, *cap_ln_w, *cap_ln_b, *cap_b;
  vb_dit_prep_block blocks[VB_DIT_MAX_DEPTH];
} vb_dit_prep_desc;
int64_t vb_dit_prepare_workspace_bytes(const vb_dit_desc* d, const vb_dit_prep_desc* p, int Bm, int Tm, int B, int L, int T);
int vb_dit_prepare(const vb_dit_desc* d, const vb_dit_prep_desc* p, const int64_t* midi, const int64_t* beats, int Bm, int Tm,
                   const float* caption, int B, int L, const vb_dit_cond* out, float* caption_out, void* workspace,
                   int64_t workspace_bytes, void* stream);

/* The classifier-free-guidance flow sampler as ONE call (reference cfm1_audio_sampler.py:87-116 + cfm1_audio.py:145-162
 * + torchdyn's fixed-step Euler; versband_b200/sampler.py:FlowSampler._run): n_steps x { DiT evaluation of the
 * stacked (cond, uncond) batch; x += dt * (v_u + guidance * (v_c - v_u)) }. cond2 holds 2 * Bc rows, cond rows
 * first. The router's Exp(1) draws come from the caller (vb_router_noise reproduces torch's Philox stream).
 * Workspace: vb_sample_workspace_bytes() bytes, ZERO-FILLED once before first use (see vb_dit_forward). */
typedef struct vb_sample_args {
  int32_t n_steps; /* Euler steps (24 for the reference's 25-point t_span) */
  int32_t Bc;      /* clips */
  float guidance;
  const int64_t* t_all; /* device [n_steps, 2*Bc]: integer timestep trunc(1000 t) of every step and row (cfm1_audio.py:156) */
  const float* dt;      /* HOST [n_steps]: Euler step sizes, derived from t_span in fp32 like torchdyn */
  const float *e0, *e1, *e2; /* device [n_steps, depth, 2*Bc*T, 2 | E | E] */
  float* x;             /* device fp32 [Bc, C, T]: x0 in, x1 out */
  float* lb;            /* optional device [depth]: load-balance loss of the last evaluation */
} vb_sample_args;
int64_t vb_sample_workspace_bytes(const vb_dit_desc* d, int Bc, int T, int n_steps);
int vb_sample_cfg(const vb_dit_desc* d, const vb_dit_cond* cond2, const vb_sample_args* args, void* workspace,
                  int64_t workspace_bytes, void* stream);

/* sizeof() of the public structs as compiled (ABI self-check for bindings): 0 = vb_gemm_args, 1 = vb_attn_args,
 * 2 = vb_voc_desc, 3 = vb_vae_desc, 4 = vb_dit_desc, 5 = vb_dit_cond, 6 = vb_dit_io,
 * 7 = vb_dit_prep_desc, 8 = vb_sample_args */
int vb_sizeof(int which);
/* number of kernels this library has launched in this process */
long long vb_launch_count(void);

/* Optional per-launch timing: when enabled every entry point brackets its kernel with CUDA events
 * on the launching stream; vb_prof_report() synchronises and returns total ms / launch count per
 * kernel family (0 gemm, 1 norm, 2 softmax, 3 attn_small, 4 router, 5 moe_sort, 6 elementwise,
 * 7 layout, 8 groupnorm, 9 conv_post). Not usable during CUDA-graph capture. */
/* haloed small-C conv path inside vb_gemm (C_in <= 64, N <= 64, fp16): 0 off, 1 on */
int vb_set_conv_halo(int mode);
/* experiments on vb_gemm_chain (tools/bench_chain.py): bit 0 ignore dependencies (WRONG results), bit 1 no
 * publish fences (WRONG), bit 2 m-fastest tile order; 0 = product behaviour */
int vb_set_chain_debug(int mode);
/* 2-CTA (tcgen05 cta_group::2, cluster of 2) GEMM path: 0 off, 1 heuristic, 2 whenever legal */
int vb_set_gemm2(int mode);
/* tuning hook: force tile width `bn` (and the 2-CTA path if cta2) for every vb_gemm with this
 * (N, K, grouped) signature; N <= 0 clears the table. Used by tools/tune_insitu.py only. */
int vb_set_gemm_override(int N, int K, int grouped, int bn, int cta2);
/* debug: the next `max_calls` vb_gemm launches (1-CTA kernel) each write a record of 256 CTAs x 8
 * uint64 %globaltimer stamps (entry, prologue done, first operands landed, first tile's MMAs
 * issued, its accumulator ready, its epilogue done, kernel exit) into consecutive 16 KB records of
 * this device buffer; NULL disables. See tools/trace_gemm.py */
int vb_set_gemm_trace(void* device_buf, int max_calls);
/* debug: writes %globaltimer (ns) to *slot when the stream reaches this point */
int vb_timestamp(uint64_t* slot, void* stream);
/* programmatic dependent launch between this library's kernels (default on) */
int vb_set_pdl(int on);
#define VB_PROF_KINDS 10
int vb_prof_enable(int on);
int vb_prof_report(double* ms_by_kind, long long* count_by_kind, int nkinds);

/* ------------------------------------------------------------------------------------------
 * Row-wise ops of the DiT step (fp32 in, fp16 GEMM operand and/or fp32 out)
 * ------------------------------------------------------------------------------------------ */
/* kind 0: RMSNorm (flag_large_dit_moe.py:52-77); kind 1: LayerNorm (vocal2music_moe.py:274-276,
 * flag_large_dit_moe.py:156). y = norm(x)[*w][+b]; then adaLN modulate y*(1+scale[bi])+shift[bi]
 * (flag_large_dit_moe.py:80-81) with bi = row / rows_per_batch, vectors at stride mod_ld. */
int vb_norm_mod(int kind, const float* x, const float* w, const float* b, const float* shift,
                const float* scale, int64_t mod_ld, int rows_per_batch, float eps, void* out_fp16,
                float* out_f32, int rows, int D, void* stream);
/* P[r,:n] = softmax(S[r,:n]) as fp16, P[r,n:ldp] = 0 (SDPA softmax, flag_large_dit_moe.py:381;
 * autoencoder1d.py:265) */
int vb_softmax_rows(const float* S, int64_t lds, void* P_fp16, int64_t ldp, int64_t rows, int n,
                    void* stream);
/* attention against a short K/V (L <= 128 keys), fp32 K,V [B,L,*] with row pitch kv_ld; q fp16 with element
 * strides (batch, token, head). out[b*T+t, h*hd+d] = base + tanh(gate[h]) * ctx (gate/base
 * optional): text cross-attention flag_large_dit_moe.py:388-402, MoE MHA vocal2music_moe.py:119 */
int vb_attn_smallkv(const void* q_fp16, int64_t q_sb, int64_t q_st, int64_t q_sh, const float* K,
                    const float* V, int64_t kv_ld, int B, int T, int H, int hd, int L, float scale,
                    const float* gate, const float* base, void* out_fp16, float* out_f32,
                    void* stream);
/* [cos(t f_i), sin(t f_i)] (flag_large_dit_moe.py:109-127); freqs = host table of f_i */
int vb_timestep_embed(const int64_t* t, const float* freqs, float* out_f32, void* out_fp16, int B,
                      int dim, void* stream);

/* out[b,:] = mean_l x[b,l,:] (caption pooling, vocal2music_moe.py:410-412) */
int vb_mean_tokens(const float* x, int B, int L, int D, float* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Band-MoE routing (vocal2music_moe.py:117-185)
 * ------------------------------------------------------------------------------------------ */
/* out[r,e] = bias[e] + <x[r,:], W[e,:]>, fp32 FMA, E <= 16 (gating networks :132,141-142) */
int vb_linear_small(const float* x, int64_t ldx, const float* W, const float* bias, float* out,
                    int rows, int D, int E, void* stream);
/* Gumbel-softmax router: e0 [N,2], e1/e2 [N,E] are the Exp(1) draws the reference obtains from
 * torch.empty_like(logits).exponential_() (:82). Emits hard top-1 indices and the combined
 * weights scale_g[n] = ((1-p)+p) * group_mask_g[n]. */
int vb_router(const float* cap_feat, const void* cap_feat_fp16, const float* Wc, const float* bc, const float* ac_logits,
              int ac_ld, const float* hl_logits, int hl_ld, const float* e0, const float* e1, const float* e2, float tau,
              int N, int T, int D, int E, int32_t* idx_c, int32_t* idx_a, float* scale_c,
              float* scale_a, float* cap_logits_out, int32_t* bucket_count, float* usage,
              int32_t* slot_c, int32_t* slot_a, const float* cap_partial, int n_heads, void* stream);
/* Features may be fp32 (cap_feat) or fp16 (cap_feat_fp16); or, with cap_partial != NULL, the caption
 * gate logits are bc + the sum over heads (in head order) of cap_partial[n][h][E] written by a
 * VB_EPI_ROUTE GEMM and cap_feat / Wc are not read. With bucket_count != NULL the router also
 * claims an (unordered) slot per (token, group) in its expert bucket with atomics; vb_moe_finalize
 * turns the counts into 128-row-padded bucket starts / tile table / lb loss (and re-zeroes them),
 * vb_moe_scatter writes row_map/row_scale and gathers the activation rows into bucket order. */
/* Step-invariant operands of a VB_EPI_ROUTE GEMM from the MoE cross-attention's fp32 k|v projection
 * kv [B*L, 2*H*hd] (L <= 96): bias_s [B, H, 96] = <k_h[j], bq_h> (-1e30 for j >= L) and
 * vg [B, H, 96, 4] = <v_h[j], capg_w[e, head h]> (capg_w [4, H*hd] = W_gate W_out_proj). */
int vb_route_prepare(const float* kv, int B, int L, int H, int hd, const float* bq, const float* capg_w,
                     float* bias_s, float* vg, void* stream);
int vb_moe_finalize(int32_t* bucket_count, float* usage, int E, int32_t* bucket_start, int32_t* row_map,
                    float* row_scale, int32_t* tile_group, int32_t* m_tiles_used, int max_rows,
                    float* lb_out, void* stream);
/* finalize + scatter in one launch; also zeroes next_count/next_usage (the next block's counters) */
int vb_moe_dispatch(const int32_t* idx_c, const int32_t* idx_a, const int32_t* slot_c, const int32_t* slot_a,
                    const float* scale_c, const float* scale_a, const int32_t* bucket_count, const float* usage,
                    int32_t* next_count, float* next_usage, int N, int E, int row_bytes, const void* x, void* xs,
                    int32_t* row_map, float* row_scale, int32_t* tile_group, int32_t* m_tiles_used,
                    int max_rows, float* lb_out, void* stream);
int vb_moe_scatter(const int32_t* idx_c, const int32_t* idx_a, const int32_t* slot_c, const int32_t* slot_a,
                   const float* scale_c, const float* scale_a, const int32_t* bucket_start, int N, int E,
                   int row_bytes, const void* x, void* xs, int32_t* row_map, float* row_scale,
                   void* stream);
/* All router noise of one sampling run in one launch, bit-identical to the reference's sequence of
 * torch `exponential_()` calls (PyTorch Philox4x32-10 stream: seed, offset advancing by 4 per call).
 * Layout: e0 [steps, depth, 2N, 2], e1/e2 [steps, depth, 2N, E], rows [0,N) = cond, [N,2N) = uncond;
 * call order (step, branch, block, {group, caption, acoustic}). */
int vb_router_noise(uint64_t seed, uint64_t offset, int N, int E, int depth, int steps, float* e0,
                    float* e1, float* e2, void* stream);
/* counting sort of the 2N (token, group) pairs into 128-row tiles per expert bucket */
int vb_moe_sort(const int32_t* idx_c, const int32_t* idx_a, const float* scale_c,
                const float* scale_a, int N, int E, int32_t* row_map, float* row_scale,
                int32_t* tile_group, int32_t* m_tiles_used, int max_rows, float* lb_out,
                void* stream);
int vb_gather_rows(const void* x_fp16, const int32_t* row_map, const int32_t* m_tiles_used,
                   int n_tokens, int D, void* out_fp16, int max_rows, void* stream);

/* ------------------------------------------------------------------------------------------
 * Elementwise / layout / GroupNorm / vocoder tail
 * ------------------------------------------------------------------------------------------ */
/* out = act(alpha*a + beta*b), rows x D fp32 in, each row written `rep` times (rep=2: nearest x2
 * upsample, autoencoder1d.py:288) */
int vb_ew(const float* a, const float* b, float alpha, float beta, int act, float slope,
          float* out_f32, void* out_fp16, int64_t rows, int D, int rep, void* stream);
/* [B,C,T] fp32 -> channels-last fp16 [B,T,Cp] (zero padded), times scale */
int vb_nct_to_ntc(const float* in, float scale, void* out, int B, int C, int T, int Cp, int out_is_f32,
                  void* stream);
/* e_u == NULL: x[b,c,t] = e_c[b,t,c]; else CFG + Euler (cfm1_audio.py:160):
 * x[b,c,t] += dt * (e_u + s*(e_c - e_u)).
 * x_cl != NULL (C <= 32): the new x is also written channels-last to x_cl[r][b][t][0..C) for r < x_cl_rep
 * ([rep][B][T][32], fp16 or fp32): the next DiT evaluation's stem operand for each CFG branch. */
int vb_ntc_to_nct(const float* e_c, const float* e_u, int64_t ld, float s, float dt, float* x, int B,
                  int C, int T, void* x_cl, int x_cl_is_f32, int x_cl_rep, void* stream);
int vb_embed_rows(const int64_t* idx, const float* table, int V, int D, void* out, int64_t rows,
                  int out_is_f32, void* stream);
/* AvgPool1d(2) of one or two branches, summed (vocal2music_moe.py:343,349,392) */
int vb_pool2_sum(const float* a, const float* b, int B, int Tin, int Tout, int D, void* out_fp16,
                 float* out_f32, void* stream);
/* GroupNorm over channels-last [B,T,C] (autoencoder1d.py:165-166) */
/* scratch: B*G*VB_GN_SPLIT*2 doubles (partial sums of the split-T first pass) */
#define VB_GN_SPLIT 8
int vb_gn_stats(const float* x, int B, int T, int C, int G, float eps, float* mean, float* rstd,
                void* scratch, void* stream);
int vb_gn_apply(const float* x, const float* mean, const float* rstd, const float* w,
                const float* bias, int B, int T, int C, int G, int act, void* out_fp16, void* stream);
/* GroupNorm + activation -> fp16 in ONE launch (a cluster of 8 CTAs per (batch, group) exchanges fp64 partial sums
 * through distributed shared memory; x is read once). vb_gn_fused_fits: 1 if the shape is supported (C/G a multiple
 * of 4, T >= 8, the slab of T/8 rows x C/G channels fits in shared memory); otherwise use vb_gn_stats + vb_gn_apply. */
int vb_gn_fused_fits(int T, int C, int G);
int vb_gn_fused(const float* x, const float* w, const float* bias, int B, int T, int C, int G, float eps, int act,
                void* out_fp16, void* stream);
/* HiFi-GAN conv_post + tanh (vocoder/hifigan/modules/hifigan.py:139-141); w is [k][C] fp32 */
int vb_conv_post(const void* x_fp16, const float* w, float bias, int B, int T, int C, int k,
                 float* wav, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VERSBAND_H_ */
