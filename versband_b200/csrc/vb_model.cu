// Model-level entry points of the C-ABI: the launch sequence of a whole network issued
// from C++, so that a host in any language can run the path with plain pointers — no Python in the loop.
//   vb_voc_forward : HiFi-GAN generator (reference vocoder/hifigan/modules/hifigan.py:126-143), channels-last
//                    fp16 activations in the activated representation (see versband_b200/vocoder.py).
// The caller owns every buffer: packed weights (device pointers in the descriptor), input, output and ONE
// workspace whose size vb_voc_workspace_bytes reports. No allocation, no hidden state.
#include <cmath>
#include <cstring>
#include "vb_internal.h"

namespace {

struct Bump {  // carve the caller's workspace
  uint8_t* base;
  int64_t off, cap;
  void* take(int64_t bytes) {
    off = (off + 255) & ~int64_t(255);
    void* p = base ? base + off : nullptr;
    off += bytes;
    return p;
  }
};

// Conv1d (k taps, dilation d, "same" padding) over channels-last fp16 [B, T, Cin] with a packed weight
// [Cout, k * kp] — the argument block vocoder.py:_conv / ops.gemm_args build
vb_gemm_args conv_args(const void* a, int B, int T, int Cin, const void* w, int Cout, int k, int dil, const float* bias,
                       int pitch = 0) {
  vb_gemm_args g;
  memset(&g, 0, sizeof(g));
  const int kp = (Cin + 63) / 64 * 64;
  if (!pitch) pitch = Cin;  // row pitch of A in elements (>= Cin: channel-padded buffers)
  g.a = a;
  g.a_dims[0] = Cin; g.a_dims[1] = T; g.a_dims[2] = B; g.a_dims[3] = 1;
  g.a_strides[0] = 1; g.a_strides[1] = pitch; g.a_strides[2] = (int64_t)T * pitch; g.a_strides[3] = ((int64_t)B * T * pitch + 7) / 8 * 8 + 8;
  g.b = w;
  g.b_dims[0] = (int64_t)k * kp; g.b_dims[1] = Cout; g.b_dims[2] = 1;
  g.b_strides[0] = 1; g.b_strides[1] = (int64_t)k * kp; g.b_strides[2] = (int64_t)Cout * k * kp + 8;
  g.M = T; g.N = Cout; g.K = k * kp;
  g.G1 = B; g.G2 = 1;
  g.kpt = kp / 64; g.dil = dil; g.pad = (k * dil - dil) / 2;
  g.a_c2_g1 = 1;
  g.alpha = 1.f;
  g.bias = bias;
  g.gate_rows = 1;
  g.row_mul = 1; g.rows_out = T;
  g.out_ld = Cout; g.out_g1 = (int64_t)T * Cout;
  g.outb_ld = Cout; g.outb_g1 = (int64_t)T * Cout;
  g.out_scale = 1.f;
  g.res16_neg = 1.f;
  return g;
}

}  // namespace

using namespace vb;

static int voc_run(const vb_voc_desc* d, const float* mel, int B, int T, float* wav, void* ws, int64_t ws_bytes,
                   void* stream, int64_t* need) {
  if (!d || B <= 0 || T <= 0) return set_error(VB_ERR_INVALID, "vb_voc_forward: bad arguments");
  if (d->n_stages < 1 || d->n_stages > VB_VOC_MAX_STAGES || d->n_kernels < 1 || d->n_kernels > 3 || d->n_mel % 8)
    return set_error(VB_ERR_INVALID, "vb_voc_forward: unsupported generator (1..%d stages, 1..3 resblock kernels, n_mel %% 8 == 0)",
                     VB_VOC_MAX_STAGES);
  const bool dry = need != nullptr;
  Bump bump{static_cast<uint8_t*>(dry ? nullptr : ws), 0, ws_bytes};
  const float slope = d->lrelu, inv = 1.f / d->lrelu;
  int rc;
#define RUN(call) do { if (!dry) { rc = (call); if (rc) return rc; } } while (0)
  // mel [B, n_mel, T] fp32 -> channels-last fp16
  void* mel_cl = bump.take((int64_t)B * T * d->n_mel * 2);
  RUN(vb_nct_to_ntc(mel, 1.f, mel_cl, B, d->n_mel, T, d->n_mel, 0, stream));
  int C = d->c0;
  void* xb = bump.take((int64_t)B * T * C * 2);  // lrelu(conv_pre(mel)), hifigan.py:128-130
  {
    vb_gemm_args g = conv_args(mel_cl, B, T, d->n_mel, d->pre_w, C, d->pre_k, 1, d->pre_b);
    g.out_fp16 = xb; g.act_out = VB_ACT_LRELU; g.act_out_slope = slope;
    RUN(vb_gemm(&g, stream));
  }
  for (int i = 0; i < d->n_stages; ++i) {
    // ---- ConvTranspose1d as u polyphase GEMMs in one launch (hifigan.py:131)
    const int u = d->rates[i], k = d->up_k[i], Cout = d->up_cout[i];
    const int Q = (k + u - 1) / u, kp = (C + 63) / 64 * 64, p = (k - u) / 2;
    const int Tout = (T - 1) * u - 2 * p + k;
    void* xl = bump.take((int64_t)B * Tout * Cout * 2);
    {
      vb_gemm_args g;
      memset(&g, 0, sizeof(g));
      g.a = xb;
      g.a_dims[0] = C; g.a_dims[1] = T; g.a_dims[2] = B; g.a_dims[3] = 1;
      g.a_strides[0] = 1; g.a_strides[1] = C; g.a_strides[2] = (int64_t)T * C; g.a_strides[3] = (int64_t)B * T * C + 8;
      g.b = d->up_w[i];
      g.b_dims[0] = (int64_t)Q * kp; g.b_dims[1] = Cout; g.b_dims[2] = u;
      g.b_strides[0] = 1; g.b_strides[1] = (int64_t)Q * kp; g.b_strides[2] = (int64_t)Cout * Q * kp;
      g.M = T + Q - 1; g.N = Cout; g.K = Q * kp;
      g.G1 = u; g.G2 = B;
      g.kpt = kp / 64; g.dil = 1; g.pad = Q - 1;
      g.a_c2_g2 = 1; g.b_c2_g1 = 1;
      g.alpha = 1.f; g.bias = d->up_b[i]; g.gate_rows = 1;
      g.row_mul = u; g.row_off = -p; g.row_off_g1 = 1; g.rows_out = Tout;
      g.out_fp16 = xl; g.outb_ld = Cout; g.outb_g2 = (int64_t)Tout * Cout;
      g.act_out = VB_ACT_LRELU; g.act_out_slope = slope; g.out_scale = 1.f; g.res16_neg = 1.f;
      RUN(vb_gemm(&g, stream));
    }
    T = Tout; C = Cout;
    // ---- MRF: mean of the ResBlock1 chains, all starting from x (hifigan.py:132-138)
    const int64_t act = (int64_t)B * T * C * 2;
    void* nxt = bump.take(act);
    void* t1 = bump.take(act);
    void* rl2[2] = {bump.take(act), bump.take(act)};
    void* rout[3] = {nullptr, nullptr, nullptr};
    for (int j = 0; j < d->n_kernels; ++j) rout[j] = bump.take(act);
    for (int j = 0; j < d->n_kernels; ++j) {
      const int kk = d->rb_k[j], nd = d->n_dil[j];
      const void* rl = xl;
      for (int c = 0; c < nd; ++c) {
        const int dd = d->rb_d[j][c];
        void* dst = c == nd - 1 ? rout[j] : rl2[c & 1];
        const void* w1 = d->rb_w1[i][j][c];
        const void* w2 = d->rb_w2[i][j][c];
        const float* b1 = d->rb_b1[i][j][c];
        const float* b2 = d->rb_b2[i][j][c];
        if ((C == 32 || C == 64) && vb_resblock_pair_fits(C, kk, dd)) {
          RUN(vb_resblock_pair(rl, dst, w1, w2, b1, b2, B, T, C, kk, dd, slope, stream));
        } else {
          vb_gemm_args g1 = conv_args(rl, B, T, C, w1, C, kk, dd, b1);
          g1.out_fp16 = t1; g1.act_out = VB_ACT_LRELU; g1.act_out_slope = slope;
          RUN(vb_gemm(&g1, stream));
          vb_gemm_args g2 = conv_args(t1, B, T, C, w2, C, kk, 1, b2);
          g2.res16 = rl; g2.res16_neg = inv;
          g2.out_fp16 = dst; g2.act_out = VB_ACT_LRELU; g2.act_out_slope = slope;
          RUN(vb_gemm(&g2, stream));
        }
        rl = dst;
      }
    }
    const bool last = i == d->n_stages - 1;
    RUN(vb_mrf_combine(rout[0], rout[1], rout[2], inv, 1.f / d->n_kernels, last ? 0.01f : slope, nxt, (int64_t)B * T * C, stream));
    xb = nxt;
  }
  RUN(vb_conv_post(xb, d->post_w, d->post_b, B, T, C, d->post_k, wav, stream));
#undef RUN
  if (dry) { *need = (bump.off + 255) & ~int64_t(255); return VB_OK; }
  if (bump.off > ws_bytes) return set_error(VB_ERR_INVALID, "vb_voc_forward: workspace too small (%lld bytes needed)", (long long)bump.off);
  return VB_OK;
}

extern "C" int64_t vb_voc_workspace_bytes(const vb_voc_desc* d, int B, int T) {
  int64_t need = 0;
  if (voc_run(d, nullptr, B, T, nullptr, nullptr, 0, nullptr, &need)) return -1;
  return need;
}

extern "C" int vb_voc_forward(const vb_voc_desc* d, const float* mel, int B, int T, float* wav, void* ws,
                              int64_t ws_bytes, void* stream) {
  if (!mel || !wav || !ws) return set_error(VB_ERR_INVALID, "vb_voc_forward: null pointer");
  const int64_t need = vb_voc_workspace_bytes(d, B, T);
  if (need < 0) return VB_ERR_INVALID;
  if (need > ws_bytes) return set_error(VB_ERR_INVALID, "vb_voc_forward: workspace too small (%lld bytes needed)", (long long)need);
  return voc_run(d, mel, B, T, wav, ws, ws_bytes, stream, nullptr);
}

// ------------------------------------------------------------------------------------------------------------
// AutoencoderKL 1-D decoder (reference ldm/models/autoencoder1d.py:480-513): the launch sequence of
// versband_b200/vae.py:VAEEngine.decode issued from C++. Activations fp32 channels-last [B, T, C]; every conv's A
// operand is the fp16 GroupNorm(+swish) output.
// ------------------------------------------------------------------------------------------------------------
namespace {

struct Scratch {  // per-layer scratch carved from the caller's workspace; reset between layers, high-water mark kept
  uint8_t* base;
  int64_t off, peak;
  void reset() { off = 0; }
  void* take(int64_t bytes) {
    off = (off + 255) & ~int64_t(255);
    void* p = base ? base + off : nullptr;
    off += bytes;
    if (off > peak) peak = off;
    return p;
  }
};

inline int64_t pad8(int64_t v) { return (v + 7) / 8 * 8; }

vb_gemm_args plain_args() {
  vb_gemm_args g;
  memset(&g, 0, sizeof(g));
  g.alpha = 1.f; g.gate_rows = 1; g.row_mul = 1; g.out_scale = 1.f; g.res16_neg = 1.f;
  return g;
}

struct VaeCtx {
  const vb_vae_desc* d;
  int B;
  bool dry;
  void* stream;
  Scratch sc;
  float *mean, *rstd;
  void* gn_scratch;
  int rc;

  template <class F> void run(F&& f) { if (!dry && !rc) rc = f(); }

  // conv over channels-last fp16 a [B, T, pitch] (first Cin channels) -> fp32 out (+ residual) and/or fp16 out
  void conv(const void* a, int T, int Cin, int pitch, const void* w, const float* bias, int Cout, int k, float* out_f32,
            const float* residual, void* out_fp16, int out16_pitch) {
    vb_gemm_args g = conv_args(a, B, T, Cin, w, Cout, k, 1, bias, pitch);
    g.residual = residual;
    g.out_f32 = out_f32;
    g.out_fp16 = out_fp16;
    g.outb_ld = out_fp16 ? out16_pitch : 0;
    g.outb_g1 = out_fp16 ? (int64_t)T * out16_pitch : 0;
    run([&] { return vb_gemm(&g, stream); });
  }
  // GroupNorm (+ optional swish) of fp32 x [B, T, C] -> fp16 scratch buffer
  void* gn(const float* x, int T, int C, const float* w, const float* b, int act) {
    void* out = sc.take((int64_t)B * T * C * 2);
    if (vb_gn_fused_fits(T, C, d->groups)) {  // one cluster launch: statistics through DSMEM, x read once
      run([&] { return vb_gn_fused(x, w, b, B, T, C, d->groups, d->eps, act, out, stream); });
      return out;
    }
    run([&] { return vb_gn_stats(x, B, T, C, d->groups, d->eps, mean, rstd, gn_scratch, stream); });
    run([&] { return vb_gn_apply(x, mean, rstd, w, b, B, T, C, d->groups, act, out, stream); });
    return out;
  }
  // ResnetBlock1D (autoencoder1d.py:211-231): x + conv2(swish(gn2(conv1(swish(gn1(x)))))), 1x1 shortcut if Cin != Cout
  void resblock(const vb_vae_op& o, const float* x, float* out, int T) {
    const int Ci = o.cin, Co = o.cout;
    void* a = gn(x, T, Ci, o.n1_w, o.n1_b, VB_ACT_SILU);
    float* h = static_cast<float*>(sc.take((int64_t)B * T * Co * 4));
    conv(a, T, Ci, Ci, o.w1, o.b1, Co, o.k1, h, nullptr, nullptr, 0);
    void* a2 = gn(h, T, Co, o.n2_w, o.n2_b, VB_ACT_SILU);
    const float* res = x;
    if (o.ws) {
      void* xb = sc.take((int64_t)B * T * Ci * 2);
      run([&] { return vb_ew(x, nullptr, 1.f, 1.f, VB_ACT_NONE, 0.f, nullptr, xb, (int64_t)B * T, Ci, 1, stream); });
      float* r = static_cast<float*>(sc.take((int64_t)B * T * Co * 4));
      conv(xb, T, Ci, Ci, o.ws, o.bs, Co, o.ks, r, nullptr, nullptr, 0);
      res = r;
    }
    conv(a2, T, Co, Co, o.w2, o.b2, Co, o.k2, out, res, nullptr, 0);
  }
  // AttnBlock1D (autoencoder1d.py:253-274): single-head attention over time, 1x1 q|k|v and proj_out convs
  void attn(const vb_vae_op& o, const float* x, float* out, int T) {
    const int C = o.cin, Tp = (T + 7) / 8 * 8;
    void* a = gn(x, T, C, o.n1_w, o.n1_b, VB_ACT_NONE);
    void* q = sc.take((int64_t)B * T * C * 2);
    void* k = sc.take((int64_t)B * T * C * 2);
    void* vt = sc.take((int64_t)B * C * Tp * 2);
    float* S = static_cast<float*>(sc.take((int64_t)B * T * Tp * 4));
    void* P = sc.take((int64_t)B * T * Tp * 2);
    void* ob = sc.take((int64_t)B * T * C * 2);
    {
      vb_gemm_args g = plain_args();  // q|k|v = a W^T + b, stored per section (V transposed)
      const int64_t rows = (int64_t)B * T;
      g.a = a;
      g.a_dims[0] = C; g.a_dims[1] = rows; g.a_dims[2] = 1; g.a_dims[3] = 1;
      g.a_strides[0] = 1; g.a_strides[1] = C; g.a_strides[2] = pad8(rows * C) + 8; g.a_strides[3] = g.a_strides[2];
      g.b = o.w1;
      g.b_dims[0] = C; g.b_dims[1] = 3 * C; g.b_dims[2] = 1;
      g.b_strides[0] = 1; g.b_strides[1] = C; g.b_strides[2] = pad8((int64_t)3 * C * C) + 8;
      g.M = (int)rows; g.N = 3 * C; g.K = (C + 63) / 64 * 64; g.G1 = 1; g.G2 = 1;
      g.rows_out = g.M;
      g.mode = VB_EPI_QKV_ROPE; g.bias = o.b1;
      g.T = T; g.heads = 1; g.head_dim = C; g.Tp = Tp;
      g.q_out = q; g.k_out = k; g.vt_out = vt;
      run([&] { return vb_gemm(&g, stream); });
    }
    {
      vb_gemm_args g = plain_args();  // S = q k^T / sqrt(C), batched over B
      g.a = q;
      g.a_dims[0] = C; g.a_dims[1] = T; g.a_dims[2] = B; g.a_dims[3] = 1;
      g.a_strides[0] = 1; g.a_strides[1] = C; g.a_strides[2] = (int64_t)T * C; g.a_strides[3] = pad8((int64_t)B * T * C) + 8;
      g.b = k;
      g.b_dims[0] = C; g.b_dims[1] = T; g.b_dims[2] = B;
      g.b_strides[0] = 1; g.b_strides[1] = C; g.b_strides[2] = (int64_t)T * C;
      g.M = T; g.N = T; g.K = (C + 63) / 64 * 64; g.G1 = B; g.G2 = 1;
      g.a_c2_g1 = 1; g.b_c2_g1 = 1;
      g.rows_out = T;
      g.alpha = (float)pow((double)C, -0.5);
      g.out_f32 = S; g.out_ld = Tp; g.out_g1 = (int64_t)T * Tp;
      run([&] { return vb_gemm(&g, stream); });
    }
    run([&] { return vb_softmax_rows(S, Tp, P, Tp, (int64_t)B * T, T, stream); });
    {
      vb_gemm_args g = plain_args();  // o = P v, batched over B
      g.a = P;
      g.a_dims[0] = T; g.a_dims[1] = T; g.a_dims[2] = B; g.a_dims[3] = 1;
      g.a_strides[0] = 1; g.a_strides[1] = Tp; g.a_strides[2] = (int64_t)T * Tp; g.a_strides[3] = pad8((int64_t)B * T * Tp) + 8;
      g.b = vt;
      g.b_dims[0] = T; g.b_dims[1] = C; g.b_dims[2] = B;
      g.b_strides[0] = 1; g.b_strides[1] = Tp; g.b_strides[2] = (int64_t)C * Tp;
      g.M = T; g.N = C; g.K = (T + 63) / 64 * 64; g.G1 = B; g.G2 = 1;
      g.a_c2_g1 = 1; g.b_c2_g1 = 1;
      g.rows_out = T;
      g.out_fp16 = ob; g.outb_ld = C; g.outb_g1 = (int64_t)T * C;
      run([&] { return vb_gemm(&g, stream); });
    }
    conv(ob, T, C, C, o.w2, o.b2, C, o.k2, out, x, nullptr, 0);
  }
};

// channel count and length after every layer: the largest fp32 activation (the size of one ping-pong slot)
int64_t vae_slot_bytes(const vb_vae_desc* d, int B, int T) {
  int64_t m = (int64_t)B * T * d->c0 * 4;
  for (int i = 0; i < d->n_ops; ++i) {
    const vb_vae_op& o = d->ops[i];
    if (o.kind == VB_VAE_UP) T *= 2;
    const int64_t b = (int64_t)B * T * o.cout * 4;
    if (b > m) m = b;
  }
  const int64_t mel = (int64_t)B * T * d->out_ch * 4;
  return ((m > mel ? m : mel) + 255) & ~int64_t(255);
}

}  // namespace

static int vae_run(const vb_vae_desc* d, const float* z, float scale, int B, int T, float* mel, void* ws, int64_t ws_bytes,
                   void* stream, int64_t* need) {
  if (!d || B <= 0 || T <= 0) return set_error(VB_ERR_INVALID, "vb_vae_decode: bad arguments");
  if (d->n_ops < 0 || d->n_ops > VB_VAE_MAX_OPS || d->groups <= 0 || d->c0 % d->groups)
    return set_error(VB_ERR_INVALID, "vb_vae_decode: bad descriptor");
  const bool dry = need != nullptr;
  uint8_t* base = static_cast<uint8_t*>(dry ? nullptr : ws);
  const int64_t slot = vae_slot_bytes(d, B, T);
  const int zp = (d->zc + 7) / 8 * 8, zqp = (d->zq_c + 7) / 8 * 8;
  // fixed area: z operands, GroupNorm statistics, two ping-pong activation slots; then per-layer scratch
  Bump fix{base, 0, ws_bytes};
  void* z_cl = fix.take((int64_t)B * T * zp * 2);
  void* zq = fix.take((int64_t)B * T * zqp * 2);
  float* mean = static_cast<float*>(fix.take((int64_t)B * d->groups * 4));
  float* rstd = static_cast<float*>(fix.take((int64_t)B * d->groups * 4));
  void* gn_scratch = fix.take((int64_t)B * d->groups * VB_GN_SPLIT * 2 * 8);
  float* slots[2] = {static_cast<float*>(fix.take(slot)), static_cast<float*>(fix.take(slot))};
  const int64_t fixed = (fix.off + 255) & ~int64_t(255);
  VaeCtx c{d, B, dry, stream, Scratch{base ? base + fixed : nullptr, 0, 0}, mean, rstd, gn_scratch, 0};

  c.run([&] { return vb_nct_to_ntc(z, scale, z_cl, B, d->zc, T, zp, 0, stream); });
  c.conv(z_cl, T, d->zc, zp, d->pq_w, d->pq_b, d->zq_c, d->pq_k, nullptr, nullptr, zq, zqp);  // post_quant_conv
  int cur = 0, C = d->c0;
  c.conv(zq, T, d->zq_c, zqp, d->in_w, d->in_b, C, d->in_k, slots[cur], nullptr, nullptr, 0);   // decoder.conv_in
  for (int i = 0; i < d->n_ops && !c.rc; ++i) {
    const vb_vae_op& o = d->ops[i];
    if (o.cin != C || o.cout % d->groups || o.cout % 8)
      return set_error(VB_ERR_INVALID, "vb_vae_decode: layer %d expects %d input channels, got %d", i, o.cin, C);
    c.sc.reset();
    float* in = slots[cur];
    float* out = slots[cur ^ 1];
    if (o.kind == VB_VAE_RES) {
      c.resblock(o, in, out, T);
    } else if (o.kind == VB_VAE_ATTN) {
      c.attn(o, in, out, T);
    } else if (o.kind == VB_VAE_UP) {  // nearest x2 + conv (autoencoder1d.py:287-291)
      void* up = c.sc.take((int64_t)B * 2 * T * C * 2);
      c.run([&] { return vb_ew(in, nullptr, 1.f, 1.f, VB_ACT_NONE, 0.f, nullptr, up, (int64_t)B * T, C, 2, stream); });
      T *= 2;
      c.conv(up, T, C, C, o.w1, o.b1, o.cout, o.k1, out, nullptr, nullptr, 0);
    } else {
      return set_error(VB_ERR_INVALID, "vb_vae_decode: unknown layer kind %d", o.kind);
    }
    C = o.cout;
    cur ^= 1;
  }
  c.sc.reset();
  void* a = c.gn(slots[cur], T, C, d->no_w, d->no_b, VB_ACT_SILU);
  float* mel_cl = slots[cur ^ 1];
  c.conv(a, T, C, C, d->out_w, d->out_b, d->out_ch, d->out_k, mel_cl, nullptr, nullptr, 0);
  c.run([&] { return vb_ntc_to_nct(mel_cl, nullptr, d->out_ch, 0.f, 0.f, mel, B, d->out_ch, T, nullptr, 0, 0, stream); });
  if (c.rc) return c.rc;
  const int64_t total = fixed + ((c.sc.peak + 255) & ~int64_t(255));
  if (dry) { *need = total; return VB_OK; }
  return VB_OK;
}

extern "C" int64_t vb_vae_workspace_bytes(const vb_vae_desc* d, int B, int T) {
  int64_t need = 0;
  if (vae_run(d, nullptr, 1.f, B, T, nullptr, nullptr, 0, nullptr, &need)) return -1;
  return need;
}

extern "C" int vb_vae_decode(const vb_vae_desc* d, const float* z, float scale, int B, int T, float* mel, void* ws,
                             int64_t ws_bytes, void* stream) {
  if (!z || !mel || !ws) return set_error(VB_ERR_INVALID, "vb_vae_decode: null pointer");
  const int64_t need = vb_vae_workspace_bytes(d, B, T);
  if (need < 0) return VB_ERR_INVALID;
  if (need > ws_bytes) return set_error(VB_ERR_INVALID, "vb_vae_decode: workspace too small (%lld bytes needed)", (long long)need);
  return vae_run(d, z, scale, B, T, mel, ws, ws_bytes, stream, nullptr);
}

// AutoencoderKL 1-D encoder (reference ldm/models/autoencoder1d.py:49-53,384-409: Encoder1D.forward then quant_conv):
// mel fp32 [B, in_ch, T] -> moments fp32 [B, 2 * embed_dim, T / 2^n_down] (mean | logvar); the launch sequence of
// versband_b200/vae.py:VAEEngine.encode_moments issued from C++. The SAME descriptor type describes the encoder:
// zc = in_ch, in_w = encoder.conv_in, ops = resblocks / attention / VB_VAE_DOWN, no_* = encoder.norm_out,
// out_w = encoder.conv_out (out_ch = 2 * z_channels), pq_w = quant_conv applied LAST (zq_c = 2 * embed_dim).
static int vae_enc_run(const vb_vae_desc* d, const float* x, int B, int T, float* moments, void* ws, int64_t ws_bytes, void* stream,
                       int64_t* need) {
  if (!d || B <= 0 || T <= 0) return set_error(VB_ERR_INVALID, "vb_vae_encode: bad arguments");
  if (d->n_ops < 0 || d->n_ops > VB_VAE_MAX_OPS || d->groups <= 0 || d->c0 % d->groups)
    return set_error(VB_ERR_INVALID, "vb_vae_encode: bad descriptor");
  const bool dry = need != nullptr;
  uint8_t* base = static_cast<uint8_t*>(dry ? nullptr : ws);
  int64_t slot = (int64_t)B * T * d->c0 * 4;
  for (int i = 0; i < d->n_ops; ++i) {  // lengths only shrink: the largest activation is at full length
    const int64_t b = (int64_t)B * T * d->ops[i].cout * 4;
    if (b > slot) slot = b;
  }
  slot = (slot + 255) & ~int64_t(255);
  const int xp = (d->zc + 7) / 8 * 8, ocp = (d->out_ch + 7) / 8 * 8;
  Bump fix{base, 0, ws_bytes};
  void* x_cl = fix.take((int64_t)B * T * xp * 2);
  void* hz = fix.take((int64_t)B * T * ocp * 2);
  float* mean = static_cast<float*>(fix.take((int64_t)B * d->groups * 4));
  float* rstd = static_cast<float*>(fix.take((int64_t)B * d->groups * 4));
  void* gn_scratch = fix.take((int64_t)B * d->groups * VB_GN_SPLIT * 2 * 8);
  float* slots[2] = {static_cast<float*>(fix.take(slot)), static_cast<float*>(fix.take(slot))};
  const int64_t fixed = (fix.off + 255) & ~int64_t(255);
  VaeCtx c{d, B, dry, stream, Scratch{base ? base + fixed : nullptr, 0, 0}, mean, rstd, gn_scratch, 0};

  c.run([&] { return vb_nct_to_ntc(x, 1.f, x_cl, B, d->zc, T, xp, 0, stream); });
  int cur = 0, C = d->c0;
  c.conv(x_cl, T, d->zc, xp, d->in_w, d->in_b, C, d->in_k, slots[cur], nullptr, nullptr, 0);  // encoder.conv_in
  for (int i = 0; i < d->n_ops && !c.rc; ++i) {
    const vb_vae_op& o = d->ops[i];
    if (o.cin != C || o.cout % d->groups || o.cout % 8)
      return set_error(VB_ERR_INVALID, "vb_vae_encode: layer %d expects %d input channels, got %d", i, o.cin, C);
    c.sc.reset();
    float* in = slots[cur];
    float* out = slots[cur ^ 1];
    if (o.kind == VB_VAE_RES) {
      c.resblock(o, in, out, T);
    } else if (o.kind == VB_VAE_ATTN) {
      c.attn(o, in, out, T);
    } else if (o.kind == VB_VAE_DOWN) {  // pad (0, 1) + conv k3 stride 2 (autoencoder1d.py:306-313) on row pairs [T/2, 2C]
      if (T % 2 || (2 * C) % 64) return set_error(VB_ERR_INVALID, "vb_vae_encode: stride-2 downsample expects an even length and 2C %% 64 == 0");
      void* hb = c.sc.take((int64_t)B * T * C * 2);
      c.run([&] { return vb_ew(in, nullptr, 1.f, 1.f, VB_ACT_NONE, 0.f, nullptr, hb, (int64_t)B * T, C, 1, stream); });
      T /= 2;
      c.conv(hb, T, 2 * C, 2 * C, o.w1, o.b1, o.cout, 2, out, nullptr, nullptr, 0);
    } else {
      return set_error(VB_ERR_INVALID, "vb_vae_encode: unknown layer kind %d", o.kind);
    }
    C = o.cout;
    cur ^= 1;
  }
  c.sc.reset();
  void* a = c.gn(slots[cur], T, C, d->no_w, d->no_b, VB_ACT_SILU);
  c.conv(a, T, C, C, d->out_w, d->out_b, d->out_ch, d->out_k, nullptr, nullptr, hz, ocp);       // encoder.conv_out
  float* mom_cl = slots[cur ^ 1];
  c.conv(hz, T, d->out_ch, ocp, d->pq_w, d->pq_b, d->zq_c, d->pq_k, mom_cl, nullptr, nullptr, 0);  // quant_conv
  c.run([&] { return vb_ntc_to_nct(mom_cl, nullptr, d->zq_c, 0.f, 0.f, moments, B, d->zq_c, T, nullptr, 0, 0, stream); });
  if (c.rc) return c.rc;
  if (dry) *need = fixed + ((c.sc.peak + 255) & ~int64_t(255));
  return VB_OK;
}

extern "C" int64_t vb_vae_encode_workspace_bytes(const vb_vae_desc* d, int B, int T) {
  int64_t need = 0;
  if (vae_enc_run(d, nullptr, B, T, nullptr, nullptr, 0, nullptr, &need)) return -1;
  return need;
}

extern "C" int vb_vae_encode(const vb_vae_desc* d, const float* x, int B, int T, float* moments, void* ws, int64_t ws_bytes,
                             void* stream) {
  if (!x || !moments || !ws) return set_error(VB_ERR_INVALID, "vb_vae_encode: null pointer");
  const int64_t need = vb_vae_encode_workspace_bytes(d, B, T);
  if (need < 0) return VB_ERR_INVALID;
  if (need > ws_bytes) return set_error(VB_ERR_INVALID, "vb_vae_encode: workspace too small (%lld bytes needed)", (long long)need);
  return vae_enc_run(d, x, B, T, moments, ws, ws_bytes, stream, nullptr);
}

// ------------------------------------------------------------------------------------------------------------
// One evaluation of the AccompBand DiT (reference TxtFlagLargeDiT.forward, vocal2music_moe.py:375-434, with the
// step-invariant part precomputed: vb_dit_cond): the launch sequence of versband_b200/dit.py:DiTEngine.forward
// issued from C++. 12 launches per block (norm+modulate, q|k|v+RoPE, fused attention, wo+gate+residual,
// norm+modulate, caption-gate GEMM, router, dispatch, 4 expert GEMMs) + stem + final layer.
// ------------------------------------------------------------------------------------------------------------
namespace {

// vb_gemm operand views, innermost dimension first; size-1 / broadcast dimensions get any 16-byte-multiple stride
// (their coordinate is always 0) — the rule of ops.py:_rev_desc
void fill_degenerate(int64_t* dims, int64_t* strides, int rank) {
  int64_t ext = 1;
  for (int i = 0; i < rank; ++i) {
    if (i > 0 && (dims[i] == 1 || strides[i] == 0)) strides[i] = pad8(ext) > 8 ? pad8(ext) : 8;
    const int64_t e = dims[i] * (strides[i] > 1 ? strides[i] : 1);
    if (e > ext) ext = e;
  }
}
void set_a(vb_gemm_args& g, const void* p, int64_t d0, int64_t d1, int64_t s1, int64_t d2 = 1, int64_t s2 = 0, int64_t d3 = 1,
           int64_t s3 = 0) {
  g.a = p;
  g.a_dims[0] = d0; g.a_dims[1] = d1; g.a_dims[2] = d2; g.a_dims[3] = d3;
  g.a_strides[0] = 1; g.a_strides[1] = s1; g.a_strides[2] = s2; g.a_strides[3] = s3;
  fill_degenerate(g.a_dims, g.a_strides, 4);
}
void set_b(vb_gemm_args& g, const void* p, int64_t d0, int64_t d1, int64_t s1, int64_t d2 = 1, int64_t s2 = 0) {
  g.b = p;
  g.b_dims[0] = d0; g.b_dims[1] = d1; g.b_dims[2] = d2;
  g.b_strides[0] = 1; g.b_strides[1] = s1; g.b_strides[2] = s2;
  fill_degenerate(g.b_dims, g.b_strides, 3);
}

struct DitCtx {
  const vb_dit_desc* d;
  void* stream;
  bool dry;
  int rc;
  bool f32;  // TF32 mode: GEMM operands are fp32 tensors
  int kq;    // elements per 128-byte k-block
  template <class F> void run(F&& f) { if (!dry && !rc) rc = f(); }
  int64_t es() const { return f32 ? 4 : 2; }
  vb_gemm_args args(int M, int N, int K) const {
    vb_gemm_args g = plain_args();
    g.operand_f32 = f32;
    g.M = M; g.N = N; g.K = (K + kq - 1) / kq * kq;
    g.G1 = 1; g.G2 = 1;
    g.rows_out = M;
    return g;
  }
  // the operand-output slot of a GEMM / row op: fp16, or fp32 in TF32 mode
  void out_op(vb_gemm_args& g, void* p, int64_t ld) const { g.out_fp16 = p; g.outb_ld = ld; g.outb_f32 = f32; }
  void gemm(const vb_gemm_args& g) { run([&] { return vb_gemm(&g, stream); }); }
  void norm_mod(int kind, const float* x, const float* w, const float* shift, const float* scale, int64_t mod_ld,
                int rows_per_batch, float eps, void* out_op_, int rows, int D) {
    run([&] {
      return vb_norm_mod(kind, x, w, nullptr, shift, scale, mod_ld, rows_per_batch, eps, f32 ? nullptr : out_op_,
                         f32 ? static_cast<float*>(out_op_) : nullptr, rows, D, stream);
    });
  }
};

}  // namespace

// timestep embedding MLP, adaLN vectors of every block + final layer, group-gate logits for R independent rows
// (vocal2music_moe.py:406,416,242,290; DiTEngine.time_conditioning)
static void dit_time_cond(DitCtx& x, int R, const int64_t* t, const float* cap_emb, void* tsin, void* th, float* temb, void* ada,
                          float* mod_out, float* hl_out) {
  const vb_dit_desc* d = x.d;
  const int D = d->D;
  const int64_t mld = d->mod_ld;
  void* stream = x.stream;
  x.run([&] { return vb_timestep_embed(t, d->t_freqs, x.f32 ? static_cast<float*>(tsin) : nullptr, x.f32 ? nullptr : tsin, R, 256, stream); });
  { vb_gemm_args g = x.args(R, D, 256); set_a(g, tsin, 256, R, 256); set_b(g, d->t_w0, 256, D, 256);
    g.bias = d->t_b0; g.act = VB_ACT_SILU; x.out_op(g, th, D); x.gemm(g); }
  { vb_gemm_args g = x.args(R, D, D); set_a(g, th, D, R, D); set_b(g, d->t_w2, D, D, D);
    g.bias = d->t_b2; g.out_f32 = temb; g.out_ld = D; x.gemm(g); }
  x.run([&] { return vb_ew(temb, cap_emb, 1.f, 1.f, VB_ACT_SILU, 0.f, x.f32 ? static_cast<float*>(ada) : nullptr, x.f32 ? nullptr : ada, R, D, 1, stream); });
  { vb_gemm_args g = x.args(R, (int)mld, D); set_a(g, ada, D, R, D); set_b(g, d->ada_w, D, mld, D);
    g.bias = d->ada_b; g.out_f32 = mod_out; g.out_ld = mld; x.gemm(g); }
  x.run([&] { return vb_linear_small(temb, D, d->hl_w, d->hl_b, hl_out, R, D, 2 * d->depth, stream); });
}

static int dit_run(const vb_dit_desc* d, const vb_dit_cond* c, const vb_dit_io* io, void* ws, int64_t ws_bytes, void* stream,
                   int B, int T, int64_t* need) {
  const bool dry = need != nullptr;
  if (!d || B <= 0 || T <= 0) return set_error(VB_ERR_INVALID, "vb_dit_forward: bad arguments");
  const int D = d->D, H = d->H, E = d->E, Fh = d->Fh, depth = d->depth, C = d->C;
  if (depth < 1 || depth > VB_DIT_MAX_DEPTH || H <= 0 || D % H || E != 4 || D % E || C > 32 || T > d->max_len)
    return set_error(VB_ERR_INVALID, "vb_dit_forward: unsupported model (depth 1..%d, 4 experts, in_channels <= 32, T <= max_len)",
                     VB_DIT_MAX_DEPTH);
  const int hd = D / H, band = D / E, LK = 96;
  const int N = B * T, Tp = (T + 7) / 8 * 8;
  const int max_rows = ((2 * N + 127) / 128 + 2 * E) * 128;
  const int64_t mld = d->mod_ld;
  DitCtx x{d, stream, dry, 0, d->operand_f32 != 0, d->operand_f32 ? 32 : 64};
  const int64_t es = x.es();
  Bump bump{static_cast<uint8_t*>(dry ? nullptr : ws), 0, ws_bytes};
  auto f32buf = [&](int64_t n) { return static_cast<float*>(bump.take(n * 4)); };
  auto i32buf = [&](int64_t n) { return static_cast<int32_t*>(bump.take(n * 4)); };
  // cross-call state first (block k's dispatch re-zeroes block k+1's counters): zero at first use
  int32_t* bucket_count = i32buf((int64_t)depth * 2 * E);
  float* usage = f32buf((int64_t)depth * 2 * E);
  void* x_cl_own = bump.take((int64_t)B * T * 32 * es);
  float* xs_ = f32buf((int64_t)N * D);
  float* hbuf = f32buf((int64_t)N * D);
  void* xa = bump.take((int64_t)N * D * es);
  void* q = bump.take((int64_t)N * D * 2);
  void* k = bump.take((int64_t)N * D * 2);
  void* vt = bump.take((int64_t)B * H * hd * Tp * 2);
  void* oa = bump.take((int64_t)N * D * es);
  void* xm = bump.take((int64_t)N * D * es);
  float* cap_part = f32buf((int64_t)N * H * E);
  int32_t *idx_c = i32buf(N), *idx_a = i32buf(N), *slot_c = i32buf(N), *slot_a = i32buf(N);
  float *scale_c = f32buf(N), *scale_a = f32buf(N);
  int32_t* bucket_start = i32buf(2 * E);
  int32_t* row_map = i32buf(max_rows);
  float* row_scale = f32buf(max_rows);
  int32_t* tile_group = i32buf(max_rows / 128);
  int32_t* m_tiles = i32buf(1);
  void* xsb = bump.take((int64_t)max_rows * D * es);
  void* hs = bump.take((int64_t)max_rows * Fh * es);
  void* y2 = bump.take((int64_t)2 * N * D * es);
  void* hf = bump.take((int64_t)E * N * Fh * es);
  void* xf = bump.take((int64_t)N * D * es);
  void* tsin = bump.take((int64_t)B * 256 * es);
  void* th = bump.take((int64_t)B * D * es);
  float* temb = f32buf((int64_t)B * D);
  void* ada = bump.take((int64_t)B * D * es);
  float* mod_own = f32buf((int64_t)B * mld);
  float* hl_own = f32buf((int64_t)B * 2 * depth);
  if (dry) { *need = (bump.off + 255) & ~int64_t(255); return VB_OK; }
  if (bump.off > ws_bytes) return set_error(VB_ERR_INVALID, "vb_dit_forward: workspace too small (%lld bytes needed)", (long long)bump.off);
  if (!c || !io || !io->v || !io->lb || c->B != B || c->T != T || c->L > LK || (!io->x_cl_ready && (!io->x || io->Bx <= 0 || B % io->Bx)))
    return set_error(VB_ERR_INVALID, "vb_dit_forward: bad conditioning / io block");
  void* x_cl = io->x_cl ? io->x_cl : x_cl_own;
  if (io->x_cl_ready && !io->x_cl) return set_error(VB_ERR_INVALID, "vb_dit_forward: x_cl_ready without x_cl");

  // ---- stem: proj_in conv + bias + acoustic (vocal2music_moe.py:395,415); CFG batching: the same latent feeds
  // every group of Bx conditioning rows
  if (!io->x_cl_ready)
    for (int r = 0; r < B / io->Bx; ++r)
      x.run([&] {
        return vb_nct_to_ntc(io->x, 1.f, static_cast<uint8_t*>(x_cl) + (int64_t)r * io->Bx * T * 32 * es, io->Bx, C, T, 32, x.f32, stream);
      });
  {
    const int K = d->proj_in_k * x.kq;
    vb_gemm_args g = x.args(T, D, K);
    set_a(g, x_cl, C, T, 32, B, (int64_t)T * 32);
    set_b(g, d->proj_in_w, K, D, K);
    g.G1 = B; g.kpt = 1; g.dil = 1; g.pad = d->proj_in_k / 2; g.a_c2_g1 = 1;
    g.bias = d->proj_in_b; g.residual = c->acoustic;
    g.out_f32 = xs_; g.out_ld = D; g.out_g1 = (int64_t)T * D;
    x.gemm(g);
  }
  // ---- timestep embedding MLP, adaLN vectors of every block + final layer, group-gate logits (:406,416,242,290)
  const float *mod = io->mod, *hl = io->hl;
  if (!mod || !hl) {
    if (!io->t) return set_error(VB_ERR_INVALID, "vb_dit_forward: neither timesteps nor precomputed modulation given");
    dit_time_cond(x, B, io->t, c->cap_emb, tsin, th, temb, ada, mod_own, hl_own);
    mod = mod_own; hl = hl_own;
  }
  float* xcur = xs_;
  const float scale = (float)(1.0 / sqrt((double)hd));  // rounded once from double, like the Python host code
  for (int kb = 0; kb < depth && !x.rc; ++kb) {
    const vb_dit_block& b = d->blocks[kb];
    const float* m0 = mod + (int64_t)kb * 6 * D;
    // ---- attention branch (vocal2music_moe.py:245-250; flag_large_dit_moe.py:323-406)
    x.norm_mod(0, xcur, b.attn_norm, m0, m0 + D, mld, T, d->eps, xa, N, D);
    { vb_gemm_args g = x.args(N, 3 * D, D); set_a(g, xa, D, N, D); set_b(g, b.wqkv, D, 3 * D, D);
      g.mode = VB_EPI_QKV_ROPE; g.rope_cs = d->rope; g.T = T; g.heads = H; g.head_dim = hd; g.Tp = Tp;
      g.q_out = q; g.k_out = k; g.vt_out = vt; x.gemm(g); }
    {
      vb_attn_args a;
      memset(&a, 0, sizeof(a));
      a.q = q; a.q_st = hd; a.q_sh = (int64_t)T * hd; a.q_sb = (int64_t)H * T * hd;
      a.k_self = k; a.vt_self = vt; a.k_text = c->k_y[kb]; a.vt_text = c->vt_y[kb];
      a.gate = b.gate;
      a.out_fp16 = x.f32 ? nullptr : oa; a.out_f32 = x.f32 ? static_cast<float*>(oa) : nullptr;
      a.B = B; a.T = T; a.H = H; a.hd = hd; a.Tp = Tp; a.L = c->L; a.Lp = c->Lp; a.scale = scale;
      x.run([&] { return vb_attn_fused(&a, stream); });
    }
    { vb_gemm_args g = x.args(N, D, D); set_a(g, oa, D, N, D); set_b(g, b.wo, D, D, D);
      g.gate = m0 + 2 * D; g.gate_rows = T; g.gate_ld = mld; g.residual = xcur; g.out_f32 = hbuf; g.out_ld = D; x.gemm(g); }
    // ---- Band-MoE branch (vocal2music_moe.py:251-257, 117-185)
    x.norm_mod(0, hbuf, b.ffn_norm, m0 + 3 * D, m0 + 4 * D, mld, T, d->eps, xm, N, D);
    { vb_gemm_args g = x.args(T, H * LK, D);  // caption gate: per-head softmax(xm . WK^T + bias) . VG in the epilogue
      set_a(g, xm, D, T, D, B, (int64_t)T * D); set_b(g, c->wk_m[kb], D, H * LK, D, B, (int64_t)H * LK * D);
      g.G1 = B; g.a_c2_g1 = 1; g.b_c2_g1 = 1; g.mode = VB_EPI_ROUTE;
      g.bias = c->bs_m[kb]; g.bias_g1 = H * LK; g.route_vg = c->vg_m[kb]; g.route_vg_g1 = (int64_t)H * LK * E;
      g.out_f32 = cap_part; g.out_ld = H * E; x.gemm(g); }
    int32_t* bc = bucket_count + (int64_t)kb * 2 * E;
    float* us = usage + (int64_t)kb * 2 * E;
    x.run([&] {
      return vb_router(nullptr, nullptr, nullptr, b.capg_b, c->ac_logits + (int64_t)kb * E, depth * E, hl + 2 * kb, 2 * depth,
                       io->e0[kb], io->e1[kb], io->e2[kb], d->temperature, N, T, 0, E, idx_c, idx_a, scale_c, scale_a, nullptr,
                       bc, us, slot_c, slot_a, cap_part, H, stream);
    });
    if (depth > 1) {
      const int kn = (kb + 1) % depth;  // this block's dispatch zeroes the next block's counters
      x.run([&] {
        return vb_moe_dispatch(idx_c, idx_a, slot_c, slot_a, scale_c, scale_a, bc, us, bucket_count + (int64_t)kn * 2 * E,
                               usage + (int64_t)kn * 2 * E, N, E, (int)(D * es), xm, xsb, row_map, row_scale, tile_group, m_tiles,
                               max_rows, io->lb + kb, stream);
      });
    } else {
      x.run([&] { return vb_moe_finalize(bc, us, E, bucket_start, row_map, row_scale, tile_group, m_tiles, max_rows, io->lb + kb, stream); });
      x.run([&] {
        return vb_moe_scatter(idx_c, idx_a, slot_c, slot_a, scale_c, scale_a, bucket_start, N, E, (int)(D * es), xm, xsb, row_map,
                              row_scale, stream);
      });
    }
    // routed experts (w1|w3 -> SwiGLU -> w2, scattered back to token rows), then the frequency experts over their
    // bands (y = y_caption + y_acoustic summed in the accumulator: 2 "taps" N rows apart) + adaLN gate + residual
    { vb_gemm_args g = x.args(max_rows, 2 * Fh, D); set_a(g, xsb, D, max_rows, D);
      set_b(g, b.w13, D, 2 * Fh, D, 2 * E, (int64_t)2 * Fh * D);
      g.b_c2_tile = 1; g.tile_group = tile_group; g.m_tiles_used = m_tiles; g.swiglu = 1; x.out_op(g, hs, Fh); x.gemm(g); }
    { vb_gemm_args g = x.args(max_rows, D, Fh); set_a(g, hs, Fh, max_rows, Fh); set_b(g, b.w2, Fh, D, Fh, 2 * E, (int64_t)D * Fh);
      g.b_c2_tile = 1; g.tile_group = tile_group; g.m_tiles_used = m_tiles; g.row_map = row_map; g.rows_out = 2 * N;
      g.row_scale = row_scale; x.out_op(g, y2, D); x.gemm(g); }
    { vb_gemm_args g = x.args(N, 2 * Fh, 2 * band); set_a(g, y2, band, 2 * N, D, E, band);
      set_b(g, b.f13, 2 * band, 2 * Fh, 2 * band, E, (int64_t)2 * Fh * 2 * band);
      g.G1 = E; g.kpt = band / x.kq; g.dil = N; g.pad = 0; g.a_c2_g1 = 1; g.b_c2_g1 = 1; g.swiglu = 1;
      x.out_op(g, hf, Fh); g.outb_g1 = (int64_t)N * Fh; x.gemm(g); }
    { vb_gemm_args g = x.args(N, band, Fh); set_a(g, hf, Fh, N, Fh, E, (int64_t)N * Fh); set_b(g, b.f2, Fh, band, Fh, E, (int64_t)band * Fh);
      g.G1 = E; g.a_c2_g1 = 1; g.b_c2_g1 = 1;
      g.gate = m0 + 5 * D; g.gate_rows = T; g.gate_ld = mld; g.gate_g1 = band; g.residual = hbuf;
      g.out_f32 = xcur; g.out_ld = D; g.out_g1 = band; x.gemm(g); }
  }
  // ---- final layer (vocal2music_moe.py:268-291,431): all 32 zero-padded output columns are computed
  const float* f0 = mod + (int64_t)depth * 6 * D;
  x.norm_mod(1, xcur, nullptr, f0, f0 + D, mld, T, 1e-6f, xf, N, D);
  { vb_gemm_args g = x.args(N, 32, D); set_a(g, xf, D, N, D); set_b(g, d->final_w, D, 32, D);
    g.bias = d->final_b; g.out_f32 = io->v; g.out_ld = 32; x.gemm(g); }
  if (io->out_v) x.run([&] { return vb_ntc_to_nct(io->v, nullptr, 32, 0.f, 0.f, io->out_v, B, C, T, nullptr, 0, 0, stream); });
  return x.rc;
}

extern "C" int64_t vb_dit_workspace_bytes(const vb_dit_desc* d, int B, int T) {
  int64_t need = 0;
  if (dit_run(d, nullptr, nullptr, nullptr, 0, nullptr, B, T, &need)) return -1;
  return need;
}

extern "C" int vb_dit_forward(const vb_dit_desc* d, const vb_dit_cond* c, const vb_dit_io* io, void* ws, int64_t ws_bytes,
                              void* stream) {
  if (!d || !c || !io || !ws) return set_error(VB_ERR_INVALID, "vb_dit_forward: null pointer");
  return dit_run(d, c, io, ws, ws_bytes, stream, c->B, c->T, nullptr);
}

// ------------------------------------------------------------------------------------------------------------
// Step-invariant conditioning of the DiT (everything in TxtFlagLargeDiT.forward that does not depend on (x, t),
// vocal2music_moe.py:388-393,397-412 + the per-block K/V projections and the folded caption gate): the launch
// sequence of versband_b200/dit.py:DiTEngine.prepare / prepare_pair issued from C++. The acoustic stem is
// computed for the Bm clips once and replicated to every group of Bm conditioning rows (B = groups * Bm captions:
// the CFG pair is 2 groups).
// ------------------------------------------------------------------------------------------------------------
static int dit_prepare_run(const vb_dit_desc* d, const vb_dit_prep_desc* p, const int64_t* midi, const int64_t* beats, int Bm,
                           int Tm, const float* caption, int B, int L, int T, const vb_dit_cond* out, float* caption_out,
                           void* ws, int64_t ws_bytes, void* stream, int64_t* need) {
  const bool dry = need != nullptr;
  if (!d || !p || Bm <= 0 || B <= 0 || B % Bm || L <= 0 || L > 96 || T <= 0 || Tm < 2)
    return set_error(VB_ERR_INVALID, "vb_dit_prepare: bad arguments (caption length <= 96, B a multiple of the clip count)");
  const int D = d->D, H = d->H, E = d->E, depth = d->depth, hd = D / H, LK = 96;
  const int Tpool = Tm / 2, Lp = (L + 7) / 8 * 8;
  if (abs(T - Tpool) > 2) return set_error(VB_ERR_INVALID, "vb_dit_prepare: latent / conditioning length mismatch (%d vs %d)", T, Tpool);
  DitCtx x{d, stream, dry, 0, d->operand_f32 != 0, d->operand_f32 ? 32 : 64};
  const int64_t es = x.es();
  const int kpad = (hd + x.kq - 1) / x.kq * x.kq;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Bump bump{static_cast<uint8_t*>(dry ? nullptr : ws), 0, ws_bytes};
  void* pooled3 = bump.take((int64_t)Bm * Tpool * 3 * D * es);
  float* ac_tmp = static_cast<float*>(bump.take((int64_t)Bm * Tpool * D * 4));
  void* cap16 = bump.take((int64_t)B * L * p->ori_dim * es);
  void* h1 = bump.take((int64_t)B * L * D * es);
  float* h2 = static_cast<float*>(bump.take((int64_t)B * L * D * 4));
  float* cap = static_cast<float*>(bump.take((int64_t)B * L * D * 4));
  void* cap_h = bump.take((int64_t)B * L * D * 2);
  float* pool = static_cast<float*>(bump.take((int64_t)B * D * 4));
  void* pool_n = bump.take((int64_t)B * D * es);
  void* y16 = bump.take((int64_t)B * L * D * es);
  float* kvf = static_cast<float*>(bump.take((int64_t)B * L * 2 * D * 4));
  void* kvh = bump.take((int64_t)B * L * 2 * D * 2);
  if (dry) { *need = (bump.off + 255) & ~int64_t(255); return VB_OK; }
  if (bump.off > ws_bytes) return set_error(VB_ERR_INVALID, "vb_dit_prepare: workspace too small (%lld bytes needed)", (long long)bump.off);
  if (!midi || !beats || !caption || !out || out->B != B || out->T != T || out->L != L || out->Lp != Lp || !out->acoustic ||
      !out->ac_logits || !out->cap_emb)
    return set_error(VB_ERR_INVALID, "vb_dit_prepare: bad output block (B, T, L, Lp = ceil8(L) and the output pointers must be set)");
  float* acoustic = const_cast<float*>(out->acoustic);
  float* ac_logits = const_cast<float*>(out->ac_logits);
  auto cuda_ok = [&](cudaError_t e) { return e == cudaSuccess ? VB_OK : set_error(VB_ERR_CUDA, "vb_dit_prepare: %s", cudaGetErrorString(e)); };

  // ---- acoustic stem: conv(k) over embedding rows as exact table gathers + lrelu + AvgPool(2) + sum, then final_proj
  // on the split-precision operand (vocal2music_moe.py:336-350,388-393)
  x.run([&] {
    return vb_stem_gather(midi, beats, p->midi_tab, p->beats_tab, p->midi_proj_b, p->beats_proj_b, p->Vm, p->Vb, p->stem_k, Bm, Tm,
                          Tpool, D, 0.01f, pooled3, x.f32, stream);
  });
  float* ac0 = Tpool == T ? acoustic : ac_tmp;
  { vb_gemm_args g = x.args(Bm * Tpool, D, 3 * D); set_a(g, pooled3, 3 * D, (int64_t)Bm * Tpool, 3 * D); set_b(g, p->final_proj_w3, 3 * D, D, 3 * D);
    g.bias = p->final_proj_b; g.out_f32 = ac0; g.out_ld = D; x.gemm(g); }
  if (Tpool != T) {  // length fix-up (vocal2music_moe.py:397-401): repeat the last frame or crop
    const int keep = T < Tpool ? T : Tpool;
    x.run([&] {
      return cuda_ok(cudaMemcpy2DAsync(acoustic, (size_t)T * D * 4, ac_tmp, (size_t)Tpool * D * 4, (size_t)keep * D * 4, Bm,
                                       cudaMemcpyDeviceToDevice, st));
    });
    for (int r = Tpool; r < T; ++r)
      x.run([&] {
        return cuda_ok(cudaMemcpy2DAsync(acoustic + (int64_t)r * D, (size_t)T * D * 4, ac_tmp + (int64_t)(Tpool - 1) * D, (size_t)Tpool * D * 4,
                                         (size_t)D * 4, Bm, cudaMemcpyDeviceToDevice, st));
      });
  }
  x.run([&] { return vb_linear_small(acoustic, D, p->acg_w, p->acg_b, ac_logits, Bm * T, D, depth * E, stream); });
  for (int r = 1; r < B / Bm; ++r) {  // the other conditioning groups share the clips' acoustic stem
    x.run([&] { return cuda_ok(cudaMemcpyAsync(acoustic + (int64_t)r * Bm * T * D, acoustic, (size_t)Bm * T * D * 4, cudaMemcpyDeviceToDevice, st)); });
    x.run([&] {
      return cuda_ok(cudaMemcpyAsync(ac_logits + (int64_t)r * Bm * T * depth * E, ac_logits, (size_t)Bm * T * depth * E * 4,
                                     cudaMemcpyDeviceToDevice, st));
    });
  }
  // ---- caption embedder (c_embedder MLP + LayerNorm), pooled caption embedding (vocal2music_moe.py:408-412)
  const int R = B * L;
  x.run([&] { return vb_ew(caption, nullptr, 1.f, 1.f, VB_ACT_NONE, 0.f, x.f32 ? static_cast<float*>(cap16) : nullptr, x.f32 ? nullptr : cap16, R, p->ori_dim, 1, stream); });
  { vb_gemm_args g = x.args(R, D, p->ori_dim); set_a(g, cap16, p->ori_dim, R, p->ori_dim); set_b(g, p->c_w0, p->ori_dim, D, p->ori_dim);
    g.bias = p->c_b0; g.act = VB_ACT_GELU_ERF; x.out_op(g, h1, D); x.gemm(g); }
  { vb_gemm_args g = x.args(R, D, D); set_a(g, h1, D, R, D); set_b(g, p->c_w2, D, D, D); g.bias = p->c_b2; g.out_f32 = h2; g.out_ld = D; x.gemm(g); }
  float* capf = caption_out ? caption_out : cap;
  x.run([&] { return vb_norm_mod(1, h2, p->c_ln_w, p->c_ln_b, nullptr, nullptr, 0, R, 1e-5f, x.f32 ? nullptr : cap_h, capf, R, D, stream); });
  const void* cap_op = x.f32 ? static_cast<const void*>(capf) : cap_h;  // fp32 operands: the fp32 caption is the operand itself
  x.run([&] { return vb_mean_tokens(capf, B, L, D, pool, stream); });
  x.run([&] { return vb_norm_mod(1, pool, p->cap_ln_w, p->cap_ln_b, nullptr, nullptr, 0, B, 1e-5f, x.f32 ? nullptr : pool_n, x.f32 ? static_cast<float*>(pool_n) : nullptr, B, D, stream); });
  { vb_gemm_args g = x.args(B, D, D); set_a(g, pool_n, D, B, D); set_b(g, p->cap_w, D, D, D); g.bias = p->cap_b;
    g.out_f32 = const_cast<float*>(out->cap_emb); g.out_ld = D; x.gemm(g); }
  // ---- per block: K / V^T of the text cross-attention, and the MoE caption cross-attention folded into its gate
  for (int kb = 0; kb < depth && !x.rc; ++kb) {
    const vb_dit_prep_block& b = p->blocks[kb];
    void* k_y = const_cast<void*>(out->k_y[kb]);
    void* vt_y = const_cast<void*>(out->vt_y[kb]);
    void* wk = const_cast<void*>(out->wk_m[kb]);
    if (!k_y || !vt_y || !wk || !out->bs_m[kb] || !out->vg_m[kb]) return set_error(VB_ERR_INVALID, "vb_dit_prepare: output pointer of block %d missing", kb);
    x.norm_mod(0, capf, b.y_norm, nullptr, nullptr, 0, R, d->eps, y16, R, D);
    x.run([&] { return cuda_ok(cudaMemsetAsync(vt_y, 0, (size_t)B * H * hd * Lp * 2, st)); });  // key padding columns L..Lp
    { vb_gemm_args g = x.args(R, 2 * D, D); set_a(g, y16, D, R, D); set_b(g, b.wkvy, D, 2 * D, D);
      g.mode = VB_EPI_QKV_ROPE; g.qkv_sec0 = 1; g.T = L; g.heads = H; g.head_dim = hd; g.Tp = Lp; g.k_out = k_y; g.vt_out = vt_y; x.gemm(g); }
    { vb_gemm_args g = x.args(R, 2 * D, D); set_a(g, cap_op, D, R, D); set_b(g, b.mkv_w, D, 2 * D, D); g.bias = b.mkv_b;
      g.out_f32 = kvf; g.out_ld = 2 * D; if (!x.f32) { g.out_fp16 = kvh; g.outb_ld = 2 * D; } x.gemm(g); }
    x.run([&] { return cuda_ok(cudaMemsetAsync(wk, 0, (size_t)B * H * LK * D * es, st)); });  // key slots L..96 of every head
    { vb_gemm_args g = x.args(L, D, kpad);  // WK[b,h] = K_h (scale W_q,h): batched over (head, batch)
      set_a(g, x.f32 ? static_cast<const void*>(kvf) : kvh, hd, L, 2 * D, H, hd, B, (int64_t)L * 2 * D);
      set_b(g, b.mq_wT, kpad, D, kpad, H, (int64_t)D * kpad);
      g.G1 = H; g.G2 = B; g.a_c2_g1 = 1; g.a_c3_g2 = 1; g.b_c2_g1 = 1;
      x.out_op(g, wk, D); g.outb_g1 = (int64_t)LK * D; g.outb_g2 = (int64_t)H * LK * D; x.gemm(g); }
    x.run([&] {
      return vb_route_prepare(kvf, B, L, H, hd, b.mq_bs, b.capg_w, const_cast<float*>(out->bs_m[kb]), const_cast<float*>(out->vg_m[kb]), stream);
    });
  }
  return x.rc;
}

extern "C" int64_t vb_dit_prepare_workspace_bytes(const vb_dit_desc* d, const vb_dit_prep_desc* p, int Bm, int Tm, int B, int L, int T) {
  int64_t need = 0;
  if (dit_prepare_run(d, p, nullptr, nullptr, Bm, Tm, nullptr, B, L, T, nullptr, nullptr, nullptr, 0, nullptr, &need)) return -1;
  return need;
}

extern "C" int vb_dit_prepare(const vb_dit_desc* d, const vb_dit_prep_desc* p, const int64_t* midi, const int64_t* beats, int Bm, int Tm,
                              const float* caption, int B, int L, const vb_dit_cond* out, float* caption_out, void* ws,
                              int64_t ws_bytes, void* stream) {
  if (!d || !p || !out || !ws) return set_error(VB_ERR_INVALID, "vb_dit_prepare: null pointer");
  return dit_prepare_run(d, p, midi, beats, Bm, Tm, caption, B, L, out->T, out, caption_out, ws, ws_bytes, stream, nullptr);
}


// ------------------------------------------------------------------------------------------------------------
// The classifier-free-guidance flow sampler (reference ldm/models/diffusion/cfm1_audio_sampler.py:87-116,
// cfm1_audio.py:145-162 `Wrapper_cfg`, torchdyn fixed-step Euler): n_steps x { one DiT evaluation of the stacked
// (cond, uncond) batch ; x += dt * (v_u + w (v_c - v_u)) } — the launch sequence of versband_b200/sampler.py:
// FlowSampler._run issued from C++. The adaLN / gate vectors of all steps are computed first in one batch of
// n_steps * B rows (they depend only on t and the caption).
// ------------------------------------------------------------------------------------------------------------
static int sample_run(const vb_dit_desc* d, const vb_dit_cond* c, const vb_sample_args* a, int Bc, int T, int S, void* ws, int64_t ws_bytes,
                      void* stream, int64_t* need) {
  const bool dry = need != nullptr;
  if (!d || Bc <= 0 || T <= 0 || S <= 0) return set_error(VB_ERR_INVALID, "vb_sample_cfg: bad arguments");
  const int B = 2 * Bc, D = d->D, depth = d->depth, E = d->E, C = d->C;
  const int64_t N = (int64_t)B * T, mld = d->mod_ld;
  int64_t dit_bytes = 0;
  int rc = dit_run(d, nullptr, nullptr, nullptr, 0, nullptr, B, T, &dit_bytes);
  if (rc) return rc;
  DitCtx x{d, stream, dry, 0, d->operand_f32 != 0, d->operand_f32 ? 32 : 64};
  const int64_t es = x.es();
  Bump bump{static_cast<uint8_t*>(dry ? nullptr : ws), 0, ws_bytes};
  void* dit_ws = bump.take(dit_bytes);  // first: the caller's one-time zero fill covers the MoE counters
  void* x_cl = bump.take(N * 32 * es);
  float* v = static_cast<float*>(bump.take(N * 32 * 4));
  float* lb_own = static_cast<float*>(bump.take((int64_t)depth * 4));
  const int R = S * B;
  void* tsin = bump.take((int64_t)R * 256 * es);
  void* th = bump.take((int64_t)R * D * es);
  float* temb = static_cast<float*>(bump.take((int64_t)R * D * 4));
  float* cap_rep = static_cast<float*>(bump.take((int64_t)R * D * 4));
  void* ada = bump.take((int64_t)R * D * es);
  float* mod_all = static_cast<float*>(bump.take((int64_t)R * mld * 4));
  float* hl_all = static_cast<float*>(bump.take((int64_t)R * 2 * depth * 4));
  if (dry) { *need = (bump.off + 255) & ~int64_t(255); return VB_OK; }
  if (bump.off > ws_bytes) return set_error(VB_ERR_INVALID, "vb_sample_cfg: workspace too small (%lld bytes needed)", (long long)bump.off);
  if (!c || !a || c->B != B || c->T != T || !a->t_all || !a->dt || !a->e0 || !a->e1 || !a->e2 || !a->x)
    return set_error(VB_ERR_INVALID, "vb_sample_cfg: bad conditioning / argument block (the conditioning batch must hold 2 * Bc rows)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  for (int s = 0; s < S && !x.rc; ++s)
    x.run([&] {
      cudaError_t e = cudaMemcpyAsync(cap_rep + (int64_t)s * B * D, c->cap_emb, (size_t)B * D * 4, cudaMemcpyDeviceToDevice, st);
      return e == cudaSuccess ? VB_OK : set_error(VB_ERR_CUDA, "vb_sample_cfg: %s", cudaGetErrorString(e));
    });
  dit_time_cond(x, R, a->t_all, cap_rep, tsin, th, temb, ada, mod_all, hl_all);
  for (int s = 0; s < S && !x.rc; ++s) {
    vb_dit_io io;
    memset(&io, 0, sizeof(io));
    io.x = a->x; io.Bx = Bc; io.x_cl = x_cl; io.x_cl_ready = s > 0;
    io.mod = mod_all + (int64_t)s * B * mld; io.hl = hl_all + (int64_t)s * B * 2 * depth;
    for (int k = 0; k < depth; ++k) {
      const int64_t blk = (int64_t)s * depth + k;
      io.e0[k] = a->e0 + blk * N * 2; io.e1[k] = a->e1 + blk * N * E; io.e2[k] = a->e2 + blk * N * E;
    }
    io.v = v; io.lb = a->lb ? a->lb : lb_own;
    x.run([&] { return dit_run(d, c, &io, dit_ws, dit_bytes, stream, B, T, nullptr); });
    // CFG mix + Euler update; the same launch rewrites the channels-last operand copies of x for the next step
    x.run([&] { return vb_ntc_to_nct(v, v + (N / 2) * 32, 32, a->guidance, a->dt[s], a->x, Bc, C, T, x_cl, x.f32, 2, stream); });
  }
  return x.rc;
}

extern "C" int64_t vb_sample_workspace_bytes(const vb_dit_desc* d, int Bc, int T, int n_steps) {
  int64_t need = 0;
  if (sample_run(d, nullptr, nullptr, Bc, T, n_steps, nullptr, 0, nullptr, &need)) return -1;
  return need;
}

extern "C" int vb_sample_cfg(const vb_dit_desc* d, const vb_dit_cond* cond2, const vb_sample_args* a, void* ws, int64_t ws_bytes,
                             void* stream) {
  if (!d || !cond2 || !a || !ws) return set_error(VB_ERR_INVALID, "vb_sample_cfg: null pointer");
  return sample_run(d, cond2, a, a->Bc, cond2->T, a->n_steps, ws, ws_bytes, stream, nullptr);
}
