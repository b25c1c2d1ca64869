// Model-level entry points of the C-ABI: the launch sequence of a whole network issued
// from C++, so that a host in any language can run the path with plain pointers — no Python in the loop.
//   vb_voc_forward : HiFi-GAN generator (reference vocoder/hifigan/modules/hifigan.py:126-143), channels-last
//                    fp16 activations in the activated representation (see versband_b200/vocoder.py).
// The caller owns every buffer: packed weights (device pointers in the descriptor), input, output and ONE
// workspace whose size vb_voc_workspace_bytes reports. No allocation, no hidden state.
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
      g.alpha = 1.f / sqrtf((float)C);
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
