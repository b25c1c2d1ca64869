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
vb_gemm_args conv_args(const void* a, int B, int T, int Cin, const void* w, int Cout, int k, int dil, const float* bias) {
  vb_gemm_args g;
  memset(&g, 0, sizeof(g));
  const int kp = (Cin + 63) / 64 * 64;
  g.a = a;
  g.a_dims[0] = Cin; g.a_dims[1] = T; g.a_dims[2] = B; g.a_dims[3] = 1;
  g.a_strides[0] = 1; g.a_strides[1] = Cin; g.a_strides[2] = (int64_t)T * Cin; g.a_strides[3] = (int64_t)B * T * Cin + 8;
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
