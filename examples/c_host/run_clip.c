/* A plain-C host for the AccompBand hot path: one clip from conditioning to waveform through the model-level
 * entry points of libversband_sm100.so (include/versband.h) — no Python, no torch.
 *
 *   run_clip <model.vbw> <case.bin> <out_wav.f32>
 *
 * model.vbw : packed weights + relocatable descriptors written by tools/export_weights.py
 * case.bin  : one test case written by tests/test_c_host_gpu.py:
 *             int32 T, Tm, L, ori_dim, n_steps, pad | uint64 seed, offset | float guidance, scale_factor |
 *             int64 t_steps[n_steps] | float dt[n_steps] | int64 midi[Tm] | int64 beats[Tm] |
 *             float captions[2][L][ori_dim] (caption, unconditional caption) | float x0[C][T]
 * out       : float wav[T_mel * hop]
 *
 * The sequence mirrors scripts/test_final.py:398-421 of the reference (gen_song): conditioning -> CFG sampler ->
 * VAE decode -> vocoder.  Build: gcc -O2 run_clip.c -I../../include -I$CUDA/include -L. -lversband_sm100 -lcudart
 */
#include <cuda_runtime_api.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "versband.h"

#define CK(call)                                                                       \
  do {                                                                                 \
    cudaError_t e_ = (call);                                                           \
    if (e_ != cudaSuccess) {                                                           \
      fprintf(stderr, "%s:%d: %s\n", __FILE__, __LINE__, cudaGetErrorString(e_));      \
      exit(2);                                                                         \
    }                                                                                  \
  } while (0)
#define VB(call)                                                                       \
  do {                                                                                 \
    if ((call) != 0) {                                                                 \
      fprintf(stderr, "%s:%d: %s\n", __FILE__, __LINE__, vb_last_error());             \
      exit(3);                                                                         \
    }                                                                                  \
  } while (0)

static void* slurp(const char* path, size_t* n) {
  FILE* f = fopen(path, "rb");
  if (!f) { perror(path); exit(1); }
  fseek(f, 0, SEEK_END);
  *n = (size_t)ftell(f);
  fseek(f, 0, SEEK_SET);
  void* p = malloc(*n);
  if (fread(p, 1, *n, f) != *n) { perror("fread"); exit(1); }
  fclose(f);
  return p;
}

static void* dmalloc(size_t bytes) {
  void* p;
  CK(cudaMalloc(&p, bytes ? bytes : 1));
  return p;
}

/* copy one descriptor out of the model file and turn its (1 + data offset) fields into device pointers */
static const uint8_t* load_struct(const uint8_t* p, const char* want, void* dst, size_t dst_bytes, const uint8_t* dev_base) {
  char name[17] = {0};
  int64_t bytes, n_reloc;
  memcpy(name, p, 16); p += 16;
  memcpy(&bytes, p, 8); p += 8;
  memcpy(&n_reloc, p, 8); p += 8;
  if (strcmp(name, want) || (size_t)bytes != dst_bytes) {
    fprintf(stderr, "model file: expected struct %s of %zu bytes, found %s of %lld (header / library mismatch)\n", want, dst_bytes, name, (long long)bytes);
    exit(1);
  }
  memcpy(dst, p, (size_t)bytes); p += bytes;
  for (int64_t i = 0; i < n_reloc; ++i) {
    int64_t off;
    uint64_t v;
    memcpy(&off, p + 8 * i, 8);
    memcpy(&v, (uint8_t*)dst + off, 8);
    v = (uint64_t)(uintptr_t)(dev_base + (v - 1));
    memcpy((uint8_t*)dst + off, &v, 8);
  }
  return p + 8 * n_reloc;
}

int main(int argc, char** argv) {
  if (argc != 4) { fprintf(stderr, "usage: %s model.vbw case.bin out.f32\n", argv[0]); return 1; }
  VB(vb_check_device());
  if (vb_sizeof(4) != (int)sizeof(vb_dit_desc) || vb_sizeof(7) != (int)sizeof(vb_dit_prep_desc) || vb_sizeof(3) != (int)sizeof(vb_vae_desc) ||
      vb_sizeof(2) != (int)sizeof(vb_voc_desc)) { fprintf(stderr, "versband.h does not match the library\n"); return 1; }
  cudaStream_t s;
  CK(cudaStreamCreate(&s));

  /* ---- model: data section to the device, descriptors relocated */
  size_t n;
  uint8_t* file = (uint8_t*)slurp(argv[1], &n);
  if (memcmp(file, "VBW1", 4)) { fprintf(stderr, "not a VBW1 file\n"); return 1; }
  int32_t n_structs; int64_t data_bytes;
  memcpy(&n_structs, file + 4, 4); memcpy(&data_bytes, file + 8, 8);
  uint8_t* dev_data = (uint8_t*)dmalloc((size_t)data_bytes);
  CK(cudaMemcpy(dev_data, file + n - (size_t)data_bytes, (size_t)data_bytes, cudaMemcpyHostToDevice));
  static vb_dit_desc dit; static vb_dit_prep_desc prep; static vb_vae_desc vae; static vb_voc_desc voc;
  const uint8_t* p = file + 16;
  p = load_struct(p, "dit", &dit, sizeof(dit), dev_data);
  p = load_struct(p, "prep", &prep, sizeof(prep), dev_data);
  p = load_struct(p, "vae", &vae, sizeof(vae), dev_data);
  p = load_struct(p, "voc", &voc, sizeof(voc), dev_data);
  (void)n_structs;

  /* ---- the case */
  size_t cn;
  uint8_t* cs = (uint8_t*)slurp(argv[2], &cn);
  int32_t hdr[6]; uint64_t seed_off[2]; float gs[2];
  memcpy(hdr, cs, 24); memcpy(seed_off, cs + 24, 16); memcpy(gs, cs + 40, 8);
  const int T = hdr[0], Tm = hdr[1], L = hdr[2], ori = hdr[3], S = hdr[4];
  const int D = dit.D, H = dit.H, E = dit.E, depth = dit.depth, C = dit.C, hd = D / H, LK = 96, Lp = (L + 7) / 8 * 8;
  const int es = dit.operand_f32 ? 4 : 2;
  const uint8_t* q = cs + 48;
  const int64_t* t_steps = (const int64_t*)q; q += 8 * (size_t)S;
  float* dt = (float*)malloc(4 * (size_t)S); memcpy(dt, q, 4 * (size_t)S); q += 4 * (size_t)S;
  const size_t midi_b = 8 * (size_t)Tm, cap_b = 4 * (size_t)2 * L * ori, x_b = 4 * (size_t)C * T;
  int64_t *midi = (int64_t*)dmalloc(midi_b), *beats = (int64_t*)dmalloc(midi_b);
  float *caps = (float*)dmalloc(cap_b), *x = (float*)dmalloc(x_b);
  CK(cudaMemcpy(midi, q, midi_b, cudaMemcpyHostToDevice)); q += midi_b;
  CK(cudaMemcpy(beats, q, midi_b, cudaMemcpyHostToDevice)); q += midi_b;
  CK(cudaMemcpy(caps, q, cap_b, cudaMemcpyHostToDevice)); q += cap_b;
  CK(cudaMemcpy(x, q, x_b, cudaMemcpyHostToDevice)); q += x_b;
  if ((size_t)(q - cs) != cn) { fprintf(stderr, "case file size mismatch\n"); return 1; }

  /* ---- step-invariant conditioning of the CFG pair (rows: caption, unconditional caption) */
  const int B = 2, N = B * T;
  vb_dit_cond cond;
  memset(&cond, 0, sizeof(cond));
  cond.B = B; cond.T = T; cond.L = L; cond.Lp = Lp;
  cond.acoustic = (float*)dmalloc(4 * (size_t)N * D);
  cond.ac_logits = (float*)dmalloc(4 * (size_t)N * depth * E);
  cond.cap_emb = (float*)dmalloc(4 * (size_t)B * D);
  for (int k = 0; k < depth; ++k) {
    cond.k_y[k] = dmalloc(2 * (size_t)B * H * L * hd);
    cond.vt_y[k] = dmalloc(2 * (size_t)B * H * hd * Lp);
    cond.wk_m[k] = dmalloc((size_t)es * B * H * LK * D);
    cond.bs_m[k] = (float*)dmalloc(4 * (size_t)B * H * LK);
    cond.vg_m[k] = (float*)dmalloc(4 * (size_t)B * H * LK * E);
  }
  int64_t wp = vb_dit_prepare_workspace_bytes(&dit, &prep, 1, Tm, B, L, T);
  if (wp < 0) { fprintf(stderr, "%s\n", vb_last_error()); return 3; }
  void* ws_prep = dmalloc((size_t)wp);
  VB(vb_dit_prepare(&dit, &prep, midi, beats, 1, Tm, caps, B, L, &cond, NULL, ws_prep, wp, s));

  /* ---- router noise (torch's exponential_() Philox stream) and the integer timesteps of every step and row */
  float *e0 = (float*)dmalloc(4 * (size_t)S * depth * N * 2), *e1 = (float*)dmalloc(4 * (size_t)S * depth * N * E),
        *e2 = (float*)dmalloc(4 * (size_t)S * depth * N * E);
  VB(vb_router_noise(seed_off[0], seed_off[1], T, E, depth, S, e0, e1, e2, s));
  int64_t* t_host = (int64_t*)malloc(8 * (size_t)S * B);
  for (int i = 0; i < S; ++i) for (int b = 0; b < B; ++b) t_host[i * B + b] = t_steps[i];
  int64_t* t_all = (int64_t*)dmalloc(8 * (size_t)S * B);
  CK(cudaMemcpy(t_all, t_host, 8 * (size_t)S * B, cudaMemcpyHostToDevice));

  /* ---- 24-step CFG Euler sampler */
  int64_t wsm = vb_sample_workspace_bytes(&dit, 1, T, S);
  if (wsm < 0) { fprintf(stderr, "%s\n", vb_last_error()); return 3; }
  void* ws_smp = dmalloc((size_t)wsm);
  CK(cudaMemsetAsync(ws_smp, 0, (size_t)wsm, s));
  vb_sample_args a;
  memset(&a, 0, sizeof(a));
  a.n_steps = S; a.Bc = 1; a.guidance = gs[0]; a.t_all = t_all; a.dt = dt; a.e0 = e0; a.e1 = e1; a.e2 = e2; a.x = x;
  VB(vb_sample_cfg(&dit, &cond, &a, ws_smp, wsm, s));

  /* ---- VAE decode (z / scale_factor) and vocoder */
  int n_up = 0, hop = 1;
  for (int i = 0; i < vae.n_ops; ++i) n_up += vae.ops[i].kind == VB_VAE_UP;
  for (int i = 0; i < voc.n_stages; ++i) hop *= voc.rates[i];
  const int T_mel = T << n_up;
  float *mel = (float*)dmalloc(4 * (size_t)vae.out_ch * T_mel), *wav = (float*)dmalloc(4 * (size_t)T_mel * hop);
  int64_t wv = vb_vae_workspace_bytes(&vae, 1, T), wh = vb_voc_workspace_bytes(&voc, 1, T_mel);
  if (wv < 0 || wh < 0) { fprintf(stderr, "%s\n", vb_last_error()); return 3; }
  void *ws_vae = dmalloc((size_t)wv), *ws_voc = dmalloc((size_t)wh);
  VB(vb_vae_decode(&vae, x, 1.f / gs[1], 1, T, mel, ws_vae, wv, s));
  VB(vb_voc_forward(&voc, mel, 1, T_mel, wav, ws_voc, wh, s));
  CK(cudaStreamSynchronize(s));

  float* out = (float*)malloc(4 * (size_t)T_mel * hop);
  CK(cudaMemcpy(out, wav, 4 * (size_t)T_mel * hop, cudaMemcpyDeviceToHost));
  FILE* f = fopen(argv[3], "wb");
  if (!f) { perror(argv[3]); return 1; }
  fwrite(out, 4, (size_t)T_mel * hop, f);
  fclose(f);
  printf("run_clip: %d latent frames -> %d samples, %lld kernels launched\n", T, T_mel * hop, vb_launch_count());
  return 0;
}
