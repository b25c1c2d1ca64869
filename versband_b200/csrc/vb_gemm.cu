// vb_gemm — host side: argument checks, tile selection, TMA descriptors and dispatch to the
// per-epilogue-class kernel instances (vb_gemm_kernels.cuh, instantiated in vb_gemm_ec*.cu).
// See include/versband.h for the contract.
#include <math.h>
#include <cstdio>
#include <cstdlib>
#include "vb_gemm_params.h"

namespace vb {

// ------------------------------------------------------------------ host side
static int encode_map(CUtensorMap* map, const void* base, int rank, const int64_t* dims,
                      const int64_t* strides, const uint32_t* box, bool tf32) {
  const int esz = tf32 ? 4 : 2;
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bx[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = static_cast<cuuint64_t>(dims[i]);
    bx[i] = box[i];
    estr[i] = 1;
    if (i > 0) {
      gstr[i - 1] = static_cast<cuuint64_t>(strides[i]) * esz;  // bytes
      if (gstr[i - 1] % 16 != 0 || gstr[i - 1] == 0)
        return set_error(VB_ERR_INVALID, "vb_gemm: tensor stride %lld of dim %d is not a positive multiple of 16 bytes",
                         (long long)strides[i], i);
    }
    if (dims[i] <= 0) return set_error(VB_ERR_INVALID, "vb_gemm: non-positive tensor dim");
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0)
    return set_error(VB_ERR_INVALID, "vb_gemm: operand base pointer not 16-byte aligned");
  auto fn = get_tensor_map_encoder();
  if (!fn) return set_error(VB_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  CUresult r = fn(map, tf32 ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank, const_cast<void*>(base), gdim, gstr,
                  bx, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(VB_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return VB_OK;
}

static int g_dbg_shift = 0;
static int g_conv_halo = 1;  // 0 off, 1 on
static int g_chain_dbg = 0;
static int g_gemm2_mode = 1;  // 0 = off, 1 = cost model, 2 = whenever legal
// in-situ tuning hook (tools/tune_insitu.py): force the tile shape of every vb_gemm whose (N, K, grouped)
// matches an entry. Not used by the product path.
struct TileOverride { int N, K, grouped, bn, cta2; };
static TileOverride g_overrides[16];
static int g_n_overrides = 0;
static unsigned long long* g_gemm_trace = nullptr;
static int g_gemm_trace_left = 0;

}  // namespace vb

extern "C" int vb_set_conv_halo(int mode) {
  vb::g_conv_halo = mode;
  return VB_OK;
}

extern "C" int vb_set_debug_shift(int rows) {
  vb::g_dbg_shift = rows;
  return VB_OK;
}

extern "C" int vb_set_gemm_override(int N, int K, int grouped, int bn, int cta2) {
  using namespace vb;
  if (N <= 0) { g_n_overrides = 0; return VB_OK; }
  for (int i = 0; i < g_n_overrides; ++i)
    if (g_overrides[i].N == N && g_overrides[i].K == K && g_overrides[i].grouped == grouped) {
      g_overrides[i].bn = bn; g_overrides[i].cta2 = cta2;
      return VB_OK;
    }
  if (g_n_overrides >= 16) return set_error(VB_ERR_INVALID, "vb_set_gemm_override: table full");
  g_overrides[g_n_overrides++] = TileOverride{N, K, grouped, bn, cta2};
  return VB_OK;
}

extern "C" int vb_set_gemm_trace(void* buf, int max_calls) {
  vb::g_gemm_trace = static_cast<unsigned long long*>(buf);
  vb::g_gemm_trace_left = buf ? max_calls : 0;
  return VB_OK;
}

extern "C" int vb_set_chain_debug(int mode) {
  vb::g_chain_dbg = mode;
  return VB_OK;
}

extern "C" int vb_set_gemm2(int mode) {
  vb::g_gemm2_mode = mode;
  return VB_OK;
}

namespace vb {
// Validation, tile selection (or `force_bn`, 1-CTA only), TMA descriptors and kernel parameter block of one
// GEMM problem. Shared by vb_gemm (one problem per launch) and vb_gemm_chain (several per launch).
int plan_gemm(const vb_gemm_args* a, int force_bn, GemmPlan* out) {
  if (!a || !a->a || !a->b) return set_error(VB_ERR_INVALID, "vb_gemm: null operand");
  const bool tf32 = a->operand_f32 != 0;
  const int bk = tf32 ? 32 : 64;  // elements per 128-byte k-block
  if (a->M <= 0 || a->N <= 0 || a->K <= 0 || a->K % bk != 0)
    return set_error(VB_ERR_INVALID, "vb_gemm: bad problem size M=%d N=%d K=%d (K must be a multiple of %d)",
                     a->M, a->N, a->K, bk);
  if (a->G1 <= 0 || a->G2 <= 0) return set_error(VB_ERR_INVALID, "vb_gemm: bad group grid");
  const int num_kb = a->K / bk;
  const int kpt = a->kpt > 0 ? a->kpt : num_kb;
  if (num_kb % kpt != 0) return set_error(VB_ERR_INVALID, "vb_gemm: K/64 not a multiple of kpt");
  if (a->a_strides[0] != 1 || a->b_strides[0] != 1)
    return set_error(VB_ERR_INVALID, "vb_gemm: reduction dim must be contiguous");

  // N tile
  int bn = force_bn > 0 ? force_bn : a->bn;
  const int m_tiles = (a->M + BM - 1) / BM;
  const int G = a->G1 * a->G2;
  // ---- tile selection -------------------------------------------------------------------
  // Cost model fitted to in-sequence device timings on B200 (tools/tune_insitu.py: every candidate
  // forced in turn inside one graph-replayed DiT evaluation): operand ingest is bounded per SM
  // (~64 B/clk) and chip-wide (L2 throughput, ~90 x 64 B/clk), every tile pays an epilogue
  // proportional to its width in 64-column rounds (two epilogue warps per TMEM quadrant) plus a
  // fixed pipeline fill. Units: 128-byte rows.
  bool use2 = false;
  {
    const bool user_bn = bn != 0;
    const int cands[6] = {32, 64, 96, 128, 192, 256};
    double best = 1e30;
    int best_bn = 32;
    const int n_cap = (a->N + 31) / 32 * 32;
    for (int i = 0; i < 6; ++i) {
      const int c = cands[i];
      if (user_bn && c != bn) continue;
      if (c > n_cap && c != 32) continue;
      const double tiles = (double)m_tiles * ((a->N + c - 1) / c) * G;
      const double waves = ceil(tiles / 148.0);
      const double u = (double)num_kb * (128 + c);
      const double cost = fmax(waves * u, tiles * u / 90.0) + waves * (40.0 * ((c + 63) / 64 * 64) + 1500.0);
      if (cost < best) { best = cost; best_bn = c; }
    }
    // 2-CTA pairs (cta_group::2): each CTA loads only half of the B tile. Non-grouped problems only.
    if (g_gemm2_mode > 0 && !user_bn && !force_bn && !a->tile_group && m_tiles >= 2 && a->N >= 128) {
      const int c2s[3] = {128, 192, 256};
      for (int i = 0; i < 3; ++i) {
        const int c = c2s[i];
        if (c > n_cap) continue;
        const double ctas = 2.0 * ((m_tiles + 1) / 2) * ((a->N + c - 1) / c) * G;
        const double waves = ceil(ctas / 148.0);
        const double u = (double)num_kb * (128 + c / 2);
        const double cost = fmax(waves * u, ctas * u / 90.0) + waves * (40.0 * ((c + 63) / 64 * 64) + 2100.0);
        if (cost < best || g_gemm2_mode == 2) {
          if (g_gemm2_mode == 2 && use2 && cost >= best) continue;
          best = cost; best_bn = c; use2 = true;
        }
      }
    }
    bn = best_bn;
    for (int i = 0; i < g_n_overrides; ++i)
      if (g_overrides[i].N == a->N && g_overrides[i].K == a->K && g_overrides[i].grouped == (a->tile_group ? 1 : 0) && !user_bn && !force_bn) {
        bn = g_overrides[i].bn;
        use2 = g_overrides[i].cta2 != 0 && !a->tile_group && m_tiles >= 2 && bn >= 128;
      }
  }
  {
    static const bool trace = getenv("VB_GEMM_TRACE") != nullptr;
    if (trace)
      fprintf(stderr, "vb_gemm M=%d N=%d K=%d G=%d grouped=%d tf32=%d kpt=%d -> bn=%d cta2=%d\n", a->M, a->N, a->K, G,
              a->tile_group ? 1 : 0, (int)tf32, kpt, bn, (int)use2);
  }
  if (a->mode == VB_EPI_ROUTE) {
    if (!a->route_vg || !a->bias || !a->out_f32 || a->N % 96 != 0 || a->bias_g1 % 4 != 0 || a->out_ld % 4 != 0 ||
        a->tile_group || a->G2 != 1 || (reinterpret_cast<uintptr_t>(a->bias) & 15) ||
        (reinterpret_cast<uintptr_t>(a->route_vg) & 15) || (reinterpret_cast<uintptr_t>(a->out_f32) & 15))
      return set_error(VB_ERR_INVALID, "vb_gemm: bad ROUTE arguments (N = heads*96, bias/route_vg/out_f32 16-byte aligned)");
    bn = 96;
    use2 = false;
  }
  if (a->mode == VB_EPI_QKV_ROPE) {
    if (!a->k_out || !a->vt_out || a->head_dim % 32 != 0 ||
        a->N != (3 - a->qkv_sec0) * a->heads * a->head_dim || a->qkv_sec0 < 0 || a->qkv_sec0 > 1 || (a->qkv_sec0 == 0 && !a->q_out) || a->T <= 0 || a->M % a->T != 0)
      return set_error(VB_ERR_INVALID, "vb_gemm: bad QKV_ROPE arguments");
    if ((bn == 96 || bn == 192) && (a->head_dim % 32) != 0) { bn = 64; use2 = false; }
  }
  if (a->swiglu && (a->N % 2 != 0))
    return set_error(VB_ERR_INVALID, "vb_gemm: swiglu needs an even N");

  CUtensorMap& tmA = out->tmA;
  CUtensorMap& tmB = out->tmB;
  {
    uint32_t box[4] = {static_cast<uint32_t>(bk), BM, 1, 1};
    int rc = encode_map(&tmA, a->a, 4, a->a_dims, a->a_strides, box, tf32);
    if (rc) return rc;
    uint32_t boxb[3] = {static_cast<uint32_t>(bk), static_cast<uint32_t>(use2 ? bn / 2 : bn), 1};
    rc = encode_map(&tmB, a->b, 3, a->b_dims, a->b_strides, boxb, tf32);
    if (rc) return rc;
  }

  GemmKParams& p = out->p;
  p = GemmKParams{};
  p.M = a->M; p.N = a->N; p.K = a->K; p.G1 = a->G1;
  p.dbg_shift = g_dbg_shift;
  if (g_gemm_trace && g_gemm_trace_left > 0) {  // one 256-CTA x 8-slot record per traced call
    p.trace = g_gemm_trace;
    g_gemm_trace += 256 * 8;
    --g_gemm_trace_left;
  }
  p.num_kb = num_kb; p.kpt = kpt; p.dil = a->dil; p.pad = a->pad;
  p.a_c2_g1 = a->a_c2_g1; p.a_c2_g2 = a->a_c2_g2; p.a_c3_g1 = a->a_c3_g1; p.a_c3_g2 = a->a_c3_g2;
  p.b_c2_g1 = a->b_c2_g1; p.b_c2_g2 = a->b_c2_g2; p.b_c2_tile = a->b_c2_tile;
  p.tile_group = a->tile_group; p.m_tiles_used = a->m_tiles_used; p.tile_row0 = a->tile_row0;
  p.mode = a->mode; p.alpha = a->alpha; p.swiglu = a->swiglu;
  p.bias = a->bias; p.bias_g1 = a->bias_g1; p.bias_tile = a->bias_tile;
  p.act = a->act; p.act_slope = a->act_slope;
  p.row_scale = a->row_scale; p.gate = a->gate;
  p.gate_rows = a->gate_rows > 0 ? a->gate_rows : 1; p.gate_ld = a->gate_ld; p.gate_g1 = a->gate_g1;
  p.row_map = a->row_map;
  p.row_mul = a->row_mul; p.row_off = a->row_off; p.row_off_g1 = a->row_off_g1;
  p.rows_out = a->rows_out;
  p.residual = a->residual; p.accumulate = a->accumulate;
  p.res16 = static_cast<const __half*>(a->res16); p.res16_neg = a->res16_neg;
  p.out_f32 = a->out_f32; p.out_ld = a->out_ld; p.out_g1 = a->out_g1; p.out_g2 = a->out_g2;
  p.out_fp16 = static_cast<__half*>(a->out_fp16);
  p.outb_f32 = a->outb_f32;
  p.outb_ld = a->outb_ld; p.outb_g1 = a->outb_g1; p.outb_g2 = a->outb_g2;
  p.act_out = a->act_out; p.act_out_slope = a->act_out_slope; p.out_scale = a->out_scale;
  p.rope_cs = a->rope_cs; p.T = a->T; p.heads = a->heads; p.head_dim = a->head_dim; p.Tp = a->Tp; p.qkv_sec0 = a->qkv_sec0;
  p.q_out = static_cast<__half*>(a->q_out);
  p.k_out = static_cast<__half*>(a->k_out);
  p.vt_out = static_cast<__half*>(a->vt_out);
  p.route_vg = a->route_vg; p.route_vg_g1 = a->route_vg_g1;
  if (a->mode == VB_EPI_GENERIC) {
    if (!a->out_f32 && !a->out_fp16)
      return set_error(VB_ERR_INVALID, "vb_gemm: no output pointer");
    if (a->rows_out <= 0) return set_error(VB_ERR_INVALID, "vb_gemm: rows_out must be set");
    auto al = [](const void* ptr, int bytes) {
      return (reinterpret_cast<uintptr_t>(ptr) % bytes) == 0;
    };
    p.vec_f32 = (a->out_ld % 4 == 0) && (a->out_g1 % 4 == 0) && (a->out_g2 % 4 == 0) &&
                (!a->out_f32 || al(a->out_f32, 16)) && (!a->residual || al(a->residual, 16));
    p.vec_fp16 = (a->outb_ld % (a->outb_f32 ? 4 : 8) == 0) && (a->outb_g1 % 8 == 0) && (a->outb_g2 % 8 == 0) &&
                 (!a->out_fp16 || al(a->out_fp16, 16));
    p.vec_aux = (!a->bias || (al(a->bias, 16) && a->bias_g1 % 4 == 0 && a->bias_tile % 4 == 0)) &&
                (!a->gate || (al(a->gate, 16) && a->gate_ld % 4 == 0 && a->gate_g1 % 4 == 0));
    const long long lim = 1ll << 31;
    p.fast_ok = p.vec_f32 && p.vec_fp16 && p.vec_aux && (long long)a->rows_out * a->out_ld < lim &&
                (long long)a->rows_out * a->outb_ld < lim && a->out_ld >= 0 && a->outb_ld >= 0;
    if (a->accumulate && !a->out_f32)
      return set_error(VB_ERR_INVALID, "vb_gemm: accumulate needs out_f32");
    if (a->residual && a->out_ld <= 0)
      return set_error(VB_ERR_INVALID, "vb_gemm: residual needs out_ld (it shares the fp32 output addressing)");
    if (a->res16 && (a->residual || a->outb_f32 || a->outb_ld <= 0 || (reinterpret_cast<uintptr_t>(a->res16) & 15)))
      return set_error(VB_ERR_INVALID, "vb_gemm: res16 excludes residual, shares the fp16 output addressing (outb_ld) and must be 16-byte aligned");
  }

  // epilogue class: the smallest specialisation that covers the requested fusion
  int ec = EC_FULL;
  {
    const bool lrelu_only = (a->act == VB_ACT_NONE || a->act == VB_ACT_LRELU) &&
                            (a->act_out == VB_ACT_NONE || a->act_out == VB_ACT_LRELU);
    if (a->mode == VB_EPI_ROUTE) ec = EC_ROUTE;
    else if (a->mode == VB_EPI_QKV_ROPE) ec = EC_QKV;
    else if (a->swiglu == 1 && a->alpha == 1.f && !a->out_f32) ec = EC_SWIGLU;
    else if (!a->swiglu && a->alpha == 1.f && lrelu_only) ec = a->res16 ? EC_RES16 : EC_LINEAR;
  }

  p.m_tiles = m_tiles;
  p.n_tiles = (a->N + bn - 1) / bn;
  p.num_tiles = p.m_tiles * p.n_tiles * G;
  out->bn = bn; out->tf32 = tf32; out->use2 = use2; out->ec = ec; out->kpt = kpt; out->num_kb = num_kb; out->G = G;
  return VB_OK;
}
}  // namespace vb

extern "C" int vb_gemm(const vb_gemm_args* a, void* stream_) {
  vb::ProfScope prof_scope(vb::PK_GEMM, stream_);
  using namespace vb;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  GemmPlan pl;
  {
    const int rc = plan_gemm(a, 0, &pl);
    if (rc) return rc;
  }
  const CUtensorMap& tmA = pl.tmA;
  const CUtensorMap& tmB = pl.tmB;
  const GemmKParams& p = pl.p;
  const int bn = pl.bn, ec = pl.ec, kpt = pl.kpt, num_kb = pl.num_kb, G = pl.G, m_tiles = p.m_tiles;
  const bool tf32 = pl.tf32, use2 = pl.use2;
  const int num_sms = device_sm_count();
  // haloed small-C conv path (see vb_conv_halo_kernel)
  if (g_conv_halo > 0 && !tf32 && !use2 && !a->tile_group && a->mode == VB_EPI_GENERIC && kpt == 1 && num_kb >= 2 &&
      num_kb <= 16 && a->N <= 64 && a->a_dims[0] <= 64 && a->dil >= 1 && !a->swiglu && a->G2 == 1 &&
      a->a_c2_g2 == 0 && a->a_c3_g1 == 0 && a->a_c3_g2 == 0 && a->b_c2_g1 == 0 && a->b_c2_g2 == 0 && m_tiles >= 64) {
    HaloParams hp{};
    hp.taps = num_kb;
    hp.row_shift = a->dil;
    hp.rows = BM + (num_kb - 1) * a->dil;
    hp.rows8 = (hp.rows + 7) / 8 * 8;
    hp.n_ksteps = (int)((a->a_dims[0] + 15) / 16);
    hp.base_off = g_conv_halo == 3;  // (experiment: base_offset must stay 0 — the 128B swizzle is a pure function of the smem address)
    const int hbn = a->N <= 32 ? 32 : 64;
    const int w_bytes = ((hp.taps * hbn * 128 + 1023) / 1024) * 1024;
    if (hp.rows <= 256) {
      int stages = (kPipeBytes - w_bytes) / (hp.rows8 * 128);
      if (stages > 6) stages = 6;
      if (stages >= 2) {
        hp.stages = stages;
        CUtensorMap tmA2, tmB2;
        uint32_t boxa[4] = {64, static_cast<uint32_t>(hp.rows), 1, 1};
        int rc = encode_map(&tmA2, a->a, 4, a->a_dims, a->a_strides, boxa, false);
        if (rc) return rc;
        uint32_t boxb2[3] = {64, static_cast<uint32_t>(hbn), 1};
        rc = encode_map(&tmB2, a->b, 3, a->b_dims, a->b_strides, boxb2, false);
        if (rc) return rc;
        GemmKParams ph = p;
        ph.n_tiles = 1;
        ph.num_tiles = m_tiles * G;
        const int smem_bytes = w_bytes + stages * hp.rows8 * 128 + 1024 + 1024 + kEpiStageBytes;
        const int grid_h = ph.num_tiles < num_sms ? ph.num_tiles : num_sms;
        const bool k2 = hp.n_ksteps <= 2;
        if (ec == EC_LINEAR) return launch_halo_class<EC_LINEAR>(hbn, k2, tmA2, tmB2, ph, hp, grid_h, smem_bytes, stream);
        if (ec == EC_RES16) return launch_halo_class<EC_RES16>(hbn, k2, tmA2, tmB2, ph, hp, grid_h, smem_bytes, stream);
        return launch_halo_class<EC_FULL>(hbn, k2, tmA2, tmB2, ph, hp, grid_h, smem_bytes, stream);
      }
    }
  }
  int grid = p.num_tiles < num_sms ? p.num_tiles : num_sms;
  if (use2) {
    const long long pairs = (long long)((m_tiles + 1) / 2) * p.n_tiles * G;
    const int clusters = (int)(pairs < num_sms / 2 ? pairs : num_sms / 2);
    grid = clusters * 2;
  }
  switch (ec) {
    case EC_ROUTE: return launch_gemm_class<EC_ROUTE>(bn, tf32, use2, tmA, tmB, p, grid, stream);
    case EC_QKV: return launch_gemm_class<EC_QKV>(bn, tf32, use2, tmA, tmB, p, grid, stream);
    case EC_SWIGLU: return launch_gemm_class<EC_SWIGLU>(bn, tf32, use2, tmA, tmB, p, grid, stream);
    case EC_LINEAR: return launch_gemm_class<EC_LINEAR>(bn, tf32, use2, tmA, tmB, p, grid, stream);
    case EC_RES16: return launch_gemm_class<EC_RES16>(bn, tf32, use2, tmA, tmB, p, grid, stream);
    default: return launch_gemm_class<EC_FULL>(bn, tf32, use2, tmA, tmB, p, grid, stream);
  }
}

// ------------------------------------------------------------------ vb_gemm_chain
extern "C" int vb_gemm_chain(const vb_gemm_args* ops, const int32_t* dep_mode, int n_ops, int bn, int32_t* counters,
                             int64_t counters_len, void* stream_) {
  vb::ProfScope prof_scope(vb::PK_GEMM, stream_);
  using namespace vb;
  if (!ops || !dep_mode || n_ops < 1 || n_ops > kChainMaxOps || !counters)
    return set_error(VB_ERR_INVALID, "vb_gemm_chain: 1..%d problems, dependency modes and a counter buffer required", kChainMaxOps);
  if (bn != 64 && bn != 128) return set_error(VB_ERR_INVALID, "vb_gemm_chain: bn must be 64 or 128");
  if (dep_mode[0] != 0) return set_error(VB_ERR_INVALID, "vb_gemm_chain: the first problem cannot have a dependency");
  ChainParams cp{};
  cp.n_ops = n_ops;
  int64_t used = 2;  // [0] = error flag, [1] spare
  int tiles = 0;
  bool tf32 = false;
  for (int o = 0; o < n_ops; ++o) {
    GemmPlan pl;
    const int rc = plan_gemm(&ops[o], bn, &pl);
    if (rc) return rc;
    if (pl.ec != EC_LINEAR && pl.ec != EC_SWIGLU)  // (EC_CHAIN = these two)
      return set_error(VB_ERR_INVALID, "vb_gemm_chain: problem %d needs an epilogue outside LINEAR/SWIGLU", o);
    if (o == 0) tf32 = pl.tf32;
    else if (tf32 != pl.tf32) return set_error(VB_ERR_INVALID, "vb_gemm_chain: mixed operand types");
    if (dep_mode[o] < 0 || dep_mode[o] > 2) return set_error(VB_ERR_INVALID, "vb_gemm_chain: bad dependency mode");
    ChainOp& op = cp.op[o];
    op.tmA = pl.tmA; op.tmB = pl.tmB; op.p = pl.p;
    op.tile_start = tiles;
    op.dep_mode = dep_mode[o];
    op.sig_mode = o + 1 < n_ops ? dep_mode[o + 1] : 0;
    if (o > 0 && dep_mode[o] == 1) {
      const GemmKParams& a = cp.op[o - 1].p;
      if (a.m_tiles != pl.p.m_tiles || a.num_tiles / (a.m_tiles * a.n_tiles) != pl.p.num_tiles / (pl.p.m_tiles * pl.p.n_tiles))
        return set_error(VB_ERR_INVALID, "vb_gemm_chain: row dependency needs the same M tiling and group count in problems %d and %d", o - 1, o);
      if ((a.m_tiles_used != nullptr) != (pl.p.m_tiles_used != nullptr))
        return set_error(VB_ERR_INVALID, "vb_gemm_chain: row dependency between a grouped and a dense problem");
    }
    const int groups = pl.p.num_tiles / (pl.p.m_tiles * pl.p.n_tiles);
    op.cnt_rows = counters + used;
    used += (int64_t)groups * pl.p.m_tiles;
    op.cnt_all = counters + used;
    used += 1;
    tiles += pl.p.num_tiles;
  }
  if (used > counters_len)
    return set_error(VB_ERR_INVALID, "vb_gemm_chain: counter buffer too small (%lld ints needed)", (long long)used);
  cp.total_tiles = tiles;
  cp.err_flag = counters;
  cp.dbg = g_chain_dbg;
  if (cp.dbg & 1)
    for (int o = 0; o < n_ops; ++o) cp.op[o].dep_mode = cp.op[o].sig_mode = 0;
  const int num_sms = device_sm_count();
  const int grid = tiles < num_sms ? tiles : num_sms;
  return launch_chain(bn, tf32, cp, grid, static_cast<cudaStream_t>(stream_));
}
