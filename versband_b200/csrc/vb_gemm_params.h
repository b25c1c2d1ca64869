// Kernel-side parameter blocks of vb_gemm and the per-epilogue-class launch entry points. The GEMM
// kernels are instantiated once per EPILOGUE CLASS (vb_gemm_ec*.cu) so that each instance carries only
// the epilogue code it executes: the all-in-one epilogue was ~9000 SASS instructions (144 KB) and ran
// instruction-fetch bound (ncu: stall_no_inst on top, 22 % instruction-cache misses).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include "vb_internal.h"

namespace vb {

constexpr int BM = 128;
constexpr int BK = 64;  // fp16 elements per k-block = one 128-byte swizzle row (32 for fp32/tf32 operands)
constexpr int kEpiWarps = 8;                      // two per TMEM lane quadrant
constexpr int kThreads = 64 + 32 * kEpiWarps;     // + TMA producer warp + MMA warp
constexpr int kEpiColStep = 32 * (kEpiWarps / 4); // columns between two chunks of the same warp
constexpr int kEpiStageBytes = kEpiWarps * 4096;  // per epilogue warp: one 32 x 32 fp32 transpose block
constexpr int kPipeBytes = 192 * 1024;            // operand ring; + barriers + transpose blocks <= 227 KB

// epilogue classes (compile-time specialisations of the fused epilogue)
enum EpiClass {
  EC_FULL = 0,    // everything: alpha, every activation, accumulate, ... (small MLPs, odd layers)
  EC_LINEAR = 1,  // bias, {none, leaky-relu} activations, row_scale/row_map, gate, residual, accumulate, fp32 and/or operand-slot output
  EC_SWIGLU = 2,  // silu(w1 x) * (w3 x) pairs -> operand-slot output (expert FFN first layers)
  EC_QKV = 3,     // bias + rotary + head-major q/k and transposed v stores
  EC_ROUTE = 4,   // per-head softmax over 96 score columns contracted with V.(gate) -> gate-logit partials
  EC_CHAIN = 5,   // LINEAR + SWIGLU selected per problem at run time (vb_gemm_chain)
  EC_RES16 = 6,   // LINEAR + fp16 activated-representation residual (the vocoder's convs): kept out of LINEAR so
                  // that the DiT's issue-bound epilogues do not carry its code (measured +1 us per DiT GEMM)
};

struct GemmKParams {
  int M, N, K;
  int G1;
  int num_kb, kpt, dil, pad;
  int m_tiles, n_tiles, num_tiles;
  int a_c2_g1, a_c2_g2, a_c3_g1, a_c3_g2;
  int b_c2_g1, b_c2_g2, b_c2_tile;
  const int* tile_group;
  const int* m_tiles_used;
  const int* tile_row0;
  // epilogue
  int mode;
  float alpha;
  int swiglu;
  const float* bias;
  long long bias_g1, bias_tile;
  int act;
  float act_slope;
  const float* row_scale;
  const float* gate;
  int gate_rows;
  long long gate_ld, gate_g1;
  const int* row_map;
  int row_mul, row_off, row_off_g1, rows_out;
  const float* residual;
  const __half* res16;  // fp16 residual in the activated representation, addressed like out_fp16
  float res16_neg;      // multiplier that undoes the leaky-relu on negative values (1/slope; 1 = plain fp16)
  int accumulate;
  float* out_f32;
  long long out_ld, out_g1, out_g2;
  __half* out_fp16;
  int outb_f32;  // the operand-output slot stores fp32 (tf32 operand mode) instead of fp16
  long long outb_ld, outb_g1, outb_g2;
  int act_out;
  float act_out_slope, out_scale;
  int dbg_shift;  // experiment: A tile loaded `dbg_shift` rows early, UMMA descriptor shifted back
  int vec_f32, vec_fp16, vec_aux;  // alignment allows 16-byte accesses (outputs / bias+gate)
  int fast_ok;  // all of the above and every row offset fits 32 bits: coalesced epilogue path allowed
  // qkv-rope
  const float* rope_cs;
  int T, heads, head_dim, Tp, qkv_sec0;
  __half* q_out;
  __half* k_out;
  __half* vt_out;
  // route
  const float* route_vg;
  long long route_vg_g1;
  unsigned long long* trace;  // debug (vb_set_gemm_trace): 8 x u64 per CTA, see tools/trace_gemm.py
};


// host-side plan of one GEMM problem (vb_gemm.cu: plan_gemm)
struct GemmPlan {
  CUtensorMap tmA, tmB;
  GemmKParams p;
  int bn, ec, kpt, num_kb, G;
  bool tf32, use2;
};
int plan_gemm(const vb_gemm_args* a, int force_bn, GemmPlan* out);

// ---- vb_gemm_chain: several dependent GEMM problems in ONE persistent launch -------------------------------
// Flat tile list = problem 0's tiles, then problem 1's, ... (n fastest inside a problem, so an M row block
// completes early). A tile of problem o may only start loading A once its producer rows exist:
//   dep_mode 0: no dependency (first problem)
//   dep_mode 1: the same (group, m_tile) of problem o-1 has stored all of its N tiles
//   dep_mode 2: every active tile of problem o-1 has been stored (row-scattering producers)
// Completion is published through global counters (one arrival per tile, after a __threadfence) that the TMA
// producer of a consuming tile polls with ld.acquire; counters are zeroed by the caller before the launch.
constexpr int kChainMaxOps = 4;
struct alignas(64) ChainOp {
  CUtensorMap tmA, tmB;
  GemmKParams p;
  int tile_start;       // first flat tile index of this problem
  int dep_mode;         // how this problem waits on problem o-1
  int sig_mode;         // what this problem publishes: = dep_mode of problem o+1 (0 for the last)
  int* cnt_rows;        // [G * m_tiles] arrivals per (group, m_tile) of THIS problem
  int* cnt_all;         // [1] arrivals over all tiles of THIS problem
};
struct ChainParams {
  ChainOp op[kChainMaxOps];
  int n_ops, total_tiles;
  int* err_flag;        // set to 1 if a dependency wait timed out (never in a correct run)
  int dbg;              // experiments (vb_set_chain_debug): 1 ignore dependencies, 2 no publish fences, 4 m-fastest order
};
int launch_chain(int bn, bool tf32, const ChainParams& cp, int grid, cudaStream_t stream);

struct HaloParams {
  int taps, row_shift, rows, rows8, n_ksteps, stages, base_off;
};

template <int EC>
int launch_gemm_class(int bn, bool tf32, bool use2, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmKParams& p,
                      int grid, cudaStream_t stream);
template <int EC>
int launch_halo_class(int hbn, bool k2, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmKParams& p,
                      const HaloParams& hp, int grid, int smem_bytes, cudaStream_t stream);

}  // namespace vb
