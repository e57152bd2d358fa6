// Fused BiDAF attention backward, tensor-core tier (sm_100a: tcgen05.mma + TMEM + TMA bulk copies).
// bf16 operands, fp32 accumulation; rel <= 2e-2 tier of north_star.  The similarity matrix, both soft-maxes and
// dS are recomputed tile by tile on chip from the saved log-sum-exp vectors: nothing of size (Lc x Lq) touches HBM.
//
// Gradient of layers/attention.py:37-75 (P = s1 row soft-max, R = s2 column soft-max, A = P q,
// T = R^T c, Bm = P T, upstream G = [G0 G1 G2 G3]):
//     dA = G1 + c o G2       dBm = c o G3       dc <- G0 + A o G2 + Bm o G3
//     Drow_i = sum_j P_ij dP_ij = dA_i.A_i + dBm_i.Bm_i          (no pass over S needed)
//     dq <- P^T dA           dT = P^T dBm       Dcol_j = sum_i R_ij dR_ij = dT_j.T_j
//     dP = dA q^T + dBm T^T  dR = c dT^T        dc += R dT
//     dS = mq o P o (dP - Drow) + mc o R o (dR - Dcol)
//     dc~ = rowsum(dS) w_c^T + (dS q~) o w_cq   dq~ = colsum(dS) w_q^T + dS^T (c~ o w_cq)
//     dw_c = sum c~^T rowsum(dS)   dw_q = sum q~^T colsum(dS)   dw_cq = sum c~ o (dS q~)   dbias = sum dS
//
// Launches (all operands are bf16 "packs" in UMMA core-matrix order, see tc_common.cuh):
//   1. bidaf_bwd_prep_kernel     per text row: dA, dBm packs, Drow, dc <- G0 + A o G2 + Bm o G3
//   2. bidaf_bwd_tc_kernel<PT>   X = 128 modality rows, streams text tiles: P^T from lse_row; dq <- P^T dA,
//                                dT = P^T dBm (fp32 -> bf16 pack), Dcol
//   3. bidaf_bwd_tc_kernel<DC>   X = 64 text rows, streams modality tiles: S, dP, dR -> dS; acc0 = dS q~, acc1 = R dT,
//                                rowsum(dS) summed in fp32 registers; epilogue: dc, partial dw_c, dw_cq, dbias
//   4. bidaf_bwd_tc_kernel<DQ>   X = 64 modality rows, streams text tiles: S^T, dP^T, dR^T -> dS^T;
//                                acc0 = dS^T (c~ o w_cq), colsum(dS) in registers; epilogue: dq, partial dw_q
//   5. bidaf_bwd_reduce_kernel   deterministic sum of the per-CTA weight-gradient partials
// 2.-4. are ONE launch (bidaf_bwd_fused_kernel): PT blocks first, then the DC and DQ blocks, which wait on a per-batch-row
// counter for the PT blocks they depend on.
//
// A DC / DQ CTA needs four X-side and four Y-side operands (1664 bytes per row and side), so it owns 64 rows and
// issues M = 64 MMAs (25 cycles for 64 x 32 x 16 against 40 for M = 128, tools/micro/umma_rate.cu); their accumulator
// rows sit in lanes 0-15 of each 32-lane TMEM quarter, so all eight warps take part with their lower half-warps.
#include <stdlib.h>
#include "tc_common.cuh"

namespace mmb {
using namespace tc;
namespace {

constexpr int TY = 32;                              // Y rows per streamed tile
constexpr int Y_PART = TY / 8 * GROUP_BYTES;        // 13312 bytes per operand per tile
constexpr int MAX_STAGES = 3;
constexpr int COL_S = 0, COL_GA = 32, COL_GB = 64, COL_ACC0 = 96, COL_ACC1 = 96 + DPAD;   // 96 + 2 * 208 = 512
constexpr int STG_STRIDE = 204;
constexpr int NTHREADS = 256;
constexpr int PART_STRIDE = 3 * DPAD;               // per-CTA partials: [sum x~ rs | sum x~ acc | sum rs]
constexpr float NEG2 = kNegFill * LOG2E;

enum Mode { PT = 0, DC = 1, DQ = 2 };

// Ask L2 for [p, p + bytes) (16-byte granules inside the range): the epilogue's fp32 rows are requested when the block starts, so
// that a whole tile loop later they are L2 hits instead of a burst of DRAM reads issued by every SM at the same moment.
__device__ __forceinline__ void l2_prefetch(const void* p, size_t bytes) {
  const uintptr_t lo = (reinterpret_cast<uintptr_t>(p) + 15) & ~uintptr_t(15), hi = (reinterpret_cast<uintptr_t>(p) + bytes) & ~uintptr_t(15);
  if (hi > lo) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(lo), "r"((uint32_t)(hi - lo)) : "memory");
}

// ---------------------------------------------------------------------------------------------------------
// 1. prep
// ---------------------------------------------------------------------------------------------------------
struct PrepArgs {
  const float* grad;         // (B, L, 4d)
  const float* text;         // (B, L, d)
  const float* out;          // (B, L, 4d) forward output: block 1 (a) is read
  const float* bm;           // (B, L, d)  b = s1 T
  __nv_bfloat16* da_pack;
  __nv_bfloat16* dbm_pack;
  float* d_text;             // (B, L, d)
  float* d_row;              // (B, L)
  int L, LP, d;
};

constexpr int PACK_CHUNK_STRIDE = 144;

__device__ __forceinline__ void ld8(const float* p, float* v) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p + 4));
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

__global__ void __launch_bounds__(256) bidaf_bwd_prep_kernel(const PrepArgs a) {
  __shared__ __align__(16) unsigned char stage[8][CHUNKS * PACK_CHUNK_STRIDE];
  const int b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = blockIdx.x * 8 + warp;
  const int d = a.d, nchunk = d >> 3;
  uint4 da_out[8], dbm_out[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int row = g * 8 + r;
    da_out[r] = dbm_out[r] = make_uint4(0u, 0u, 0u, 0u);
    float dot = 0.f;
    const bool on = row < a.L && lane < nchunk;
    if (on) {
      const size_t r4 = ((size_t)b * a.L + row) * 4 * d + lane * 8, r1 = ((size_t)b * a.L + row) * d + lane * 8;
      float g0[8], g1[8], g2[8], g3[8], c[8], av[8], bv[8];
      ld8(a.grad + r4, g0); ld8(a.grad + r4 + d, g1); ld8(a.grad + r4 + 2 * d, g2); ld8(a.grad + r4 + 3 * d, g3);
      ld8(a.text + r1, c); ld8(a.out + r4 + d, av); ld8(a.bm + r1, bv);
      float da[8], db[8], dc[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        da[e] = fmaf(c[e], g2[e], g1[e]);
        db[e] = c[e] * g3[e];
        dc[e] = fmaf(bv[e], g3[e], fmaf(av[e], g2[e], g0[e]));
        dot = fmaf(da[e], av[e], fmaf(db[e], bv[e], dot));           // Drow = dA.A + dBm.Bm
      }
      *reinterpret_cast<float4*>(a.d_text + r1) = make_float4(dc[0], dc[1], dc[2], dc[3]);
      *reinterpret_cast<float4*>(a.d_text + r1 + 4) = make_float4(dc[4], dc[5], dc[6], dc[7]);
      __nv_bfloat162 pa[4], pb[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        pa[e] = __floats2bfloat162_rn(da[2 * e], da[2 * e + 1]);
        pb[e] = __floats2bfloat162_rn(db[2 * e], db[2 * e + 1]);
      }
      da_out[r] = *reinterpret_cast<uint4*>(pa);
      dbm_out[r] = *reinterpret_cast<uint4*>(pb);
    }
    dot = warp_sum(dot);
    if (lane == 0 && row < a.L) a.d_row[(size_t)b * a.L + row] = dot;
  }
  unsigned char* buf = stage[warp];
  const size_t base = ((size_t)b * (a.LP / 8) + g) * GROUP_BYTES;
#pragma unroll 1
  for (int which = 0; which < 2; ++which) {
    if (lane < CHUNKS) {
#pragma unroll
      for (int r = 0; r < 8; ++r)
        *reinterpret_cast<uint4*>(buf + lane * PACK_CHUNK_STRIDE + r * 16) = which == 0 ? da_out[r] : dbm_out[r];
    }
    __syncwarp();
    uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<char*>(which == 0 ? a.da_pack : a.dbm_pack) + base);
    for (int i = lane; i < CHUNKS * 8; i += 32)
      dst[i] = *reinterpret_cast<const uint4*>(buf + (i >> 3) * PACK_CHUNK_STRIDE + (i & 7) * 16);
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------------------
// 2.-4. tensor-core passes
// ---------------------------------------------------------------------------------------------------------
struct BwdArgs {
  const __nv_bfloat16* x_ops[4];       // X-side packs; [0] is the S operand
  const __nv_bfloat16* y_ops[4];       // Y-side packs; [0] is the S operand and the value operand of acc0
  const unsigned long long* x_words;   // (B, LXP/64, 2)
  const unsigned long long* y_words;   // (B, LYP/64, 2)
  const float* bias;
  const float* norm_x;                 // DC/DQ: lse of the soft-max that runs along Y, per X row (B, LX)
  const float* norm_y;                 // lse of the soft-max that runs along X, per Y row (B, LY)
  const float* dlt_x;                  // DC/DQ: sum_y W1 G1 per X row (Drow for DC, Dcol for DQ)
  const float* dlt_y;                  // DC/DQ: per Y row
  // epilogue
  const float* x_feat;                 // DC/DQ: fp32 X rows (B, LX, d) (un-dropped)
  const uint8_t* x_keep;               // DC/DQ: nullable keep mask (B, LX, d)
  const float* w_term;                 // DC/DQ: (d) weight of the additive term of the X side
  const float* w_fold;                 // DC: (d) text_modality_weight; DQ: null (already folded into the value operand)
  const float* t_feat;                 // PT: fp32 T (B, LX, d)
  float* dx;                           // PT: d_modality (B, LX, d) written; DC/DQ: d_text / d_modality accumulated
  __nv_bfloat16* dt_pack;              // PT: packed dT
  float* d_col;                        // PT: Dcol (B, LX)
  float* part;                         // DC/DQ: (B, nxb, PART_STRIDE) weight-gradient partials
  long long* trace;                    // debugging aid: clock64() stamps of CTA (0,0), or null
  float keep_scale;
  int LX, LXP, LY, LYP, d;
  int prefetch;                        // request the epilogue's rows from L2 at block start
};

// One X block of one pass.  `ready` (per batch row) orders the passes inside ONE launch: a PT block bumps ready[b]
// once its dq rows, dT pack and Dcol are in memory; the DC and DQ blocks of that batch row wait for all of them.
template <int MODE>
__device__ __forceinline__ void bwd_block(const BwdArgs& a, const int b, const int xblk, const int nxb, int* ready,
                                          const int ready_target) {
  constexpr bool IS_PT = MODE == PT;
  constexpr int ROWS = IS_PT ? 128 : 64;                 // real X rows per CTA
  constexpr int NX = IS_PT ? 1 : 4, NY = IS_PT ? 3 : 4;
  constexpr int X_BYTES = ROWS / 8 * GROUP_BYTES;
  constexpr int TILE_BYTES = ROWS * TY * 2;
  constexpr int TILE_LBO = ROWS * 16;
  constexpr int STAGE_BYTES = NY * Y_PART;
  constexpr int HALF = TY / 2;
  constexpr int STAGES = IS_PT ? 3 : 2;                  // what fits next to the resident X operands
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* Xs = smem;
  unsigned char* Ts = Xs + NX * X_BYTES;                 // bf16 tiles fed to the second MMAs: [0] P^T / dS, [1] R (DC)
  unsigned char* St = Ts + 2 * TILE_BYTES;
  float* ycol = reinterpret_cast<float*>(St + STAGES * STAGE_BYTES);     // [2][3][TY]: n2, d2, fill
  uint64_t* bars = reinterpret_cast<uint64_t*>(ycol + 2 * 3 * TY);       // [0] x, [1] mma, [2..4] full, [5..7] free
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 + 2 * MAX_STAGES);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int half = warp >> 2, wq = warp & 3;
  // PT: M = 128 MMAs, accumulator row = TMEM lane (32 rows per warp quarter).  DC / DQ: M = 64 MMAs, whose sixteen rows per
  // quarter sit in lanes 0-15: every warp works, its upper half-warp carries garbage and never stores.
  constexpr int MMA_M = IS_PT ? 128 : 64;
  const int row = IS_PT ? wq * 32 + lane : wq * 16 + (lane & 15);
  const bool owner = IS_PT || lane < 16;
  // issuing warps (one elected lane each, see tc_common.cuh): DC/DQ use two warps that own no rows
  constexpr int MMA_WARP = IS_PT ? 0 : 2, TMA_WARP = IS_PT ? 1 : 3;
  const int warp_u = uniform_warp_idx();
  const uint32_t leader = elect_one();                   // one lane per warp: the issuer in the MMA / TMA warps
  const int x0 = xblk * ROWS;
  if (x0 >= a.LX) return;
  const uint32_t bar_x = smem_u32(bars), bar_mma = smem_u32(bars + 1), bar_full0 = smem_u32(bars + 2);
  const uint32_t bar_free0 = smem_u32(bars + 2 + MAX_STAGES);

  if (tid == 0) {
    mbar_init(bar_x, 1);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full0 + 8 * s, 1);
      mbar_init(bar_free0 + 8 * s, 1);
    }
    mbar_init(bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);

  int nty = (a.LY + TY - 1) / TY;
  if (MODE == DC) {
    // Y rows (modality positions) past the last un-masked one: P = 0 there, hence dT = Dcol = 0 and every term of dS and
    // of R dT vanishes -- stop at the last tile that has an un-masked row (all tiles if nothing is un-masked: then the
    // row soft-max is uniform and dT is not zero).
    int last = 0;
    for (int w = lane; w < (a.LY + 63) / 64; w += 32) {
      const unsigned long long open = a.y_words[((size_t)b * (a.LYP / 64) + w) * 2 + 1];
      if (open != 0ull) last = 2 * w + ((open >> 32) != 0ull ? 2 : 1);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) last = max(last, __shfl_xor_sync(0xffffffffu, last, o));
    if (last > 0) nty = min(nty, last);
  }
  const size_t x_off = ((size_t)b * (a.LXP / 8) + x0 / 8) * GROUP_BYTES;
  const size_t y_batch = (size_t)b * (a.LYP / 8) * GROUP_BYTES;
  if (a.prefetch && tid == 32 * 5) {                     // (a warp that issues neither MMAs nor TMA loads)
    const size_t e0 = ((size_t)b * a.LX + x0) * a.d, ne = (size_t)min(ROWS, a.LX - x0) * a.d;
    if (IS_PT) {
      l2_prefetch(a.t_feat + e0, ne * 4);
    } else {
      l2_prefetch(a.x_feat + e0, ne * 4);
      l2_prefetch(a.dx + e0, ne * 4);
      if (a.x_keep) l2_prefetch(a.x_keep + e0, ne);
    }
  }
  const float log2_lx = log2f((float)a.LX);
  // per-column scalars of tile t -> ycol[t & 1], by one warp (an idle one in DC/DQ), in two steps so that the global
  // loads are issued a whole tile before their values are stored (visible after the next __syncthreads)
  const int yl = tid - (IS_PT ? 0 : 6 * 32);
  float yl_lse = 0.f, yl_d2 = 0.f;
  auto fetch_ycol = [&](int t) {
    const int y = t * TY + yl;
    if ((unsigned)yl < (unsigned)TY && y < a.LY) {
      yl_lse = a.norm_y[(size_t)b * a.LY + y];
      if (!IS_PT) yl_d2 = a.dlt_y[(size_t)b * a.LY + y];
    }
  };
  auto store_ycol = [&](int t) {
    if ((unsigned)yl < (unsigned)TY) {
      float n2 = 0.f, d2 = 0.f, fill = NEG2;
      if (t * TY + yl < a.LY) {
        // a fully masked soft-max is uniform (attention.py:94 uses -1e30, not -inf): every logit of the column
        // is the fill value, so logit - lse = -log(LX) -- which fp32 cannot hold next to 1e30; re-base both to 0.
        const bool degenerate = yl_lse < -5e29f;
        n2 = degenerate ? log2_lx : yl_lse * LOG2E;
        fill = degenerate ? 0.f : NEG2;
        d2 = yl_d2;
      }
      float* yc = ycol + (t & 1) * 3 * TY;
      yc[yl] = n2;
      yc[TY + yl] = d2;
      yc[2 * TY + yl] = fill;
    }
  };
  if (!IS_PT && ready) {                                 // dT pack, Dcol and dq come from this batch row's PT blocks
    if (tid == 0) wait_counter(ready + b, ready_target);
    __syncthreads();
    fence_proxy_async_all();                             // their generic-proxy stores -> our async-proxy (TMA) loads
  }
  fetch_ycol(0);
  store_ycol(0);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  auto issue_stage = [&](int t) {
    const int s = t % STAGES;
    const uint32_t bar = bar_full0 + 8 * s;
    const uint32_t dst = smem_u32(St + s * STAGE_BYTES);
    const size_t off = y_batch + (size_t)t * Y_PART;
    mbar_expect_tx(bar, STAGE_BYTES, leader);
#pragma unroll
    for (int p = 0; p < NY; ++p)
      tma_bulk_g2s(dst + p * Y_PART, reinterpret_cast<const char*>(a.y_ops[p]) + off, Y_PART, bar, leader);
  };
  if (warp_u == TMA_WARP) {
    mbar_expect_tx(bar_x, NX * X_BYTES, leader);
#pragma unroll
    for (int p = 0; p < NX; ++p)
      tma_bulk_g2s(smem_u32(Xs + p * X_BYTES), reinterpret_cast<const char*>(a.x_ops[p]) + x_off, X_BYTES, bar_x, leader);
    for (int t = 0; t < STAGES && t < nty; ++t) issue_stage(t);
  }

  // ---- per-thread constants of this X row ------------------------------------------------------------------------
  const float bias2 = a.bias[0] * LOG2E;
  const int gx = x0 + row;
  bool valid_x = false, open_x = false;
  float n1 = 0.f, d1 = 0.f;
  if (owner) {
    const ulonglong2 xw = *reinterpret_cast<const ulonglong2*>(a.x_words + ((size_t)b * (a.LXP / 64) + gx / 64) * 2);
    valid_x = (xw.x >> (gx & 63)) & 1ull;
    open_x = (xw.y >> (gx & 63)) & 1ull;
    if (!IS_PT && valid_x) {
      n1 = a.norm_x[(size_t)b * a.LX + gx] * LOG2E;
      d1 = a.dlt_x[(size_t)b * a.LX + gx];
    }
  }
  const uint32_t lane_base = tmem + ((uint32_t)(wq * 32) << 16);
  uint32_t mma_phase = 0;
  float rsum = 0.f;                                      // DC/DQ: fp32 sum over y of this thread's dS columns
  constexpr uint32_t IDESC_S = idesc_bf16(TY, 0, MMA_M), IDESC_PV = idesc_bf16(DPAD, 1, MMA_M);
  const uint32_t xs_lo = desc_lo(smem_u32(Xs), 128), ts_lo = desc_lo(smem_u32(Ts), TILE_LBO);

  const bool tracing = a.trace != nullptr && b == 0 && xblk == 0 && tid == MMA_WARP * 32;
  int ntrace = 0;
  auto stamp = [&]() {
    if (tracing && ntrace < 250) a.trace[ntrace++] = clock64();
  };
  stamp();
  if (warp_u == MMA_WARP) mbar_wait(bar_x, 0);
  stamp();
  for (int t = 0; t < nty; ++t) {
    const int s = t % STAGES;
    const uint32_t st_addr = smem_u32(St + s * STAGE_BYTES);
    if (warp_u == MMA_WARP) {
      mbar_wait(bar_full0 + 8 * s, (t / STAGES) & 1);
      stamp();
      tc_fence_after();
      const uint32_t st_lo = desc_lo(st_addr, 128);
      // product p: X operand p times Y operand p, K-major both, into S / GA / GB
      constexpr int NPROD = IS_PT ? 1 : 4;
#pragma unroll
      for (int p = 0; p < NPROD; ++p) {
        // DC: GA = dA q^T + dBm T^T (p = 1, 2), GB = c dT^T (p = 3);  DQ: GA = dT c^T (p = 1), GB = q dA^T + T dBm^T (p = 2, 3)
        const int col = p == 0 ? COL_S : p == 1 ? COL_GA : p == 3 ? COL_GB : (MODE == DC ? COL_GA : COL_GB);
        const bool fresh = p == 0 || p == 1 || (p == 2 && MODE == DQ) || (p == 3 && MODE == DC);
#pragma unroll
        for (int k = 0; k < DPAD / 16; ++k)
          umma_bf16_lh(tmem + col, xs_lo + (p * X_BYTES + k * 256) / 16, desc_hi(GROUP_BYTES),
                       st_lo + (p * Y_PART + k * 256) / 16, desc_hi(GROUP_BYTES), IDESC_S, !(fresh && k == 0), leader);
      }
      umma_commit(bar_mma, leader);
      stamp();
    }
    const ulonglong2 words = *reinterpret_cast<const ulonglong2*>(a.y_words + ((size_t)b * (a.LYP / 64) + (t >> 1)) * 2);
    const int sh = (t & 1) * 32 + half * HALF;
    const uint32_t wvalid = (uint32_t)(words.x >> sh) & 0xffffu, wopen = (uint32_t)(words.y >> sh) & 0xffffu;
    if (t + 1 < nty) fetch_ycol(t + 1);
    // the previous tile's second MMAs commit to the "free" barrier of their stage: refill it while this tile's
    // first MMAs run, so the load has a whole tile of tensor-core time to land
    if (warp_u == TMA_WARP && t >= 1 && t - 1 + STAGES < nty) {
      mbar_wait(bar_free0 + 8 * ((t - 1) % STAGES), ((t - 1) / STAGES) & 1);
      issue_stage(t - 1 + STAGES);
    }
    mbar_wait(bar_mma, mma_phase);
    mma_phase ^= 1;
    tc_fence_after();
    stamp();

    {
      const float* yc = ycol + (t & 1) * 3 * TY + half * HALF;
      float sv[HALF];
      tmem_ld16(lane_base + COL_S + half * HALF, sv);
      uint32_t pk0[HALF / 2], pk1[HALF / 2];
      if (IS_PT) {
        float n2[HALF];                                       // broadcast LDS.128 (2.3 cycles per warp each), not 16 scalar reads
#pragma unroll
        for (int c = 0; c < HALF; c += 4) *reinterpret_cast<float4*>(n2 + c) = *reinterpret_cast<const float4*>(yc + c);
        const bool masked_row = __any_sync(0xffffffffu, valid_x && !open_x);   // rare: only then are the fill values needed
#pragma unroll
        for (int c = 0; c < HALF; c += 2) {
          float w[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            float logit = fmaf(sv[c + e], LOG2E, bias2);
            if (masked_row && !open_x) logit = yc[2 * TY + c + e];
            w[e] = (valid_x && ((wvalid >> (c + e)) & 1u)) ? fast_exp2(logit - n2[c + e]) : 0.f;
          }
          const __nv_bfloat162 v = __floats2bfloat162_rn(w[0], w[1]);
          pk0[c / 2] = *reinterpret_cast<const uint32_t*>(&v);
        }
      } else {
        float ga[HALF], gb[HALF], n2[HALF], d2[HALF];
        tmem_ld16(lane_base + COL_GA + half * HALF, ga);
        tmem_ld16(lane_base + COL_GB + half * HALF, gb);
#pragma unroll
        for (int c = 0; c < HALF; c += 4) {
          *reinterpret_cast<float4*>(n2 + c) = *reinterpret_cast<const float4*>(yc + c);
          *reinterpret_cast<float4*>(d2 + c) = *reinterpret_cast<const float4*>(yc + TY + c);
        }
        // interior tile (every column in range and un-masked: uniform over the CTA) and no masked row in this warp:
        // no bit tests, no re-based logits
        const bool interior = (wvalid & wopen) == 0xffffu && __all_sync(0xffffffffu, open_x || !valid_x);
        if (interior) {
#pragma unroll
          for (int c = 0; c < HALF; c += 2) {
            float ds[2], rr[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const float s2 = fmaf(sv[c + e], LOG2E, bias2);
              const float w1 = valid_x ? fast_exp2(s2 - n1) : 0.f, w2 = valid_x ? fast_exp2(s2 - n2[c + e]) : 0.f;   // rows past LX: zero
              const float t2 = w2 * (gb[c + e] - d2[c + e]);
              ds[e] = fmaf(w1, ga[c + e] - d1, t2);
              rsum += t2;
              rr[e] = w2;
            }
            const __nv_bfloat162 v = __floats2bfloat162_rn(ds[0], ds[1]);
            pk0[c / 2] = *reinterpret_cast<const uint32_t*>(&v);
            if (MODE == DC) {
              const __nv_bfloat162 r2 = __floats2bfloat162_rn(rr[0], rr[1]);
              pk1[c / 2] = *reinterpret_cast<const uint32_t*>(&r2);
            }
          }
        } else {
#pragma unroll
          for (int c = 0; c < HALF; c += 2) {
            float ds[2], rr[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const float s2 = fmaf(sv[c + e], LOG2E, bias2);
              const bool vy = (wvalid >> (c + e)) & 1u, oy = (wopen >> (c + e)) & 1u;
              const float w1 = fast_exp2(s2 - n1), w2 = fast_exp2(s2 - n2[c + e]);
              const float t1 = (valid_x && oy) ? w1 * (ga[c + e] - d1) : 0.f;
              const float t2 = (open_x && vy) ? w2 * (gb[c + e] - d2[c + e]) : 0.f;
              ds[e] = t1 + t2;
              rsum += t2;            // sum_y of the W1 part is zero analytically (soft-max along y): leave its noise out
              if (MODE == DC) rr[e] = (valid_x && vy) ? (open_x ? w2 : fast_exp2(yc[2 * TY + c + e] - n2[c + e])) : 0.f;
            }
            const __nv_bfloat162 v = __floats2bfloat162_rn(ds[0], ds[1]);
            pk0[c / 2] = *reinterpret_cast<const uint32_t*>(&v);
            if (MODE == DC) {
              const __nv_bfloat162 r2 = __floats2bfloat162_rn(rr[0], rr[1]);
              pk1[c / 2] = *reinterpret_cast<const uint32_t*>(&r2);
            }
          }
        }
      }
      // tiles in core-matrix order: chunk c8 (8 columns) at c8 * TILE_LBO + row * 16
      unsigned char* trow = Ts + row * 16 + (half * (HALF / 8)) * TILE_LBO;
      if (owner) {
#pragma unroll
        for (int c8 = 0; c8 < HALF / 8; ++c8) {
          *reinterpret_cast<uint4*>(trow + c8 * TILE_LBO) = make_uint4(pk0[c8 * 4], pk0[c8 * 4 + 1], pk0[c8 * 4 + 2], pk0[c8 * 4 + 3]);
          if (MODE == DC)
            *reinterpret_cast<uint4*>(trow + TILE_BYTES + c8 * TILE_LBO) =
                make_uint4(pk1[c8 * 4], pk1[c8 * 4 + 1], pk1[c8 * 4 + 2], pk1[c8 * 4 + 3]);
        }
      }
    }
    if (t + 1 < nty) store_ycol(t + 1);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    stamp();
    if (warp_u == MMA_WARP) {
      tc_fence_after();
      // acc0 += tile0 * value0;  PT: value0 = dA (part 1), acc1 += tile0 * dBm (part 2);  DC: acc1 += R * dT (part 3)
      const uint32_t v_lo = desc_lo(st_addr, GROUP_BYTES);
#pragma unroll
      for (int k = 0; k < TY / 16; ++k)
        umma_bf16_lh(tmem + COL_ACC0, ts_lo + k * 2 * TILE_LBO / 16, desc_hi(128),
                     v_lo + ((IS_PT ? 1 : 0) * Y_PART + k * 2 * GROUP_BYTES) / 16, desc_hi(128), IDESC_PV, (t > 0) || (k > 0),
                     leader);
      if (MODE != DQ) {
#pragma unroll
        for (int k = 0; k < TY / 16; ++k)
          umma_bf16_lh(tmem + COL_ACC1, ts_lo + ((IS_PT ? 0 : TILE_BYTES) + k * 2 * TILE_LBO) / 16, desc_hi(128),
                       v_lo + ((IS_PT ? 2 : 3) * Y_PART + k * 2 * GROUP_BYTES) / 16, desc_hi(128), IDESC_PV,
                       (t > 0) || (k > 0), leader);
      }
      umma_commit(t == nty - 1 ? bar_mma : bar_free0 + 8 * s, leader);
    }
  }
  mbar_wait(bar_mma, mma_phase);
  tc_fence_after();
  __syncthreads();
  stamp();

  // ---- epilogue: TMEM -> fp32 staging in shared memory (over the operands) -> coalesced global traffic ---------------
  float* stg = reinterpret_cast<float*>(smem);
  const int d = a.d, dv4 = d >> 2;
  auto drain = [&](int col0, float* dst) {
#pragma unroll 1
    for (int q = half; q < DPAD / 16; q += 2) {
      float o[16];
      tmem_ld16(lane_base + col0 + q * 16, o);
      if (owner) {
#pragma unroll
        for (int i = 0; i < 16; i += 4)
          if (q * 16 + i < STG_STRIDE)
            *reinterpret_cast<float4*>(dst + row * STG_STRIDE + q * 16 + i) = make_float4(o[i], o[i + 1], o[i + 2], o[i + 3]);
      }
    }
  };
  if (IS_PT) {
    drain(COL_ACC0, stg);                                           // dq <- P^T dA
    __syncthreads();
    for (int i = tid; i < ROWS * dv4; i += NTHREADS) {
      const int r = i / dv4, c4 = i - r * dv4;
      if (x0 + r < a.LX)
        *reinterpret_cast<float4*>(a.dx + ((size_t)b * a.LX + x0 + r) * d + c4 * 4) =
            *reinterpret_cast<const float4*>(stg + r * STG_STRIDE + c4 * 4);
    }
    __syncthreads();
    drain(COL_ACC1, stg);                                           // dT
    __syncthreads();
    char* tp = reinterpret_cast<char*>(a.dt_pack) + x_off;
    for (int i = tid; i < ROWS * CHUNKS; i += NTHREADS) {
      const int g8 = i / (CHUNKS * 8), rem = i - g8 * CHUNKS * 8, ch = rem >> 3, r8 = rem & 7;
      const int r = g8 * 8 + r8;
      __nv_bfloat162 v[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int col = ch * 8 + 2 * e;
        const bool ok = (x0 + r < a.LX) && col < d;
        v[e] = __floats2bfloat162_rn(ok ? stg[r * STG_STRIDE + col] : 0.f, ok ? stg[r * STG_STRIDE + col + 1] : 0.f);
      }
      *reinterpret_cast<uint4*>(tp + (size_t)i * 16) = *reinterpret_cast<uint4*>(v);
    }
    // Dcol_j = dT_j . T_j: a warp per row, the T loads of eight rows in flight at a time
    constexpr int NW = NTHREADS / 32, RB = 8;
#pragma unroll 1
    for (int r0 = warp; r0 < ROWS; r0 += NW * RB) {
      float4 tv[RB][2];
#pragma unroll
      for (int u = 0; u < RB; ++u) {
        const int r = r0 + u * NW;
        const float* trow = a.t_feat + ((size_t)b * a.LX + min(x0 + r, a.LX - 1)) * d;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int c4 = lane + 32 * h;
          tv[u][h] = c4 < dv4 ? __ldg(reinterpret_cast<const float4*>(trow + c4 * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
#pragma unroll
      for (int u = 0; u < RB; ++u) {
        const int r = r0 + u * NW;
        float dot = 0.f;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int c4 = lane + 32 * h;
          if (c4 < dv4) {
            const float4 sv4 = *reinterpret_cast<const float4*>(stg + r * STG_STRIDE + c4 * 4);
            dot += tv[u][h].x * sv4.x + tv[u][h].y * sv4.y + tv[u][h].z * sv4.z + tv[u][h].w * sv4.w;
          }
        }
        dot = warp_sum(dot);
        if (lane == 0 && x0 + r < a.LX) a.d_col[(size_t)b * a.LX + x0 + r] = dot;
      }
    }
  } else {
    float* stg1 = stg + ROWS * STG_STRIDE;
    float* xsum = ycol;                                              // [2][ROWS] halves of sum_y dS (the tile scalars are dead)
    static_assert(IS_PT || 2 * ROWS <= 2 * 3 * TY, "xsum fits in ycol");
    drain(COL_ACC0, stg);
    if (MODE == DC) drain(COL_ACC1, stg1);
    if (owner) xsum[half * ROWS + row] = rsum;
    __syncthreads();
    // Thread = one float4 column x a strided set of rows, loads batched ahead of the stores (the loop would
    // otherwise serialise on DRAM latency: the dx store may alias the next row's loads).
    const int ngrp = NTHREADS / dv4, rg = tid / dv4, c4 = tid - rg * dv4;
    const int nrow = min(ROWS, a.LX - x0);
    float* red = reinterpret_cast<float*>(St);                       // [ngrp][2][d] + [ngrp] partial sums
    float pt[4] = {0.f, 0.f, 0.f, 0.f}, pf[4] = {0.f, 0.f, 0.f, 0.f}, p_sum = 0.f;
    if (rg < ngrp) {
      float wtv[4], wfv[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        wtv[e] = a.w_term[c4 * 4 + e];
        wfv[e] = a.w_fold ? a.w_fold[c4 * 4 + e] : 1.f;
      }
      constexpr int BATCH = 7;                                         // 13 rows per thread at d = 200: two rounds of loads
#pragma unroll 1
      for (int r0 = rg; r0 < nrow; r0 += BATCH * ngrp) {
        float4 xv[BATCH], dv[BATCH];
        uchar4 kv[BATCH];
#pragma unroll
        for (int u = 0; u < BATCH; ++u) {
          const int r = r0 + u * ngrp;
          if (r < nrow) {
            const size_t gi = ((size_t)b * a.LX + x0 + r) * d + c4 * 4;
            xv[u] = __ldg(reinterpret_cast<const float4*>(a.x_feat + gi));
            dv[u] = *reinterpret_cast<const float4*>(a.dx + gi);
            kv[u] = a.x_keep ? __ldg(reinterpret_cast<const uchar4*>(a.x_keep + gi)) : make_uchar4(1, 1, 1, 1);
          }
        }
#pragma unroll
        for (int u = 0; u < BATCH; ++u) {
          const int r = r0 + u * ngrp;
          if (r < nrow) {
            const float rs = xsum[r] + xsum[ROWS + r];
            const float4 v04 = *reinterpret_cast<const float4*>(stg + r * STG_STRIDE + c4 * 4);
            float4 v14 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (MODE == DC) v14 = *reinterpret_cast<const float4*>(stg1 + r * STG_STRIDE + c4 * 4);
            const float x[4] = {xv[u].x, xv[u].y, xv[u].z, xv[u].w}, v0[4] = {v04.x, v04.y, v04.z, v04.w};
            const float v1[4] = {v14.x, v14.y, v14.z, v14.w};
            const unsigned char kk[4] = {kv[u].x, kv[u].y, kv[u].z, kv[u].w};
            float o[4] = {dv[u].x, dv[u].y, dv[u].z, dv[u].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float ks = a.x_keep ? (kk[e] ? a.keep_scale : 0.f) : 1.f;
              const float xd = x[e] * ks;
              o[e] += ks * fmaf(rs, wtv[e], v0[e] * wfv[e]) + v1[e];
              pt[e] = fmaf(xd, rs, pt[e]);
              pf[e] = fmaf(xd, v0[e], pf[e]);
            }
            p_sum += rs;
            const size_t gi = ((size_t)b * a.LX + x0 + r) * d + c4 * 4;
            *reinterpret_cast<float4*>(a.dx + gi) = make_float4(o[0], o[1], o[2], o[3]);
          }
        }
      }
      *reinterpret_cast<float4*>(red + (rg * 2 + 0) * d + c4 * 4) = make_float4(pt[0], pt[1], pt[2], pt[3]);
      *reinterpret_cast<float4*>(red + (rg * 2 + 1) * d + c4 * 4) = make_float4(pf[0], pf[1], pf[2], pf[3]);
      if (c4 == 0) red[ngrp * 2 * d + rg] = p_sum;
    }
    __syncthreads();
    if (tid < d) {                                                   // fixed-order sum over the row groups
      float s_term = 0.f, s_fold = 0.f, s_sum = 0.f;
      for (int g = 0; g < ngrp; ++g) {
        s_term += red[(g * 2 + 0) * d + tid];
        s_fold += red[(g * 2 + 1) * d + tid];
        if (tid == 0) s_sum += red[ngrp * 2 * d + g];
      }
      float* part = a.part + ((size_t)b * nxb + xblk) * PART_STRIDE;
      part[tid] = s_term;
      part[DPAD + tid] = s_fold;
      if (tid == 0) part[2 * DPAD] = s_sum;
    }
  }
  tc_fence_before();
  __syncthreads();
  stamp();
  if (tracing) a.trace[255] = ntrace;
  if (warp == 0) tmem_dealloc(tmem, TMEM_COLS);
  if (IS_PT && ready && tid == 0) signal_counter(ready + b);     // after the barrier: every thread's stores are ordered before it
}

// (separate launches: debugging aid, MMB_BIDAF_BWD_STAGES)
template <int MODE>
__global__ void __launch_bounds__(NTHREADS, 1) bidaf_bwd_tc_kernel(const BwdArgs a) {
  bwd_block<MODE>(a, blockIdx.y, blockIdx.x, gridDim.x, nullptr, 0);
}

// All three passes in one launch: PT blocks first in block order, then the DC and DQ blocks, each of which starts as
// soon as the PT blocks of ITS batch row are done -- no launch boundaries, no partially filled waves in between.
// Block order: PT, then the LONGER of the two dependent kinds first (a DQ block streams Lc / 32 tiles, a DC block Lq / 32): with the
// long blocks at the end of the grid the launch ended on a few SMs running one of them alone (config 2: 6 waves of 148, DQ blocks of
// 35 us last).
struct BwdFusedArgs {
  BwdArgs k[3];
  int* ready;                  // (B) zeroed before the launch
  int n_pt, n_dc, n_dq;        // blocks of PT, DC and DQ (B * blocks per batch row)
  int nb[3];                   // X blocks per batch row of PT, DC, DQ
  int dq_first;
};

__global__ void __launch_bounds__(NTHREADS, 1) bidaf_bwd_fused_kernel(const BwdFusedArgs f) {
  const int blk = blockIdx.x;
  if (blk < f.n_pt) {
    bwd_block<PT>(f.k[PT], blk / f.nb[PT], blk % f.nb[PT], f.nb[PT], f.ready, 0);
    return;
  }
  int i = blk - f.n_pt;
  bool is_dq;
  if (f.dq_first) {
    is_dq = i < f.n_dq;
    if (!is_dq) i -= f.n_dq;
  } else {
    is_dq = i >= f.n_dc;
    if (is_dq) i -= f.n_dc;
  }
  if (is_dq) bwd_block<DQ>(f.k[DQ], i / f.nb[DQ], i % f.nb[DQ], f.nb[DQ], f.ready, f.nb[PT]);
  else bwd_block<DC>(f.k[DC], i / f.nb[DC], i % f.nb[DC], f.nb[DC], f.ready, f.nb[PT]);
}

template <int MODE>
constexpr size_t bwd_smem_bytes() {
  constexpr int ROWS = MODE == PT ? 128 : 64;
  constexpr int NX = MODE == PT ? 1 : 4, NY = MODE == PT ? 3 : 4;
  constexpr int STAGES = MODE == PT ? 3 : 2;
  return (size_t)NX * (ROWS / 8 * GROUP_BYTES) + 2 * (ROWS * TY * 2) + (size_t)STAGES * NY * Y_PART + 2 * 3 * TY * 4 +
         (2 + 2 * MAX_STAGES) * 8 + 16;
}

// ---------------------------------------------------------------------------------------------------------
// 5. weight-gradient reduction (fixed order: deterministic)
// ---------------------------------------------------------------------------------------------------------
struct ReduceArgs {
  const float* part_c;   // (B, gx_c, PART_STRIDE), the first nb_c blocks of every batch row are live
  const float* part_q;
  float *d_w_text, *d_w_cross, *d_w_modality, *d_bias;
  int B, gx_c, nb_c, gx_q, nb_q, d;
};

constexpr int RED_COLS = 16;       // columns per block; 256 threads = 16 columns x 16 partial-row groups

__global__ void __launch_bounds__(256) bidaf_bwd_reduce_kernel(const ReduceArgs a) {
  __shared__ float red[256];
  const int nblk = (a.d + RED_COLS - 1) / RED_COLS;
  const int which = blockIdx.x / nblk;                 // 0: dw_text, 1: dw_cross, 2: dw_modality, 3: dbias
  const int tid = threadIdx.x;
  float acc = 0.f;
  if (which < 3) {
    const int k = (blockIdx.x - which * nblk) * RED_COLS + (tid & (RED_COLS - 1)), grp = tid / RED_COLS;
    const float* part = which == 2 ? a.part_q : a.part_c;
    const int gx = which == 2 ? a.gx_q : a.gx_c, nb = which == 2 ? a.nb_q : a.nb_c;
    const int off = which == 1 ? DPAD : 0;
    if (k < a.d)
      for (int i = grp; i < a.B * nb; i += 256 / RED_COLS) {
        const int b = i / nb, xb = i - b * nb;
        acc += part[((size_t)b * gx + xb) * PART_STRIDE + off + k];
      }
    red[tid] = acc;
    __syncthreads();
    if (grp == 0 && k < a.d) {
      float s = 0.f;
      for (int g = 0; g < 256 / RED_COLS; ++g) s += red[g * RED_COLS + tid];
      (which == 0 ? a.d_w_text : which == 1 ? a.d_w_cross : a.d_w_modality)[k] = s;
    }
  } else {
    for (int i = tid; i < a.B * a.nb_c; i += 256) {
      const int b = i / a.nb_c, xb = i - b * a.nb_c;
      acc += a.part_c[((size_t)b * a.gx_c + xb) * PART_STRIDE + 2 * DPAD];
    }
    red[tid] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
      if (tid < s) red[tid] += red[tid + s];
      __syncthreads();
    }
    if (tid == 0) a.d_bias[0] = red[0];
  }
}

struct BwdWorkspace {
  __nv_bfloat16 *da_pack, *dbm_pack, *dt_pack;
  float *d_row, *d_col, *part_c, *part_q;
  int* ready;                // (B) PT -> DC / DQ dependency counters of the fused launch
  long long* trace;          // 3 x 256 clock stamps (MMB_BIDAF_BWD_TRACE=1)
  size_t bytes;
};
BwdWorkspace bwd_workspace(void* workspace, int B, int Lc, int Lq) {
  const BidafPacks pk = bidaf_packs(nullptr, B, Lc, Lq, false);
  char* ws = static_cast<char*>(workspace);
  BwdWorkspace w;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    char* p = ws + off;
    off += (bytes + 255) / 256 * 256;
    return p;
  };
  w.da_pack = reinterpret_cast<__nv_bfloat16*>(take(pk.c_pack));
  w.dbm_pack = reinterpret_cast<__nv_bfloat16*>(take(pk.c_pack));
  w.dt_pack = reinterpret_cast<__nv_bfloat16*>(take(pk.q_pack));
  w.d_row = reinterpret_cast<float*>(take((size_t)B * Lc * 4));
  w.d_col = reinterpret_cast<float*>(take((size_t)B * Lq * 4));
  w.part_c = reinterpret_cast<float*>(take((size_t)B * (pk.LcP / 64) * PART_STRIDE * 4));
  w.part_q = reinterpret_cast<float*>(take((size_t)B * (pk.LqP / 64) * PART_STRIDE * 4));
  w.ready = reinterpret_cast<int*>(take((size_t)B * 4));
  w.trace = reinterpret_cast<long long*>(take(3 * 256 * 8));
  w.bytes = off;
  return w;
}

}  // namespace

size_t bidaf_bwd_tc_workspace_bytes(int B, int Lc, int Lq) { return bwd_workspace(nullptr, B, Lc, Lq).bytes; }

int bidaf_bwd_tc(const float* grad_out, const float* text, const float* modality, const float* w_text,
                 const float* w_modality, const float* w_cross, const float* bias, const uint8_t* keep_text,
                 const uint8_t* keep_modality, float keep_scale, const float* out, const float* bm, const float* q2c,
                 const float* lse_row, const float* lse_col, const void* fwd_workspace, void* workspace, float* d_text,
                 float* d_modality, float* d_w_text, float* d_w_modality, float* d_w_cross, float* d_bias, int B, int Lc,
                 int Lq, int d, cudaStream_t stream) {
  MMB_REQUIRE(d % 8 == 0 && d <= 200, MMB_ERR_UNSUPPORTED, "mmb_bidaf_bwd (bf16 tier): d=%d (need d %% 8 == 0, d <= 200)", d);
  MMB_REQUIRE(fwd_workspace && workspace, MMB_ERR_INVALID, "mmb_bidaf_bwd (bf16 tier): workspace is null");
  const BidafPacks pk = bidaf_packs(const_cast<void*>(fwd_workspace), B, Lc, Lq, keep_modality != nullptr);
  const BwdWorkspace w = bwd_workspace(workspace, B, Lc, Lq);
  const int LcP = pk.LcP, LqP = pk.LqP;
  // debugging aid: MMB_BIDAF_BWD_STAGES = bit mask of the launches to run (1 prep, 2 PT, 4 DC, 8 DQ, 16 reduce)
  const char* env = getenv("MMB_BIDAF_BWD_STAGES");
  const int stages = env ? atoi(env) : 31;
  const bool tracing = getenv("MMB_BIDAF_BWD_TRACE") != nullptr;   // stamps land at the end of `workspace`

  if (stages & 1) {
  PrepArgs pa{grad_out, text, out, bm, w.da_pack, w.dbm_pack, d_text, w.d_row, Lc, LcP, d};
  bidaf_bwd_prep_kernel<<<dim3(LcP / 64, B), 256, 0, stream>>>(pa);
  if (int rc = check_launch("bidaf_bwd_prep_kernel")) return rc;
  }
  BwdArgs pt{}, dc{}, dq{};
  {   // PT: X = modality rows, Y = text rows
    BwdArgs& a = pt;
    a.x_ops[0] = pk.qs;
    a.y_ops[0] = pk.cw; a.y_ops[1] = w.da_pack; a.y_ops[2] = w.dbm_pack;
    a.x_words = pk.q_words; a.y_words = pk.c_words; a.bias = bias;
    a.norm_y = lse_row;
    a.t_feat = q2c; a.dx = d_modality; a.dt_pack = w.dt_pack; a.d_col = w.d_col;
    a.LX = Lq; a.LXP = LqP; a.LY = Lc; a.LYP = LcP; a.d = d;
    a.trace = tracing ? w.trace : nullptr;
  }
  {   // DC: X = text rows, Y = modality rows
    BwdArgs& a = dc;
    a.x_ops[0] = pk.cw; a.x_ops[1] = w.da_pack; a.x_ops[2] = w.dbm_pack; a.x_ops[3] = pk.cp;
    a.y_ops[0] = pk.qs; a.y_ops[1] = pk.qp; a.y_ops[2] = pk.tp; a.y_ops[3] = w.dt_pack;
    a.x_words = pk.c_words; a.y_words = pk.q_words; a.bias = bias;
    a.norm_x = lse_row; a.norm_y = lse_col; a.dlt_x = w.d_row; a.dlt_y = w.d_col;
    a.x_feat = text; a.x_keep = keep_text; a.w_term = w_text; a.w_fold = w_cross;
    a.dx = d_text; a.part = w.part_c; a.keep_scale = keep_scale;
    a.LX = Lc; a.LXP = LcP; a.LY = Lq; a.LYP = LqP; a.d = d;
    a.trace = tracing ? w.trace + 256 : nullptr;
  }
  {   // DQ: X = modality rows, Y = text rows
    BwdArgs& a = dq;
    a.x_ops[0] = pk.qs; a.x_ops[1] = w.dt_pack; a.x_ops[2] = pk.qp; a.x_ops[3] = pk.tp;
    a.y_ops[0] = pk.cw; a.y_ops[1] = pk.cp; a.y_ops[2] = w.da_pack; a.y_ops[3] = w.dbm_pack;
    a.x_words = pk.q_words; a.y_words = pk.c_words; a.bias = bias;
    a.norm_x = lse_col; a.norm_y = lse_row; a.dlt_x = w.d_col; a.dlt_y = w.d_row;
    a.x_feat = modality; a.x_keep = keep_modality; a.w_term = w_modality; a.w_fold = nullptr;
    a.dx = d_modality; a.part = w.part_q; a.keep_scale = keep_scale;
    a.LX = Lq; a.LXP = LqP; a.LY = Lc; a.LYP = LcP; a.d = d;
    a.trace = tracing ? w.trace + 512 : nullptr;
  }
  static const char* pf_env = getenv("MMB_BIDAF_BWD_PREFETCH");
  pt.prefetch = dc.prefetch = dq.prefetch = pf_env ? atoi(pf_env) : 1;
  constexpr size_t smem_pt = bwd_smem_bytes<PT>(), smem_dc = bwd_smem_bytes<DC>(), smem_dq = bwd_smem_bytes<DQ>();
  static_assert(smem_dc <= 227 * 1024 && smem_dq <= 227 * 1024 && smem_pt <= 227 * 1024, "shared memory");
  if (!env) {   // the normal path: one launch for the three tensor-core passes
    constexpr size_t smem = smem_dc > smem_pt ? (smem_dc > smem_dq ? smem_dc : smem_dq) : (smem_pt > smem_dq ? smem_pt : smem_dq);
    static const char* ord_env = getenv("MMB_BIDAF_BWD_ORDER");         // 0: PT, DC, DQ (round 1); default: longest kind first
    const int dq_first = ord_env ? atoi(ord_env) : (Lc >= Lq ? 1 : 0);
    BwdFusedArgs f{{pt, dc, dq}, w.ready, B * (LqP / 128), B * (LcP / 64), B * (LqP / 64), {LqP / 128, LcP / 64, LqP / 64}, dq_first};
    MMB_CUDA(cudaMemsetAsync(w.ready, 0, sizeof(int) * (size_t)B, stream));
    MMB_CUDA(cudaFuncSetAttribute(bidaf_bwd_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    bidaf_bwd_fused_kernel<<<f.n_pt + f.n_dc + f.n_dq, NTHREADS, smem, stream>>>(f);
    if (int rc = check_launch("bidaf_bwd_fused_kernel")) return rc;
  } else {
    if (stages & 2) {
      MMB_CUDA(cudaFuncSetAttribute(bidaf_bwd_tc_kernel<PT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_pt));
      bidaf_bwd_tc_kernel<PT><<<dim3(LqP / 128, B), NTHREADS, smem_pt, stream>>>(pt);
      if (int rc = check_launch("bidaf_bwd_tc_kernel<PT>")) return rc;
    }
    if (stages & 4) {
      MMB_CUDA(cudaFuncSetAttribute(bidaf_bwd_tc_kernel<DC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_dc));
      bidaf_bwd_tc_kernel<DC><<<dim3(LcP / 64, B), NTHREADS, smem_dc, stream>>>(dc);
      if (int rc = check_launch("bidaf_bwd_tc_kernel<DC>")) return rc;
    }
    if (stages & 8) {
      MMB_CUDA(cudaFuncSetAttribute(bidaf_bwd_tc_kernel<DQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_dq));
      bidaf_bwd_tc_kernel<DQ><<<dim3(LqP / 64, B), NTHREADS, smem_dq, stream>>>(dq);
      if (int rc = check_launch("bidaf_bwd_tc_kernel<DQ>")) return rc;
    }
  }
  if (!(stages & 16)) return MMB_OK;
  ReduceArgs ra{w.part_c, w.part_q, d_w_text, d_w_cross, d_w_modality, d_bias, B, LcP / 64, (Lc + 63) / 64, LqP / 64,
                (Lq + 63) / 64, d};
  bidaf_bwd_reduce_kernel<<<3 * ((d + RED_COLS - 1) / RED_COLS) + 1, 256, 0, stream>>>(ra);
  return check_launch("bidaf_bwd_reduce_kernel");
}

}  // namespace mmb
