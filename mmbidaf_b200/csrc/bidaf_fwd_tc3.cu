// Third cut of the fused BiDAF forward, tensor-core tier (layers/attention.py:37-75), built on two measurements of round 1
// (profiles/r01_bidaf_tc_ncu.md):
//   * a 128 x N x 16 tcgen05.mma costs 17.4 / 33.4 cycles (N = 32 / 64) with the A operand in TENSOR MEMORY against 41.4 / 49.5
//     with A in shared memory (tools/micro/umma_tmem_a.cu) -- the X tile is constant over the whole tile loop, so it is staged
//     into TMEM once (104 columns) and S = X Y^T runs at its floor: 13 x 33.4 = 434 cycles per 64 columns instead of 2 x 538;
//   * in the two cuts of round 1 the S product, the soft-max and the P V product of a tile are one serial chain per block.
//     Here S is double-buffered in TMEM and P in shared memory: S(t+1) is ISSUED BEFORE P V(t), so the tensor pipe computes it
//     while the threads work on the soft-max of S(t).
// TMEM (one block per SM, 512 columns): X 112 | S0 64 | S1 64 | O 208 = 448.  One accumulator per block, so the c2q pass is
// split into C2QA (a = s1 q) and C2QB (b = s1 T) blocks as in bidaf_fwd_tc2.cu (rebuilding S is cheap now).
// Shared memory: two operand rings -- the S operand of tile t+1 is needed a whole iteration before the value operand of tile t
// and is free again as soon as S(t+1) has run, so the rings turn independently: 4 slots for S operands, 3 for value operands
// (7 x 26 624 B) -- plus P x 2 (32 768 B).  The X tile lands in the last two value slots and is dead once it is in TMEM.
//
// Ordering argument (tensor pipe executes MMAs in issue order; a commit arrives when everything issued before it is done):
//   issue order:  S(0) | S(1) PV(0) | S(2) PV(1) | ...      iteration t issues S(t+1) first and PV(t) last
//   - threads read S(t) after bar_s[t&1]; S(t) was issued after PV(t-2), so P buffer t&1 (read by PV(t-2)) is free to overwrite;
//   - S(t+1) overwrites S buffer (t+1)&1, last read (tcgen05.ld, waited) by every thread in iteration t-1, before its barrier;
//   - the accumulator is only touched by threads (lazy rescale) after waiting for PV(t-1) (v_free of its slot).
#include <stdlib.h>
#include "tc_common.cuh"

namespace mmb {
using namespace tc;
namespace {

constexpr int TX = 128, TY = 64, HALF = TY / 2, NTHREADS = 256;
constexpr int X_BYTES = TX / 8 * GROUP_BYTES;   // 53248
constexpr int Y_BYTES = TY / 8 * GROUP_BYTES;   // 26624
constexpr int P_BYTES = TX * TY * 2;            // 16384: chunk c8 (8 columns) at c8 * 2048 + row * 16
constexpr int S_SLOTS = 4, V_SLOTS = 3;
constexpr int MMA_WARP = 0, TMA_WARP = 1;
constexpr int COL_X = 0, COL_S = 112, COL_O = 240;            // X: 104 columns used of 112; S: 2 x 64; O: 208
constexpr int STG_STRIDE = 204;
constexpr float TAU2 = 11.0f;
constexpr float NEG2 = kNegFill * LOG2E;

enum Kind { Q2C = 0, C2QA = 1, C2QB = 2 };

struct BlockArgs {
  const __nv_bfloat16* x_pack;       // S operand of the X side
  const __nv_bfloat16* s_pack;       // S operand of the Y side
  const __nv_bfloat16* v_pack;       // value operand of the Y side (may equal s_pack)
  const __nv_bfloat16* x_plain;      // C2QA / C2QB: plain text pack for the c*a / c*b products
  const unsigned long long* y_words; // (B, LYP/64, 2)
  const float* bias;
  float* out;                        // Q2C: T fp32 (B, LX, d);  C2QA / C2QB: out (B, LX, 4d)
  __nv_bfloat16* t_pack;             // Q2C: packed T
  float* lse;                        // Q2C: lse_col; C2QA: lse_row; C2QB: null
  float* bm;                         // C2QB: optional (B, LX, d)
  int LX, LXP, LY, LYP, d;
};

struct FusedArgs {
  BlockArgs k[3];
  int* ready;            // (B) zeroed before the launch
  int nq, nc;
  int n_q2c, n_c2q;
  long long* cta_times;  // debugging aid (tools/bidaf_fwd_timeline.py) or null
};

// D[tmem] (+)= A[tmem] * B[smem] (tools/micro/umma_tmem_a.cu: row i of A = lane i, column c = bf16 pair K = 2c, 2c+1)
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                             uint32_t acc, uint32_t leader) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, q;\n\t"
      ".reg .b64 db;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "setp.ne.b32 q, %6, 0;\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t"
      "}" ::"r"(d_tmem), "r"(a_tmem), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc), "r"(leader)
      : "memory");
}

template <int KIND>
__device__ __forceinline__ void block_body(const BlockArgs& a, const int b, const int xblk, int* ready, const int ready_target,
                                           long long* cta_times) {
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* Sr = smem;                                       // S-operand ring: S_SLOTS x Y_BYTES
  unsigned char* Vr = Sr + S_SLOTS * Y_BYTES;                     // value-operand ring: V_SLOTS x Y_BYTES
  unsigned char* Xs = Vr + (V_SLOTS - 2) * Y_BYTES;               // the X tile lands in the last two value slots
  unsigned char* Ps = Vr + V_SLOTS * Y_BYTES;                     // 2 x P_BYTES
  uint64_t* bars = reinterpret_cast<uint64_t*>(Ps + 2 * P_BYTES); // [0] x, [1] final, [2..3] s ready, then full / free of both rings
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4 + 2 * (S_SLOTS + V_SLOTS));
  float* xbuf = reinterpret_cast<float*>(tmem_slot + 4);          // [2][TX] cross-half exchange (max, then sum)
  const bool same_v = a.v_pack == a.s_pack;                       // C2QA without dropout: the S operand is the value operand

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int half = warp >> 2, wq = warp & 3;
  const int warp_u = uniform_warp_idx();
  const uint32_t leader = elect_one();
  const int row = wq * 32 + lane;
  const int x0 = xblk * TX;
  const uint32_t bar_x = smem_u32(bars), bar_final = smem_u32(bars + 1), bar_s0 = smem_u32(bars + 2);
  const uint32_t s_full0 = smem_u32(bars + 4), s_free0 = s_full0 + 8 * S_SLOTS;
  const uint32_t v_full0 = s_free0 + 8 * S_SLOTS, v_free0 = v_full0 + 8 * V_SLOTS;

  if (tid == 0) {
    for (int i = 0; i < 4 + 2 * (S_SLOTS + V_SLOTS); ++i) mbar_init(bar_x + 8 * i, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  if (cta_times && tid == 0) cta_times[4] = globaltimer_ns();

  int nty = (a.LY + TY - 1) / TY;                                  // tiles past the last un-masked Y row add nothing
  {
    int last = 0;
    for (int t = lane; t < nty; t += 32)
      if (a.y_words[((size_t)b * (a.LYP / 64) + t) * 2 + 1] != 0ull) last = t + 1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) last = max(last, __shfl_xor_sync(0xffffffffu, last, o));
    if (last > 0) nty = last;
  }
  const size_t x_off = ((size_t)b * (a.LXP / 8) + x0 / 8) * GROUP_BYTES;
  const size_t y_batch = (size_t)b * (a.LYP / 8) * GROUP_BYTES;
  auto load_s = [&](int t) {                                      // S operand of tile t -> slot t % S_SLOTS
    const uint32_t bar = s_full0 + 8 * (t % S_SLOTS);
    mbar_expect_tx(bar, Y_BYTES, leader);
    tma_bulk_g2s(smem_u32(Sr + (t % S_SLOTS) * Y_BYTES), reinterpret_cast<const char*>(a.s_pack) + y_batch + (size_t)t * Y_BYTES,
                 Y_BYTES, bar, leader);
  };
  auto load_v = [&](int t) {                                      // value operand of tile t -> slot t % V_SLOTS
    const uint32_t bar = v_full0 + 8 * (t % V_SLOTS);
    mbar_expect_tx(bar, Y_BYTES, leader);
    tma_bulk_g2s(smem_u32(Vr + (t % V_SLOTS) * Y_BYTES), reinterpret_cast<const char*>(a.v_pack) + y_batch + (size_t)t * Y_BYTES,
                 Y_BYTES, bar, leader);
  };
  if (warp_u == TMA_WARP) {
    mbar_expect_tx(bar_x, X_BYTES, leader);
    tma_bulk_g2s(smem_u32(Xs), reinterpret_cast<const char*>(a.x_pack) + x_off, X_BYTES, bar_x, leader);
    for (int t = 0; t < S_SLOTS && t < nty; ++t) load_s(t);       // the S operands never depend on the Q2C blocks
    if (KIND == C2QB && ready) {                                  // the value operand is T: wait for this batch row's Q2C blocks
      wait_counter(ready + b, ready_target);
      fence_proxy_async_all();
    }
    if (!same_v && nty > 0) load_v(0);                            // slot 0 is not under the X tile
  }

  // ---- X tile: shared memory -> TMEM, once.  Thread (row, half) copies chunks [16 half, 16 half + 16) of its row:
  //      4 bf16 pairs per 16-byte chunk -> 4 columns; half 1 has 10 real chunks, the rest of its 48 columns is zero padding.
  const uint32_t lane_base = tmem + ((uint32_t)(wq * 32) << 16);
  mbar_wait(bar_x, 0);
  if (cta_times && tid == 0) { cta_times[5] = globaltimer_ns(); cta_times[7] = nty; }
  {
    const unsigned char* xrow = Xs + (row >> 3) * GROUP_BYTES + (row & 7) * 16;
    const int nq16 = half == 0 ? 4 : 3;
#pragma unroll 1
    for (int q = 0; q < nq16; ++q) {
      float v[16];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int ch = half * 16 + q * 4 + c;
        uint4 u = make_uint4(0u, 0u, 0u, 0u);
        if (ch < CHUNKS) u = *reinterpret_cast<const uint4*>(xrow + ch * 128);
        v[c * 4 + 0] = __uint_as_float(u.x); v[c * 4 + 1] = __uint_as_float(u.y);
        v[c * 4 + 2] = __uint_as_float(u.z); v[c * 4 + 3] = __uint_as_float(u.w);
      }
      tmem_st16(lane_base + COL_X + half * 64 + q * 16, v);
    }
    tmem_wait_st();
  }
  fence_proxy_async();                                            // our reads of the X tile precede the TMA writes that reuse it
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp_u == TMA_WARP && !same_v)
    for (int t = 1; t < V_SLOTS && t < nty; ++t) load_v(t);       // the slots the X tile occupied

  const float bias2 = a.bias[0] * LOG2E;
  float m_ref = -INFINITY, l_part = 0.f;
  constexpr uint32_t IDESC_S = idesc_bf16(TY, 0), IDESC_PV = idesc_bf16(DPAD, 1);
  auto issue_s = [&](int t) {                                     // S(t) = X Y_t^T into S buffer t & 1; frees its ring slot
    const int sl = t % S_SLOTS;
    mbar_wait(s_full0 + 8 * sl, (t / S_SLOTS) & 1);
    tc_fence_after();
    const uint32_t y_lo = desc_lo(smem_u32(Sr + sl * Y_BYTES), 128);
#pragma unroll
    for (int k = 0; k < DPAD / 16; ++k)
      umma_bf16_ts(tmem + COL_S + (t & 1) * TY, tmem + COL_X + k * 8, y_lo + k * 16, desc_hi(GROUP_BYTES), IDESC_S, k > 0, leader);
    umma_commit(bar_s0 + 8 * (t & 1), leader);
    if (!same_v) umma_commit(s_free0 + 8 * sl, leader);           // (same_v: the slot is freed by P V(t), which reads it too)
  };
  if (warp_u == MMA_WARP && nty > 0) issue_s(0);

  for (int t = 0; t < nty; ++t) {
    if (warp_u == MMA_WARP && t + 1 < nty) issue_s(t + 1);        // runs on the tensor pipe under this tile's soft-max
    if (warp_u == TMA_WARP && !same_v && t + S_SLOTS < nty) {     // S(t) has run by the time anybody gets past bar_s below:
      mbar_wait(s_free0 + 8 * (t % S_SLOTS), (t / S_SLOTS) & 1);  // its ring slot takes the S operand of tile t + S_SLOTS
      load_s(t + S_SLOTS);
    }
    const ulonglong2 words = *reinterpret_cast<const ulonglong2*>(a.y_words + ((size_t)b * (a.LYP / 64) + t) * 2);
    const uint32_t wvalid = (uint32_t)(words.x >> (HALF * half)), wopen = (uint32_t)(words.y >> (HALF * half));
    const bool all_open = (words.x & words.y) == ~0ull;
    mbar_wait(bar_s0 + 8 * (t & 1), (t >> 1) & 1);
    tc_fence_after();
    if (cta_times && tid == 0 && t == 0) cta_times[6] = globaltimer_ns();

    // ---- this thread's half row of S(t): masked streaming soft-max (base 2) ----------------------------------------
    float sv[HALF];
    tmem_ld16(lane_base + COL_S + (t & 1) * TY + half * HALF, sv);
    tmem_ld16(lane_base + COL_S + (t & 1) * TY + half * HALF + 16, sv + 16);
    float tile_max = -INFINITY;
    if (all_open) {
#pragma unroll
      for (int c = 0; c < HALF; ++c) {
        sv[c] = fmaf(sv[c], LOG2E, bias2);
        tile_max = fmaxf(tile_max, sv[c]);
      }
    } else {
#pragma unroll
      for (int c = 0; c < HALF; ++c) {
        const float v = ((wopen >> c) & 1u) ? fmaf(sv[c], LOG2E, bias2) : NEG2;      // attention.py:94
        sv[c] = v;
        if ((wvalid >> c) & 1u) tile_max = fmaxf(tile_max, v);
      }
    }
    xbuf[half * TX + row] = tile_max;
    tc_fence_before();                                            // our tcgen05.ld of S(t) precede the barrier: its buffer may be
    __syncthreads();                                              // overwritten by S(t+2), issued after the NEXT iteration's barriers
    tile_max = fmaxf(tile_max, xbuf[(half ^ 1) * TX + row]);
    float alpha = 1.f;
    const bool bump = tile_max > m_ref + TAU2;
    if (bump) {
      alpha = fast_exp2(m_ref - tile_max);
      m_ref = tile_max;
    }
    float psum = 0.f;
    uint32_t packed[HALF / 2];
#pragma unroll
    for (int c = 0; c < HALF; c += 2) {
      const float p0 = (all_open || ((wvalid >> c) & 1u)) ? fast_exp2(sv[c] - m_ref) : 0.f;
      const float p1 = (all_open || ((wvalid >> (c + 1)) & 1u)) ? fast_exp2(sv[c + 1] - m_ref) : 0.f;
      psum += p0 + p1;
      const __nv_bfloat162 pk = __floats2bfloat162_rn(p0, p1);
      packed[c / 2] = *reinterpret_cast<const uint32_t*>(&pk);
    }
    l_part = l_part * alpha + psum;
    {
      unsigned char* prow = Ps + (t & 1) * P_BYTES + row * 16 + (half * (HALF / 8)) * 2048;
#pragma unroll
      for (int c8 = 0; c8 < HALF / 8; ++c8)
        *reinterpret_cast<uint4*>(prow + c8 * 2048) = make_uint4(packed[c8 * 4], packed[c8 * 4 + 1], packed[c8 * 4 + 2], packed[c8 * 4 + 3]);
    }
    fence_proxy_async();
    tc_fence_before();
    const int any_bump = __syncthreads_or(bump && t > 0);
    if (any_bump) {                                               // lazy rescale: P V(t-1) must have landed first
      const int pv = t - 1;
      const uint32_t done = same_v ? s_free0 + 8 * (pv % S_SLOTS) : v_free0 + 8 * (pv % V_SLOTS);
      mbar_wait(done, (pv / (same_v ? S_SLOTS : V_SLOTS)) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int q = half; q < DPAD / 16; q += 2) {
        float o[16];
        tmem_ld16(lane_base + COL_O + q * 16, o);
#pragma unroll
        for (int i = 0; i < 16; ++i) o[i] *= alpha;
        tmem_st16(lane_base + COL_O + q * 16, o);
      }
      tmem_wait_st();
      tc_fence_before();
      __syncthreads();
    }
    if (warp_u == TMA_WARP && t >= 1) {                           // P V(t-1) ran under this tile's soft-max: refill what it released
      if (!same_v && t - 1 + V_SLOTS < nty) {                     // (waiting for it at the top of the iteration would stall the soft-max)
        mbar_wait(v_free0 + 8 * ((t - 1) % V_SLOTS), ((t - 1) / V_SLOTS) & 1);
        load_v(t - 1 + V_SLOTS);
      } else if (same_v && t - 1 + S_SLOTS < nty) {               // one ring: the slot of tile t-1 is free after P V(t-1)
        mbar_wait(s_free0 + 8 * ((t - 1) % S_SLOTS), ((t - 1) / S_SLOTS) & 1);
        load_s(t - 1 + S_SLOTS);
      }
    }
    if (warp_u == MMA_WARP) {                                     // O += P(t) V_t
      tc_fence_after();
      uint32_t v_addr;
      if (same_v) {
        v_addr = smem_u32(Sr + (t % S_SLOTS) * Y_BYTES);          // already waited for by issue_s(t)
      } else {
        mbar_wait(v_full0 + 8 * (t % V_SLOTS), (t / V_SLOTS) & 1);
        tc_fence_after();
        v_addr = smem_u32(Vr + (t % V_SLOTS) * Y_BYTES);
      }
      const uint32_t v_lo = desc_lo(v_addr, GROUP_BYTES), ps_lo = desc_lo(smem_u32(Ps + (t & 1) * P_BYTES), 2048);
#pragma unroll
      for (int k = 0; k < TY / 16; ++k)
        umma_bf16_lh(tmem + COL_O, ps_lo + k * 256, desc_hi(128), v_lo + k * 2 * GROUP_BYTES / 16, desc_hi(128), IDESC_PV,
                     (t > 0) || (k > 0), leader);
      if (t == nty - 1) umma_commit(bar_final, leader);
      umma_commit(same_v ? s_free0 + 8 * (t % S_SLOTS) : v_free0 + 8 * (t % V_SLOTS), leader);
    }
  }

  // ---- epilogue: TMEM -> registers -> fp32 staging in shared memory (over the rings) -> coalesced global stores -----------
  xbuf[half * TX + row] = l_part;
  if (nty > 0) mbar_wait(bar_final, 0);
  tc_fence_after();
  __syncthreads();
  if (cta_times && tid == 0) cta_times[1] = globaltimer_ns();
  if (KIND != Q2C && warp_u == TMA_WARP) {                        // the plain text tile for the products: over the value ring's tail
    mbar_expect_tx(bar_x, X_BYTES, leader);
    tma_bulk_g2s(smem_u32(Xs), reinterpret_cast<const char*>(a.x_plain) + x_off, X_BYTES, bar_x, leader);
  }
  const float l_run = xbuf[row] + xbuf[TX + row];
  const int gx = x0 + row;
  const float inv_l = 1.f / l_run;
  if (half == 0 && gx < a.LX && a.lse) a.lse[(size_t)b * a.LX + gx] = (m_ref + log2f(l_run)) * LN2;
  float* stg = reinterpret_cast<float*>(smem);                    // 128 x 204 fp32 = 104448 B = the S ring (106496 B)
  static_assert(TX * STG_STRIDE * 4 <= S_SLOTS * Y_BYTES, "staging fits under the S-operand ring, next to the text tile");
  const int d = a.d, dv4 = d >> 2;
#pragma unroll 1
  for (int q = half; q < DPAD / 16; q += 2) {
    float o[16];
    tmem_ld16(lane_base + COL_O + q * 16, o);
#pragma unroll
    for (int i = 0; i < 16; i += 4)
      if (q * 16 + i < STG_STRIDE)
        *reinterpret_cast<float4*>(stg + row * STG_STRIDE + q * 16 + i) =
            make_float4(o[i] * inv_l, o[i + 1] * inv_l, o[i + 2] * inv_l, o[i + 3] * inv_l);
  }
  __syncthreads();
  if (cta_times && tid == 0) cta_times[8] = globaltimer_ns();
  constexpr int NW = NTHREADS / 32;
  if (KIND == Q2C) {
#pragma unroll 1
    for (int r = warp; r < TX; r += NW) {                         // fp32 T rows: a warp writes one row
      if (x0 + r >= a.LX) break;
      float* trow = a.out + ((size_t)b * a.LX + x0 + r) * d;
      for (int c4 = lane; c4 < dv4; c4 += 32)
        *reinterpret_cast<float4*>(trow + c4 * 4) = *reinterpret_cast<const float4*>(stg + r * STG_STRIDE + c4 * 4);
    }
    char* tp = reinterpret_cast<char*>(a.t_pack) + x_off;         // packed bf16 T, value operand of the C2QB blocks
    for (int i = tid; i < TX * CHUNKS; i += NTHREADS) {
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
  } else {
    mbar_wait(bar_x, 1);                                          // the plain text tile
    if (cta_times && tid == 0) cta_times[9] = globaltimer_ns();
#pragma unroll 2
    for (int r = warp; r < TX; r += NW) {
      if (x0 + r >= a.LX) break;
      float* orow = a.out + ((size_t)b * a.LX + x0 + r) * 4 * d;
      const unsigned char* ctile = Xs + (r >> 3) * GROUP_BYTES + (r & 7) * 16;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c4 = lane + 32 * h;
        if (c4 >= dv4) continue;
        const float4 v = *reinterpret_cast<const float4*>(stg + r * STG_STRIDE + c4 * 4);
        const uint2 cb = *reinterpret_cast<const uint2*>(ctile + (c4 >> 1) * 128 + (c4 & 1) * 8);
        const float2 c01 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&cb.x));
        const float2 c23 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&cb.y));
        const float4 p = make_float4(c01.x * v.x, c01.y * v.y, c23.x * v.z, c23.y * v.w);
        if (KIND == C2QA) {
          *reinterpret_cast<float4*>(orow + d + c4 * 4) = v;
          *reinterpret_cast<float4*>(orow + 2 * d + c4 * 4) = p;
        } else {
          *reinterpret_cast<float4*>(orow + 3 * d + c4 * 4) = p;
          if (a.bm) *reinterpret_cast<float4*>(a.bm + ((size_t)b * a.LX + x0 + r) * d + c4 * 4) = v;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, TMEM_COLS);
  if (KIND == Q2C && ready && tid == 0) signal_counter(ready + b);
}

__global__ void __launch_bounds__(NTHREADS, 1) bidaf_tc3_kernel(const FusedArgs f) {
  const int blk = blockIdx.x;
  long long* times = f.cta_times ? f.cta_times + 12 * (size_t)blk : nullptr;
  if (times && threadIdx.x == 0) {
    times[0] = globaltimer_ns();
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    times[3] = smid;
  }
  if (blk < f.n_q2c) {
    block_body<Q2C>(f.k[Q2C], blk / f.nq, blk % f.nq, f.ready, 0, times);
  } else if (blk < f.n_q2c + f.n_c2q) {
    const int i = blk - f.n_q2c;
    block_body<C2QA>(f.k[C2QA], i / f.nc, i % f.nc, nullptr, 0, times);
  } else {
    const int i = blk - f.n_q2c - f.n_c2q;
    block_body<C2QB>(f.k[C2QB], i / f.nc, i % f.nc, f.ready, f.nq, times);
  }
  if (times && threadIdx.x == 0) times[2] = globaltimer_ns();
}

constexpr size_t SMEM_BYTES = (size_t)(S_SLOTS + V_SLOTS) * Y_BYTES + 2 * P_BYTES + (4 + 2 * (S_SLOTS + V_SLOTS)) * 8 + 16 + 2 * TX * 4;
static_assert(SMEM_BYTES <= 227 * 1024, "one block per SM");
static_assert(X_BYTES <= 2 * Y_BYTES, "the X tile fits in two value slots");

}  // namespace

// Same contract as bidaf_fwd_tc2_launch (bidaf_fwd_tc2.cu): after bidaf_pack_kernel, one launch for Q2C, C2QA and C2QB blocks.
int bidaf_fwd_tc3_launch(const BidafPacks& pk, const float* bias, float* out, float* q2c, float* bm, float* lse_row,
                         float* lse_col, int B, int Lc, int Lq, int d, cudaStream_t stream) {
  static_assert(PACK_ROWS == TX, "pack padding must match the X tile");
  const int LcP = pk.LcP, LqP = pk.LqP;
  FusedArgs f{};
  f.k[Q2C] = BlockArgs{pk.qs, pk.cw, pk.cp, nullptr, pk.c_words, bias, q2c, pk.tp, lse_col, nullptr, Lq, LqP, Lc, LcP, d};
  f.k[C2QA] = BlockArgs{pk.cw, pk.qs, pk.qp, pk.cp, pk.q_words, bias, out, nullptr, lse_row, nullptr, Lc, LcP, Lq, LqP, d};
  f.k[C2QB] = BlockArgs{pk.cw, pk.qs, pk.tp, pk.cp, pk.q_words, bias, out, nullptr, nullptr, bm, Lc, LcP, Lq, LqP, d};
  f.ready = pk.ready;
  f.nq = LqP / TX;
  f.nc = LcP / TX;
  f.n_q2c = B * f.nq;
  f.n_c2q = B * f.nc;
  const char* ct = getenv("MMB_BIDAF_FWD_CTA_TIMES");
  f.cta_times = ct ? reinterpret_cast<long long*>(strtoull(ct, nullptr, 0)) : nullptr;
  MMB_CUDA(cudaMemsetAsync(pk.ready, 0, sizeof(int) * (size_t)B, stream));
  MMB_CUDA(cudaFuncSetAttribute(bidaf_tc3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
  bidaf_tc3_kernel<<<f.n_q2c + 2 * f.n_c2q, NTHREADS, SMEM_BYTES, stream>>>(f);
  return check_launch("bidaf_tc3_kernel");
}

}  // namespace mmb
