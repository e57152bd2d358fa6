// Fused BiDAF attention forward, tensor-core tier, main launch -- two CTAs per SM.
//
// Same operands ("packs", tc_common.cuh), same streaming soft-max and the same tcgen05 / TMEM / TMA machinery as
// the one-block-per-SM cut in bidaf_fwd_tc.cu (which long sequences still use), re-cut so that TWO thread blocks share an
// SM: measurements ("What the measurements say") showed that with one 512-column CTA per SM the load,
// MMA / soft-max and store phases of a block serialise, and the store phase alone (~15 us per C2Q block, bound by
// the ~30 B/clk an SM can write) was half of its life.  Here every block owns ONE 208-column accumulator and
// 32-column S tiles (240 TMEM columns, 256 allocated), 128 threads (one per X row: no cross-thread exchange in the
// soft-max) and <= 113 KB of shared memory, so a second block's MMAs and soft-max run under the first one's stores.
//
//   Q2C  block: X = 128 modality rows, streams text tiles      T = softmax_i(S)^T c           (+ packed bf16 T)
//   C2QA block: X = 128 text rows, streams modality tiles      a = softmax_j(S) q  -> out blocks 1, 2; lse_row
//   C2QB block: X = 128 text rows, streams modality tiles      b = softmax_j(S) T  -> out block 3 (and bm)
// The c2q pass is split in two blocks that each rebuild S (cheap: 13 MMAs of 128x32x16 per tile) because two
// accumulators do not fit in half of TMEM.  All blocks live in ONE launch, ordered Q2C, C2QA, C2QB: a C2QB block
// waits for the Q2C blocks of its batch row (ready[b] counter); C2QA blocks need nothing and fill the SMs meanwhile.
#include <stdlib.h>
#include "tc_common.cuh"

namespace mmb {
using namespace tc;
namespace {

constexpr int TX = 128, TY = 32, NTHREADS = 128;
constexpr int X_BYTES = TX / 8 * GROUP_BYTES;   // 53248
constexpr int Y_BYTES = TY / 8 * GROUP_BYTES;   // 13312
constexpr int P_BYTES = TX * TY * 2;            // 8192: chunk c8 (8 columns) at c8 * 2048 + row * 16
constexpr int MAX_STAGES = 4;
constexpr int MMA_WARP = 0, TMA_WARP = 1;
constexpr int TMEM_ALLOC = 256, COL_S = 0, COL_O = 32;
constexpr int STG_STRIDE = 204;
constexpr float TAU2 = 11.0f;                   // lazy-rescale threshold in log2 units
constexpr float NEG2 = kNegFill * LOG2E;

enum Kind { Q2C = 0, C2QA = 1, C2QB = 2 };

struct BlockArgs {
  const __nv_bfloat16* x_pack;       // S operand of the X side
  const __nv_bfloat16* parts[2];     // per-stage Y operands: [0] S operand, [1] value operand (may be null: values = parts[0])
  const __nv_bfloat16* x_plain;      // C2QA / C2QB: plain (un-dropped, un-folded) text pack for the c*a / c*b products
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
  int nq, nc;            // X blocks per batch row: modality side, text side
  int n_q2c, n_c2q;      // B * nq, B * nc
  long long* cta_times;  // debugging aid (tools/bidaf_fwd_timeline.py) or null
};

template <int KIND>
__device__ __forceinline__ void block_body(const BlockArgs& a, const int b, const int xblk, int* ready, const int ready_target,
                                           long long* cta_times) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int nparts = a.parts[1] ? 2 : 1;
  const int stage_bytes = nparts * Y_BYTES;
  const int STAGES = nparts == 1 ? 4 : 2;                        // what fits in ~113 KB next to X and P
  unsigned char* Xs = smem;
  unsigned char* Ps = Xs + X_BYTES;
  unsigned char* St = Ps + P_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(St + STAGES * stage_bytes);   // [0] x, [1] mma, [2..5] full, [6..9] free
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 + 2 * MAX_STAGES);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row = tid;                                            // one thread per X row = TMEM lane
  const int warp_u = uniform_warp_idx();
  const uint32_t leader = elect_one();
  const int x0 = xblk * TX;
  const uint32_t bar_x = smem_u32(bars), bar_mma = smem_u32(bars + 1), bar_full0 = smem_u32(bars + 2);
  const uint32_t bar_free0 = smem_u32(bars + 2 + MAX_STAGES);

  if (tid == 0) {
    mbar_init(bar_x, 1);
    mbar_init(bar_mma, 1);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full0 + 8 * s, 1);
      mbar_init(bar_free0 + 8 * s, 1);
    }
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), TMEM_ALLOC);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  if (cta_times && tid == 0) cta_times[4] = globaltimer_ns();

  // Tiles past the last un-masked Y row contribute exp(-1e30 - m) = 0 to every soft-max: stop there.  (If nothing at
  // all is un-masked the soft-max is uniform over the whole range, attention.py:94, and every tile is needed.)
  int nty = (a.LY + TY - 1) / TY;
  {
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
  auto issue_stage = [&](int t) {
    const int s = t % STAGES;
    const uint32_t bar = bar_full0 + 8 * s;
    const uint32_t dst = smem_u32(St + s * stage_bytes);
    const size_t off = y_batch + (size_t)t * Y_BYTES;
    mbar_expect_tx(bar, stage_bytes, leader);
    tma_bulk_g2s(dst, reinterpret_cast<const char*>(a.parts[0]) + off, Y_BYTES, bar, leader);
    if (nparts == 2) tma_bulk_g2s(dst + Y_BYTES, reinterpret_cast<const char*>(a.parts[1]) + off, Y_BYTES, bar, leader);
  };
  if (warp_u == TMA_WARP) {
    mbar_expect_tx(bar_x, X_BYTES, leader);
    tma_bulk_g2s(smem_u32(Xs), reinterpret_cast<const char*>(a.x_pack) + x_off, X_BYTES, bar_x, leader);
    if (KIND == C2QB && ready) {                                  // the stages carry T: wait for this batch row's Q2C blocks
      wait_counter(ready + b, ready_target);
      fence_proxy_async_all();                                    // their generic-proxy stores -> our async-proxy (TMA) loads
    }
    for (int t = 0; t < STAGES && t < nty; ++t) issue_stage(t);
  }

  const float bias2 = a.bias[0] * LOG2E;
  const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
  float m_ref = -INFINITY, l_run = 0.f;                              // log2 domain
  uint32_t mma_phase = 0;
  constexpr uint32_t IDESC_S = idesc_bf16(TY, 0), IDESC_PV = idesc_bf16(DPAD, 1);
  const uint32_t xs_lo = desc_lo(smem_u32(Xs), 128), ps_lo = desc_lo(smem_u32(Ps), 2048);

  if (warp_u == MMA_WARP) mbar_wait(bar_x, 0);
  if (cta_times && tid == 0) { cta_times[5] = globaltimer_ns(); cta_times[7] = nty; }
  for (int t = 0; t < nty; ++t) {
    const int s = t % STAGES;
    const uint32_t st_addr = smem_u32(St + s * stage_bytes);
    if (warp_u == MMA_WARP) {
      mbar_wait(bar_full0 + 8 * s, (t / STAGES) & 1);
      tc_fence_after();
      const uint32_t st_lo = desc_lo(st_addr, 128);
#pragma unroll
      for (int k = 0; k < DPAD / 16; ++k)                       // S = X Y^T, both K-major
        umma_bf16_lh(tmem + COL_S, xs_lo + k * 16, desc_hi(GROUP_BYTES), st_lo + k * 16, desc_hi(GROUP_BYTES), IDESC_S, k > 0,
                     leader);
      umma_commit(bar_mma, leader);
    }
    const ulonglong2 words = *reinterpret_cast<const ulonglong2*>(a.y_words + ((size_t)b * (a.LYP / 64) + (t >> 1)) * 2);
    const uint32_t wvalid = (uint32_t)(words.x >> ((t & 1) * 32)), wopen = (uint32_t)(words.y >> ((t & 1) * 32));
    const bool all_open = (wvalid & wopen) == ~0u;              // CTA-uniform: interior tile, nothing masked
    // the previous tile's P V MMAs commit to the "free" barrier of their stage: refill it while this tile's S MMAs run
    if (warp_u == TMA_WARP && t >= 1 && t - 1 + STAGES < nty) {
      mbar_wait(bar_free0 + 8 * ((t - 1) % STAGES), ((t - 1) / STAGES) & 1);
      issue_stage(t - 1 + STAGES);
    }
    mbar_wait(bar_mma, mma_phase);
    mma_phase ^= 1;
    tc_fence_after();
    if (cta_times && tid == 0 && t == 0) cta_times[6] = globaltimer_ns();
    if (KIND != Q2C && warp_u == TMA_WARP && t == nty - 1) {    // X operand no longer needed: fetch the plain text tile
      mbar_expect_tx(bar_x, X_BYTES, leader);
      tma_bulk_g2s(smem_u32(Xs), reinterpret_cast<const char*>(a.x_plain) + x_off, X_BYTES, bar_x, leader);
    }

    // ---- this thread's row of S: masked streaming soft-max (base 2) ------------------------------------------------
    float sv[TY];
    tmem_ld16(lane_base + COL_S, sv);
    tmem_ld16(lane_base + COL_S + 16, sv + 16);
    float tile_max = -INFINITY;
    if (all_open) {
      float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};   // four independent chains
#pragma unroll
      for (int c = 0; c < TY; ++c) {
        sv[c] = fmaf(sv[c], LOG2E, bias2);
        mx[c & 3] = fmaxf(mx[c & 3], sv[c]);
      }
      tile_max = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
    } else {
#pragma unroll
      for (int c = 0; c < TY; ++c) {
        const float v = ((wopen >> c) & 1u) ? fmaf(sv[c], LOG2E, bias2) : NEG2;   // attention.py:94
        sv[c] = v;
        if ((wvalid >> c) & 1u) tile_max = fmaxf(tile_max, v);
      }
    }
    float alpha = 1.f;
    const bool bump = tile_max > m_ref + TAU2;                  // first tile: m_ref = -inf -> always
    if (bump) {
      alpha = fast_exp2(m_ref - tile_max);                          // 0 on the first tile
      m_ref = tile_max;
    }
    float psum = 0.f;
    uint32_t packed[TY / 2];
    if (all_open) {
      float ps[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int c = 0; c < TY; c += 2) {
        const float p0 = fast_exp2(sv[c] - m_ref), p1 = fast_exp2(sv[c + 1] - m_ref);
        ps[(c >> 1) & 3] += p0 + p1;
        const __nv_bfloat162 pk = __floats2bfloat162_rn(p0, p1);
        packed[c / 2] = *reinterpret_cast<const uint32_t*>(&pk);
      }
      psum = (ps[0] + ps[1]) + (ps[2] + ps[3]);
    } else {
#pragma unroll
      for (int c = 0; c < TY; c += 2) {
        const float p0 = ((wvalid >> c) & 1u) ? fast_exp2(sv[c] - m_ref) : 0.f;
        const float p1 = ((wvalid >> (c + 1)) & 1u) ? fast_exp2(sv[c + 1] - m_ref) : 0.f;
        psum += p0 + p1;
        const __nv_bfloat162 pk = __floats2bfloat162_rn(p0, p1);
        packed[c / 2] = *reinterpret_cast<const uint32_t*>(&pk);
      }
    }
    l_run = l_run * alpha + psum;
    {
      unsigned char* prow = Ps + row * 16;
#pragma unroll
      for (int c8 = 0; c8 < TY / 8; ++c8)
        *reinterpret_cast<uint4*>(prow + c8 * 2048) =
            make_uint4(packed[c8 * 4], packed[c8 * 4 + 1], packed[c8 * 4 + 2], packed[c8 * 4 + 3]);
    }
    if (__any_sync(0xffffffffu, bump && t > 0)) {               // lazy rescale of this warp's accumulator rows (alpha = 1 where no bump)
#pragma unroll 1
      for (int q = 0; q < DPAD / 16; ++q) {
        float o[16];
        tmem_ld16(lane_base + COL_O + q * 16, o);
#pragma unroll
        for (int i = 0; i < 16; ++i) o[i] *= alpha;
        tmem_st16(lane_base + COL_O + q * 16, o);
      }
      tmem_wait_st();
    }
    fence_proxy_async();                                        // st.shared P -> visible to the tensor core
    tc_fence_before();
    __syncthreads();
    if (warp_u == MMA_WARP) {
      tc_fence_after();
      const uint32_t v_lo = desc_lo(st_addr + (nparts - 1) * Y_BYTES, GROUP_BYTES);
#pragma unroll
      for (int k = 0; k < TY / 16; ++k)                         // O += P V (V MN-major: LBO = group stride)
        umma_bf16_lh(tmem + COL_O, ps_lo + k * 256, desc_hi(128), v_lo + k * 2 * GROUP_BYTES / 16, desc_hi(128), IDESC_PV,
                     (t > 0) || (k > 0), leader);
      umma_commit(t == nty - 1 ? bar_mma : bar_free0 + 8 * s, leader);
    }
  }
  // ---- epilogue: TMEM -> registers -> fp32 staging in smem (over X, P and the stages) -> coalesced global stores ----
  mbar_wait(bar_mma, mma_phase);
  tc_fence_after();
  if (cta_times && tid == 0) cta_times[1] = globaltimer_ns();
  const int gx = x0 + row;
  const float inv_l = 1.f / l_run;
  if (a.lse && gx < a.LX) a.lse[(size_t)b * a.LX + gx] = (m_ref + log2f(l_run)) * LN2;
  const int d = a.d, dv4 = d >> 2;
  constexpr int NW = NTHREADS / 32;
  auto drain = [&](float* dst_row) {                            // this thread's accumulator row, normalised, -> staging
#pragma unroll 1
    for (int q = 0; q < DPAD / 16; ++q) {
      float o[16];
      tmem_ld16(lane_base + COL_O + q * 16, o);
#pragma unroll
      for (int i = 0; i < 16; i += 4)
        if (q * 16 + i < STG_STRIDE)
          *reinterpret_cast<float4*>(dst_row + q * 16 + i) = make_float4(o[i] * inv_l, o[i + 1] * inv_l, o[i + 2] * inv_l, o[i + 3] * inv_l);
    }
  };
  if (KIND == Q2C) {
    float* stg = reinterpret_cast<float*>(smem);                // 128 x 204 fp32 = 104448 B over X, P and the stages
    __syncthreads();
    drain(stg + row * STG_STRIDE);
    __syncthreads();
#pragma unroll 1
    for (int r = warp; r < TX; r += NW) {                       // fp32 T rows: a warp writes one row (512 + 288 contiguous bytes)
      if (x0 + r >= a.LX) break;
      float* trow = a.out + ((size_t)b * a.LX + x0 + r) * d;
      for (int c4 = lane; c4 < dv4; c4 += 32)
        *reinterpret_cast<float4*>(trow + c4 * 4) = *reinterpret_cast<const float4*>(stg + r * STG_STRIDE + c4 * 4);
    }
    // packed bf16 T (value operand of the C2QB blocks): one 16-byte chunk per (row, chunk), contiguous per 8-row group
    char* tp = reinterpret_cast<char*>(a.t_pack) + x_off;
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
    // blocks 1..3 of the concat (attention.py:52); block 0 (the text itself) was written by the pack kernel.  The
    // products take c from the bf16 text tile that a last TMA brought into the X buffer (global loads here cost ~9 us
    // of latency per block), so the fp32 staging holds 64 rows at a time, over P and the stages.  A warp instruction
    // stores 512 contiguous bytes of ONE row: measured 30 B/clk/SM against 15 for 64-byte runs over 8 rows
    // (tools/micro/store_rate.cu); the price is a bank-conflicted 8-byte read of the core-matrix text tile.
    float* stg = reinterpret_cast<float*>(Ps);                  // 64 x 204 fp32 = 52224 B <= P + stages (61440 B)
    mbar_wait(bar_x, 1);
#pragma unroll 1
    for (int hr = 0; hr < 2; ++hr) {
      __syncthreads();
      if ((warp >> 1) == hr) drain(stg + (row - 64 * hr) * STG_STRIDE);
      __syncthreads();
      if (cta_times && tid == 0) cta_times[8 + 2 * hr] = globaltimer_ns();
#pragma unroll 2
      for (int r = 64 * hr + warp; r < 64 * hr + 64; r += NW) {
        if (x0 + r >= a.LX) break;
        float* orow = a.out + ((size_t)b * a.LX + x0 + r) * 4 * d;
        const unsigned char* ctile = Xs + (r >> 3) * GROUP_BYTES + (r & 7) * 16;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int c4 = lane + 32 * h;
          if (c4 >= dv4) continue;
          const float4 v = *reinterpret_cast<const float4*>(stg + (r - 64 * hr) * STG_STRIDE + c4 * 4);
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
      if (cta_times && tid == 0) cta_times[9 + 2 * hr] = globaltimer_ns();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, TMEM_ALLOC);
  if (KIND == Q2C && ready && tid == 0) signal_counter(ready + b);   // after the barrier: every thread's T stores are ordered before it
}

__global__ void __launch_bounds__(NTHREADS, 2) bidaf_tc2_kernel(const FusedArgs f) {
  const int blk = blockIdx.x;
  long long* times = f.cta_times ? f.cta_times + 12 * (size_t)blk : nullptr;   // debugging aid: [start, loop end, end, SM id, prologue end, X landed,
                                                                               //  first S tile done, tiles]
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

constexpr size_t SMEM_BYTES = (size_t)X_BYTES + P_BYTES + 4 * Y_BYTES + (2 + 2 * MAX_STAGES) * 8 + 16;
static_assert(2 * (SMEM_BYTES + 1024) <= 228 * 1024, "two blocks per SM");
static_assert(TX * STG_STRIDE * 4 <= X_BYTES + P_BYTES + 4 * Y_BYTES, "staging fits over the operands");
static_assert(64 * STG_STRIDE * 4 <= P_BYTES + 4 * Y_BYTES, "half staging fits next to the text tile");

}  // namespace

// The main launch of the bf16 tier (after bidaf_pack_kernel): Q2C, C2QA and C2QB blocks, two per SM.
int bidaf_fwd_tc2_launch(const BidafPacks& pk, const float* bias, float* out, float* q2c, float* bm, float* lse_row,
                         float* lse_col, int B, int Lc, int Lq, int d, cudaStream_t stream) {
  static_assert(PACK_ROWS == TX, "pack padding must match the X tile");
  const int LcP = pk.LcP, LqP = pk.LqP;
  FusedArgs f{};
  // Q2C: X = modality rows (qs), Y = text rows: S operand cw, values cp
  f.k[Q2C] = BlockArgs{pk.qs, {pk.cw, pk.cp}, nullptr, pk.c_words, bias, q2c, pk.tp, lse_col, nullptr, Lq, LqP, Lc, LcP, d};
  // C2QA: X = text rows (cw), Y = modality rows: S operand qs, values qp (the same pack without dropout)
  f.k[C2QA] = BlockArgs{pk.cw, {pk.qs, pk.qp != pk.qs ? pk.qp : nullptr}, pk.cp, pk.q_words, bias, out, nullptr, lse_row, nullptr,
                        Lc, LcP, Lq, LqP, d};
  // C2QB: the same S, values = packed T
  f.k[C2QB] = BlockArgs{pk.cw, {pk.qs, pk.tp}, pk.cp, pk.q_words, bias, out, nullptr, nullptr, bm, Lc, LcP, Lq, LqP, d};
  f.ready = pk.ready;
  f.nq = LqP / TX;
  f.nc = LcP / TX;
  f.n_q2c = B * f.nq;
  f.n_c2q = B * f.nc;
  const char* ct = getenv("MMB_BIDAF_FWD_CTA_TIMES");
  f.cta_times = ct ? reinterpret_cast<long long*>(strtoull(ct, nullptr, 0)) : nullptr;
  MMB_CUDA(cudaMemsetAsync(pk.ready, 0, sizeof(int) * (size_t)B, stream));
  MMB_CUDA(cudaFuncSetAttribute(bidaf_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
  bidaf_tc2_kernel<<<f.n_q2c + 2 * f.n_c2q, NTHREADS, SMEM_BYTES, stream>>>(f);
  return check_launch("bidaf_tc2_kernel");
}

}  // namespace mmb
