// Fused BiDAF attention forward, tensor-core tier, cut 4: ONE persistent, warp-specialised CTA per SM.
//
// Why (profiles/r02_bidaf_fwd.md): in cuts 1-3 a thread block is one serial chain -- load X, tile loop, drain, store -- run
// by 4-8 warps, every phase latency bound (a 64-column tile took ~2 us against 0.45 us of tensor time), and all blocks
// of a launch run in lock step, so the chip alternates between "everybody computes" and "everybody stores" (the store phases
// hit the ~5.5 TB/s the chip can write, the loops leave memory idle).  Here the phases of DIFFERENT work items overlap
// inside one CTA that lives for the whole launch:
//
//   warps 0-7   soft-max      two threads per X row (16 columns of a 32-column S tile each): masked streaming soft-max
//                             (base 2, lazy rescale), P as bf16 into shared memory
//   warps 8-15  epilogue      drain the FINISHED item's accumulator (TMEM -> registers -> per-warp transpose in shared
//                             memory -> coalesced 128-byte runs) while the next item's tile loop runs
//   warp 16     MMA issuer    S(t+1) = X Y^T is issued BEFORE P V(t): the tensor pipe computes the next S under the soft-max
//   warp 17     scheduler+TMA work items from a global atomic queue; Y tiles (S operand + value operand) through a 4-stage ring
//   warp 18     X loader      the next item's X tile is requested the moment the current item's last S product has completed
//                             (its own warp, so that the Y ring is never held up behind that wait)
//
// TMEM (512 columns): S x 2 (32 columns each) | O x 2 (208 columns each): the accumulator of item n is drained while item
// n + 1 accumulates into the other one.  Work items are the blocks of cut 2 (bidaf_fwd_tc2.cu):
//   Q2C  X = 128 modality rows, streams text tiles      T = softmax_i(S)^T c      -> T fp32 (optional), packed bf16 T, lse_col
//   C2QA X = 128 text rows, streams modality tiles      a = softmax_j(S) q        -> out blocks 1, 2; lse_row
//   C2QB X = 128 text rows, streams modality tiles      b = softmax_j(S) T        -> out block 3 (and bm)
// in that order in the queue; a C2QB item waits for its batch row's Q2C items (ready[b]), which were claimed earlier by CTAs
// that are running -- no assumption on block scheduling.  The c * a / c * b products read c from the fp32 text (coalesced,
// overlapped), so no plain text tile is staged.
//
// Barrier protocol (all mbarriers; "k-th use" parities; every wait is bounded -> a protocol bug traps instead of hanging):
//   item_full/empty[4]   scheduler -> all roles: {kind, b, xblk, tiles}
//   x_full/free          TMA -> MMA (tx bytes) / MMA commit after the item's last S product
//   y_full/free[NST]     TMA -> MMA (tx bytes) / MMA commit after P V(t)
//   s_full/free[2]       MMA commit after S(t) / soft-max warps after their tcgen05.ld of S(t)
//   p_full/free[2]       soft-max warps after writing P(t) / MMA commit after P V(t)
//   o_full/free[2]       MMA commit after the item's last P V / epilogue warps after their last tcgen05.ld of it
//   st_full/free[2]      soft-max warps: row statistics (reference maximum, partial sums) -> epilogue warps
#include <stdlib.h>
#include "tc_common.cuh"

namespace mmb {
using namespace tc;
namespace {

constexpr int TX = 128, TY = 32;
constexpr int X_BYTES = TX / 8 * GROUP_BYTES;   // 53248
constexpr int Y_BYTES = TY / 8 * GROUP_BYTES;   // 13312
constexpr int STAGE_BYTES = 2 * Y_BYTES;        // S operand + value operand
constexpr int NST = 4;
constexpr int P_BYTES = TX * TY * 2;            // 8192: chunk c8 (8 columns) at c8 * 2048 + row * 16
constexpr int EPI_COLS = 32, EPI_STRIDE = 36;   // per-warp transpose buffer: 32 rows x 32 columns fp32, padded
constexpr int EPI_WARP_BYTES = 32 * EPI_STRIDE * 4;
constexpr int NSOFT = 8, NEPI = 8;
constexpr int EPI_WARP0 = 8, MMA_WARP = 16, TMA_WARP = 17, XLOAD_WARP = 18, NTHREADS = 19 * 32;
constexpr int NSB = 3;                          // S buffers in TMEM = P buffers in shared memory
constexpr int COL_S = 0, COL_O = NSB * TY;      // S: 3 x 32 columns; O: 2 x 208 columns (96 + 416 = 512)
constexpr int ITEM_SLOTS = 4;
constexpr float TAU2 = 11.0f;
constexpr float NEG2 = kNegFill * LOG2E;

enum Kind { Q2C = 0, C2QA = 1, C2QB = 2, DONE = 3 };

struct KindArgs {
  const __nv_bfloat16* x_pack;       // S operand of the X side
  const __nv_bfloat16* s_pack;       // S operand of the Y side
  const __nv_bfloat16* v_pack;       // value operand of the Y side (may equal s_pack)
  const unsigned long long* y_words; // (B, LYP/64, 2)
  const float* c_src;                // C2QA / C2QB: fp32 text (B, LX, d) for the products
  float* out;                        // Q2C: T fp32 (B, LX, d) or null;  C2QA / C2QB: out (B, LX, 4d)
  __nv_bfloat16* t_pack;             // Q2C: packed T
  float* lse;                        // Q2C: lse_col; C2QA: lse_row; C2QB: null
  float* bm;                         // C2QB: optional (B, LX, d)
  int LX, LXP, LY, LYP;
};

struct Args {
  KindArgs k[3];
  const float* bias;
  int* ready;            // (B) Q2C -> C2QB counters (epilogue warps of finished Q2C items), zeroed before the launch
  int* queue;            // work queue head, zeroed before the launch
  int nq, nc;            // X blocks per batch row: modality side, text side
  int n_q2c, n_c2q;      // B * nq, B * nc
  int d;
  long long* trace;      // debugging aid: 8 x int64 per item, or null
  int dbg;               // debugging aid (MMB_TC4_DEBUG): 1 no global stores, 2 no c loads, 4 no TMEM drain, 8 no transpose
};

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void st_cs_f4(float* p, const float4 v) {      // streaming store: the output is not re-read here
  asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__device__ __forceinline__ int4 ld_item(const int4* p) {                   // (ordered by the acquire of the barrier wait before it)
  int4 v;
  asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(smem_u32(p)) : "memory");
  return v;
}

// One count on a global dependency counter with release semantics: this thread's earlier writes, and those it has observed through
// the __syncwarp before it, are visible to whoever acquires the counter.
__device__ __forceinline__ void signal_release(int* counter) {
  asm volatile("red.release.gpu.global.add.s32 [%0], 1;" ::"l"(counter) : "memory");
}

struct Smem {
  unsigned char* X;        // X_BYTES
  unsigned char* Y;        // NST x STAGE_BYTES
  unsigned char* P;        // NSB x P_BYTES
  unsigned char* E;        // NEPI x EPI_WARP_BYTES
  float* xbuf;             // [2][2][TX] row-max exchange between the two threads of a row
  float* stat;             // [2][3][TX] m_ref, l (half 0), l (half 1)
  int4* items;             // [ITEM_SLOTS]
  uint64_t* bars;
  uint32_t* tmem_slot;
};
constexpr int NBARS = 2 * ITEM_SLOTS + 2 + 2 + 2 * NST + 4 * NSB + 2 + 2 + 2 + 2;
constexpr size_t SMEM_BYTES = (size_t)X_BYTES + NST * STAGE_BYTES + NSB * P_BYTES + NEPI * EPI_WARP_BYTES + 2 * 2 * TX * 4 +
                              2 * 3 * TX * 4 + ITEM_SLOTS * 16 + NBARS * 8 + 16;
static_assert(SMEM_BYTES <= 227 * 1024, "one CTA per SM");

// Drain one finished accumulator (this warp's 32 rows, every second 32-column chunk) to global memory.  KIND is a compile-time
// constant so that each item kind gets straight-line code with its pointers in registers: the first version selected the kind at
// run time inside the unrolled loops and spent 420 instructions per chunk, 13 us per item, on address arithmetic and predicates.
template <int KIND>
__device__ __forceinline__ void epilogue_item(const Args& f, const uint32_t acc, float* ebuf, const int b, const int xblk, const int q4,
                                              const int eh, const int lane, const float inv_l, const uint32_t o_free_bar) {
  const KindArgs& a = f.k[KIND];
  const int d = f.d;
  const int x0 = xblk * TX;
  const int wrows = max(0, min(32, a.LX - x0 - q4 * 32));           // valid rows among this warp's 32
  const size_t row0 = (size_t)b * a.LX + x0 + q4 * 32;              // global row of this warp's first row
  const int c4 = (lane & 7) * 4, rsub = lane >> 3;                  // transposed role: 8 lanes x float4 = one row's 32 columns
  const int nvalid = (wrows - rsub + 3) >> 2;                       // rows rsub, rsub + 4, ... below wrows: rr < nvalid
  constexpr int NCHUNK = (DPAD + EPI_COLS - 1) / EPI_COLS;          // 7: columns 0..223, 208 allocated
  const int last_cc = ((NCHUNK - 1 - eh) & ~1) + eh;                // this warp's last chunk
  // running pointers, made once per item: everything below is pointer + small constant (the first version rebuilt 64-bit
  // row offsets for every access: ~35 integer instructions per load)
  const uint32_t ostride = KIND == Q2C ? (uint32_t)d : 4u * (uint32_t)d;      // floats per output row
  float* const obase = a.out ? a.out + (row0 + rsub) * ostride + (KIND == C2QA ? d : KIND == C2QB ? 3 * d : 0) + c4 : nullptr;
  const float* const cbase = KIND != Q2C ? a.c_src + (row0 + rsub) * d + c4 : nullptr;
  float* const bbase = (KIND == C2QB && a.bm) ? a.bm + (row0 + rsub) * d + c4 : nullptr;
  char* const tpack = KIND == Q2C ? reinterpret_cast<char*>(a.t_pack) +
                                        ((size_t)b * (a.LXP / 8) + (size_t)(x0 + q4 * 32) / 8 + (lane >> 3)) * GROUP_BYTES + (lane & 7) * 16
                                  : nullptr;
  const uint32_t oinc = 4u * ostride, cinc = 4u * (uint32_t)d;      // four rows further
#pragma unroll 1
  for (int cc = eh; cc < NCHUNK; cc += 2) {
    const int col0 = cc * EPI_COLS;
    const bool col_ok = col0 + c4 < d;
    uint32_t raw[32];
    if (!(f.dbg & 4)) {
      tmem_ld16_nowait(acc + col0, raw);
      if (cc < NCHUNK - 1) tmem_ld16_nowait(acc + col0 + 16, raw + 16);
      tmem_wait_ld();
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) raw[i] = 0x3f800000u;
    }
    if (cc == NCHUNK - 1) {
#pragma unroll
      for (int i = 16; i < 32; ++i) raw[i] = 0u;
    }
    if (cc == last_cc) {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_free_bar);                       // the accumulator may be overwritten (item n + 2)
    }
    if (KIND == Q2C) {
      // packed bf16 T (value operand of the C2QB items): this thread's 8-column chunks go straight to their place --
      // 8 consecutive rows x 16 bytes = one 128-byte core matrix, so a warp writes whole 128-byte lines
      const bool row_ok = lane < wrows;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int ch = cc * 4 + j;
        if (ch < CHUNKS) {
          __nv_bfloat162 h[4];
#pragma unroll
          for (int e2 = 0; e2 < 4; ++e2) {
            const bool ok = row_ok && ch * 8 + 2 * e2 < d;
            h[e2] = __floats2bfloat162_rn(ok ? __uint_as_float(raw[j * 8 + 2 * e2]) * inv_l : 0.f,
                                          ok ? __uint_as_float(raw[j * 8 + 2 * e2 + 1]) * inv_l : 0.f);
          }
          if (!(f.dbg & 1)) *reinterpret_cast<uint4*>(tpack + ch * 128) = *reinterpret_cast<uint4*>(h);
        }
      }
      if (obase == nullptr) continue;                               // inference: the fp32 T is only saved for the backward pass
    }
    // transpose through this warp's buffer: thread = row  ->  8 lanes x float4 = 128 contiguous bytes of one row
#pragma unroll
    for (int i = 0; i < 32; i += 4)
      *reinterpret_cast<float4*>(ebuf + lane * EPI_STRIDE + i) =
          make_float4(__uint_as_float(raw[i]) * inv_l, __uint_as_float(raw[i + 1]) * inv_l, __uint_as_float(raw[i + 2]) * inv_l,
                      __uint_as_float(raw[i + 3]) * inv_l);
    __syncwarp();
    const int nv = col_ok ? nvalid : 0;
    float4 cv[8];
    if (KIND != Q2C) {                                              // the c values this chunk multiplies (raw's registers are free now)
      const float* cp = cbase + col0;
#pragma unroll
      for (int rr = 0; rr < 8; ++rr) {
        cv[rr] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (rr < nv && !(f.dbg & 2)) cv[rr] = __ldg(reinterpret_cast<const float4*>(cp));
        cp += cinc;
      }
    }
    float* op = obase + col0;
    float* bp = bbase ? bbase + col0 : nullptr;
    const float* erow = ebuf + rsub * EPI_STRIDE + c4;
#pragma unroll
    for (int rr = 0; rr < 8; ++rr) {
      if (rr < nv && !(f.dbg & 1)) {
        const float4 o = *reinterpret_cast<const float4*>(erow + rr * 4 * EPI_STRIDE);
        if (KIND == Q2C) {
          *reinterpret_cast<float4*>(op) = o;                       // re-read by the backward pass: default cache policy
        } else {
          const float4 c = cv[rr];
          const float4 p = make_float4(c.x * o.x, c.y * o.y, c.z * o.z, c.w * o.w);
          if (KIND == C2QA) {
            st_cs_f4(op, o);
            st_cs_f4(op + d, p);
          } else {
            st_cs_f4(op, p);
            if (bp) *reinterpret_cast<float4*>(bp) = o;
          }
        }
      }
      op += oinc;
      if (KIND == C2QB && bp) bp += cinc;
    }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(NTHREADS, 1) bidaf_tc4_kernel(const Args f) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Smem sm;
  sm.X = smem_raw;
  sm.Y = sm.X + X_BYTES;
  sm.P = sm.Y + NST * STAGE_BYTES;
  sm.E = sm.P + NSB * P_BYTES;
  sm.xbuf = reinterpret_cast<float*>(sm.E + NEPI * EPI_WARP_BYTES);
  sm.stat = sm.xbuf + 2 * 2 * TX;
  sm.items = reinterpret_cast<int4*>(sm.stat + 2 * 3 * TX);
  sm.bars = reinterpret_cast<uint64_t*>(sm.items + ITEM_SLOTS);
  sm.tmem_slot = reinterpret_cast<uint32_t*>(sm.bars + NBARS);

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp_u = uniform_warp_idx();
  const uint32_t leader = elect_one();

  // barrier addresses
  const uint32_t b0 = smem_u32(sm.bars);
  const uint32_t item_full0 = b0, item_empty0 = item_full0 + 8 * ITEM_SLOTS;
  const uint32_t x_full0 = item_empty0 + 8 * ITEM_SLOTS, x_free0 = x_full0 + 16;
  const uint32_t y_full0 = x_free0 + 16, y_free0 = y_full0 + 8 * NST;
  const uint32_t s_full0 = y_free0 + 8 * NST, s_free0 = s_full0 + 8 * NSB;
  const uint32_t p_full0 = s_free0 + 8 * NSB, p_free0 = p_full0 + 8 * NSB;
  const uint32_t o_full0 = p_free0 + 8 * NSB, o_free0 = o_full0 + 16;
  const uint32_t st_full0 = o_free0 + 16, st_free0 = st_full0 + 16;

  if (tid == 0) {
    for (int i = 0; i < ITEM_SLOTS; ++i) {
      mbar_init(item_full0 + 8 * i, 1);
      mbar_init(item_empty0 + 8 * i, NSOFT + NEPI + 2);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(x_full0 + 8 * i, 1);
      mbar_init(x_free0 + 8 * i, 1);
      mbar_init(o_full0 + 8 * i, 1);
      mbar_init(o_free0 + 8 * i, NEPI);
      mbar_init(st_full0 + 8 * i, NSOFT);
      mbar_init(st_free0 + 8 * i, NEPI);
    }
    for (int i = 0; i < NST; ++i) {
      mbar_init(y_full0 + 8 * i, 1);
      mbar_init(y_free0 + 8 * i, 1);
    }
    for (int i = 0; i < NSB; ++i) {
      mbar_init(s_full0 + 8 * i, 1);
      mbar_init(s_free0 + 8 * i, NSOFT);
      mbar_init(p_full0 + 8 * i, NSOFT);
      mbar_init(p_free0 + 8 * i, 1);
    }
    fence_barrier_init();
  }
  if (warp_u == 0) tmem_alloc(smem_u32(sm.tmem_slot), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(sm.tmem_slot);
  const int n_items = f.n_q2c + 2 * f.n_c2q;

  if (warp_u == TMA_WARP) {
    // =========================================== scheduler + TMA producer ===========================================
    uint32_t g = 0;                                                 // tiles issued so far (ring position)
    for (uint32_t n = 0;; ++n) {
      int item = 0;
      if (leader) item = atomicAdd(f.queue, 1);
      item = __shfl_sync(0xffffffffu, item, __ffs(__ballot_sync(0xffffffffu, leader)) - 1);
      int kind = DONE, b = 0, xblk = 0, nty = 0;
      int canon = 0;                                                // kind-major item number (trace index)
      if (item < n_items) {
        // Queue order: Q2C and C2QA items interleaved in proportion (neither depends on anything), then the C2QB items.  With all
        // Q2C items first every SM ran the same kind at the same time and the store phases of all SMs coincided.
        const int n1 = f.n_q2c + f.n_c2q;
        if (item < n1) {
          const int qb = (int)((long long)item * f.n_q2c / n1), qa = (int)((long long)(item + 1) * f.n_q2c / n1);
          if (qa > qb) { kind = Q2C; b = qb / f.nq; xblk = qb - b * f.nq; canon = qb; }
          else { kind = C2QA; const int i = item - qb; b = i / f.nc; xblk = i - b * f.nc; canon = f.n_q2c + i; }
        } else {
          kind = C2QB;
          const int i = item - n1;
          b = i / f.nc;
          xblk = i - b * f.nc;
          canon = item;
        }
        const KindArgs& a = f.k[kind];
        // tiles past the last un-masked Y row contribute exp(-1e30 - m) = 0 to every soft-max: stop there.  (If nothing at
        // all is un-masked the soft-max is uniform over the whole range, attention.py:94, and every tile is needed.)
        nty = (a.LY + TY - 1) / TY;
        int last = 0;
        for (int w = lane; w < (a.LY + 63) / 64; w += 32) {
          const unsigned long long open = a.y_words[((size_t)b * (a.LYP / 64) + w) * 2 + 1];
          if (open != 0ull) last = 2 * w + ((open >> 32) != 0ull ? 2 : 1);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) last = max(last, __shfl_xor_sync(0xffffffffu, last, o));
        if (last > 0) nty = min(nty, last);
      }
      const int slot = n & (ITEM_SLOTS - 1);
      mbar_wait(item_empty0 + 8 * slot, ((n / ITEM_SLOTS) & 1) ^ 1);
      if (lane == 0) {
        sm.items[slot] = make_int4(kind, b, xblk, nty);
        if (f.trace && kind != DONE) f.trace[(size_t)canon * 8 + 0] = globaltimer_ns();
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(item_full0 + 8 * slot);            // release semantics: the slot contents are visible
      if (kind == DONE) break;
      const KindArgs& a = f.k[kind];
      const size_t y_batch = (size_t)b * (a.LYP / 8) * GROUP_BYTES;
      if (kind == C2QB) {                                           // the value operand is T: wait for this batch row's Q2C items
        wait_counter(f.ready + b, f.nq * NEPI);
        fence_proxy_async_all();                                    // their generic-proxy stores -> our async-proxy (TMA) loads
      }
      const bool same_v = a.v_pack == a.s_pack;
      for (int t = 0; t < nty; ++t, ++g) {
        const int s = g % NST;
        mbar_wait(y_free0 + 8 * s, ((g / NST) & 1) ^ 1);
        const uint32_t dst = smem_u32(sm.Y + s * STAGE_BYTES);
        const size_t off = y_batch + (size_t)t * Y_BYTES;
        mbar_expect_tx(y_full0 + 8 * s, same_v ? Y_BYTES : STAGE_BYTES, leader);
        tma_bulk_g2s(dst, reinterpret_cast<const char*>(a.s_pack) + off, Y_BYTES, y_full0 + 8 * s, leader);
        if (!same_v) tma_bulk_g2s(dst + Y_BYTES, reinterpret_cast<const char*>(a.v_pack) + off, Y_BYTES, y_full0 + 8 * s, leader);
      }
    }
  } else if (warp_u == MMA_WARP) {
    // ================================================= MMA issuer ==================================================
    constexpr uint32_t IDESC_S = idesc_bf16(TY, 0), IDESC_PV = idesc_bf16(DPAD, 1);
    uint32_t g = 0;
    for (uint32_t n = 0;; ++n) {
      const int slot = n & (ITEM_SLOTS - 1);
      mbar_wait(item_full0 + 8 * slot, (n / ITEM_SLOTS) & 1);
      const int4 it = ld_item(sm.items + slot);
      __syncwarp();
      if (lane == 0) mbar_arrive(item_empty0 + 8 * slot);
      const int kind = it.x, nty = it.w;
      if (kind == DONE) break;
      const bool same_v = f.k[kind].v_pack == f.k[kind].s_pack;
      const int ob = n & 1;
      mbar_wait(x_full0, n & 1);
      tc_fence_after();
      const uint32_t xs_lo = desc_lo(smem_u32(sm.X), 128);
      auto issue_s = [&](int t) {                                   // S(t) = X Y_t^T into S buffer (g + t) % NSB
        const uint32_t gt = g + t;
        const int s = gt % NST, sb = gt % NSB;
        mbar_wait(y_full0 + 8 * s, (gt / NST) & 1);
        mbar_wait(s_free0 + 8 * sb, ((gt / NSB) & 1) ^ 1);
        tc_fence_after();
        const uint32_t y_lo = desc_lo(smem_u32(sm.Y + s * STAGE_BYTES), 128);
#pragma unroll
        for (int k = 0; k < DPAD / 16; ++k)
          umma_bf16_lh(tmem + COL_S + sb * TY, xs_lo + k * 16, desc_hi(GROUP_BYTES), y_lo + k * 16, desc_hi(GROUP_BYTES), IDESC_S,
                       k > 0, leader);
        umma_commit(s_full0 + 8 * sb, leader);
        if (t == nty - 1) umma_commit(x_free0, leader);             // the X tile is dead after the item's last S product
      };
      int s_issued = 0;
      for (int t = 0; t < nty; ++t) {
        // the S products run ahead of the soft-max by up to NSB - 1 tiles: the tensor pipe always has the next S queued, and a
        // soft-max warp that finishes a tile finds the next one waiting
        while (s_issued < nty && s_issued < t + NSB) issue_s(s_issued++);
        const uint32_t gt = g + t;
        const int s = gt % NST, pb = gt % NSB;
        if (t == 0) mbar_wait(o_free0 + 8 * ob, ((n >> 1) & 1) ^ 1);   // the accumulator of item n - 2 has been drained
        mbar_wait(p_full0 + 8 * pb, (gt / NSB) & 1);
        tc_fence_after();
        const uint32_t v_lo = desc_lo(smem_u32(sm.Y + s * STAGE_BYTES + (same_v ? 0 : Y_BYTES)), GROUP_BYTES);
        const uint32_t ps_lo = desc_lo(smem_u32(sm.P + pb * P_BYTES), 2048);
#pragma unroll
        for (int k = 0; k < TY / 16; ++k)                           // O += P V (V MN-major: LBO = group stride)
          umma_bf16_lh(tmem + COL_O + ob * DPAD, ps_lo + k * 256, desc_hi(128), v_lo + k * 2 * GROUP_BYTES / 16, desc_hi(128),
                       IDESC_PV, (t > 0) || (k > 0), leader);
        umma_commit(y_free0 + 8 * s, leader);
        umma_commit(p_free0 + 8 * pb, leader);
        if (t == nty - 1) umma_commit(o_full0 + 8 * ob, leader);
      }
      g += nty;
    }
  } else if (warp_u == XLOAD_WARP) {
    // ================================================== X loader ===================================================
    for (uint32_t n = 0;; ++n) {
      const int slot = n & (ITEM_SLOTS - 1);
      mbar_wait(item_full0 + 8 * slot, (n / ITEM_SLOTS) & 1);
      const int4 it = ld_item(sm.items + slot);
      __syncwarp();
      if (lane == 0) mbar_arrive(item_empty0 + 8 * slot);
      if (it.x == DONE) break;
      const KindArgs& a = f.k[it.x];
      const size_t x_off = ((size_t)it.y * (a.LXP / 8) + (size_t)it.z * (TX / 8)) * GROUP_BYTES;
      mbar_wait(x_free0, (n & 1) ^ 1);                              // the previous item's last S product has read the tile
      mbar_expect_tx(x_full0, X_BYTES, leader);
      tma_bulk_g2s(smem_u32(sm.X), reinterpret_cast<const char*>(a.x_pack) + x_off, X_BYTES, x_full0, leader);
    }
  } else if (warp_u < NSOFT) {
    // ================================================ soft-max warps ================================================
    const int q4 = warp_u & 3, half = warp_u >> 2;
    const int row = q4 * 32 + lane;
    const uint32_t lane_base = tmem + ((uint32_t)(q4 * 32) << 16);
    const float bias2 = f.bias[0] * LOG2E;
    uint32_t g = 0;
    for (uint32_t n = 0;; ++n) {
      const int slot = n & (ITEM_SLOTS - 1);
      mbar_wait(item_full0 + 8 * slot, (n / ITEM_SLOTS) & 1);
      const int4 it = ld_item(sm.items + slot);
      __syncwarp();
      if (lane == 0) mbar_arrive(item_empty0 + 8 * slot);
      const int kind = it.x, b = it.y, nty = it.w;
      if (kind == DONE) break;
      const KindArgs& a = f.k[kind];
      const int ob = n & 1;
      const unsigned long long* words_b = a.y_words + (size_t)b * (a.LYP / 64) * 2;
      float m_ref = -INFINITY, l_part = 0.f;                        // log2 domain; l over this thread's columns
      ulonglong2 words = *reinterpret_cast<const ulonglong2*>(words_b);
      for (int t = 0; t < nty; ++t, ++g) {
        const int sb = g % NSB;
        const uint32_t sk = (g / NSB) & 1;                          // parity of this use of buffer sb
        const uint32_t sh = (t & 1) * 32 + half * 16;
        const uint32_t wvalid = (uint32_t)(words.x >> sh) & 0xffffu, wopen = (uint32_t)(words.y >> sh) & 0xffffu;
        const bool all_open = (wvalid & wopen) == 0xffffu;
        if (t + 1 < nty) words = *reinterpret_cast<const ulonglong2*>(words_b + (size_t)((t + 1) >> 1) * 2);   // next tile's masks
        mbar_wait(s_full0 + 8 * sb, sk);
        tc_fence_after();
        float sv[16];
        tmem_ld16(lane_base + COL_S + sb * TY + half * 16, sv);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(s_free0 + 8 * sb);               // S(t + 2) may overwrite the buffer
        float tile_max = -INFINITY;
        if (all_open) {
          float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            sv[c] = fmaf(sv[c], LOG2E, bias2);
            mx[c & 3] = fmaxf(mx[c & 3], sv[c]);
          }
          tile_max = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
        } else {
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            const float v = ((wopen >> c) & 1u) ? fmaf(sv[c], LOG2E, bias2) : NEG2;   // attention.py:94
            sv[c] = v;
            if ((wvalid >> c) & 1u) tile_max = fmaxf(tile_max, v);
          }
        }
        float* xb = sm.xbuf + (g & 1) * 2 * TX;
        xb[half * TX + row] = tile_max;
        named_bar_sync(1 + q4, 64);                                 // the two warps that share these 32 rows
        tile_max = fmaxf(tile_max, xb[(half ^ 1) * TX + row]);
        float alpha = 1.f;
        const bool bump = tile_max > m_ref + TAU2;                  // first tile: m_ref = -inf -> always
        if (bump) {
          alpha = fast_exp2(m_ref - tile_max);                      // 0 on the first tile
          m_ref = tile_max;
        }
        float psum = 0.f;
        uint32_t packed[8];
        if (all_open) {
          float ps[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int c = 0; c < 16; c += 2) {
            const float p0 = fast_exp2(sv[c] - m_ref), p1 = fast_exp2(sv[c + 1] - m_ref);
            ps[(c >> 1) & 3] += p0 + p1;
            const __nv_bfloat162 pk = __floats2bfloat162_rn(p0, p1);
            packed[c / 2] = *reinterpret_cast<const uint32_t*>(&pk);
          }
          psum = (ps[0] + ps[1]) + (ps[2] + ps[3]);
        } else {
#pragma unroll
          for (int c = 0; c < 16; c += 2) {
            const float p0 = ((wvalid >> c) & 1u) ? fast_exp2(sv[c] - m_ref) : 0.f;
            const float p1 = ((wvalid >> (c + 1)) & 1u) ? fast_exp2(sv[c + 1] - m_ref) : 0.f;
            psum += p0 + p1;
            const __nv_bfloat162 pk = __floats2bfloat162_rn(p0, p1);
            packed[c / 2] = *reinterpret_cast<const uint32_t*>(&pk);
          }
        }
        l_part = l_part * alpha + psum;
        mbar_wait(p_free0 + 8 * sb, sk ^ 1);                        // P V(t - NSB) has read this P buffer
        {
          unsigned char* prow = sm.P + sb * P_BYTES + (2 * half) * 2048 + row * 16;
          *reinterpret_cast<uint4*>(prow) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
          *reinterpret_cast<uint4*>(prow + 2048) = make_uint4(packed[4], packed[5], packed[6], packed[7]);
        }
        if (__any_sync(0xffffffffu, bump && t > 0)) {               // lazy rescale of this warp's rows (alpha = 1 where no bump)
          mbar_wait(p_free0 + 8 * ((g - 1) % NSB), ((g - 1) / NSB) & 1);   // P V(t - 1) has landed in the accumulator
          tc_fence_after();
#pragma unroll 1
          for (int qq = half; qq < DPAD / 16; qq += 2) {            // the two threads of a row split the columns
            float o[16];
            tmem_ld16(lane_base + COL_O + ob * DPAD + qq * 16, o);
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] *= alpha;
            tmem_st16(lane_base + COL_O + ob * DPAD + qq * 16, o);
          }
          tmem_wait_st();
          tc_fence_before();
        }
        fence_proxy_async();                                        // st.shared P -> visible to the tensor core
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full0 + 8 * sb);
      }
      // row statistics -> epilogue warps
      mbar_wait(st_free0 + 8 * ob, ((n >> 1) & 1) ^ 1);
      float* st = sm.stat + ob * 3 * TX;
      if (half == 0) st[row] = m_ref;
      st[(1 + half) * TX + row] = l_part;
      __syncwarp();
      if (lane == 0) mbar_arrive(st_full0 + 8 * ob);
    }
  } else if (warp_u < EPI_WARP0 + NEPI) {
    // ================================================ epilogue warps ================================================
    const int e = warp_u - EPI_WARP0, q4 = e & 3, eh = e >> 2;     // two warps per TMEM lane quarter: even / odd column chunks
    const uint32_t lane_base = tmem + ((uint32_t)(q4 * 32) << 16);
    float* ebuf = reinterpret_cast<float*>(sm.E + e * EPI_WARP_BYTES);
    for (uint32_t n = 0;; ++n) {
      const int slot = n & (ITEM_SLOTS - 1);
      mbar_wait(item_full0 + 8 * slot, (n / ITEM_SLOTS) & 1);
      const int4 it = ld_item(sm.items + slot);
      __syncwarp();
      if (lane == 0) mbar_arrive(item_empty0 + 8 * slot);
      const int kind = it.x, b = it.y, xblk = it.z;
      if (kind == DONE) break;
      const int ob = n & 1;
      const int row = q4 * 32 + lane;                               // this thread's accumulator row (TMEM lane)
      mbar_wait(st_full0 + 8 * ob, (n >> 1) & 1);
      const float* st = sm.stat + ob * 3 * TX;
      const float m_ref = st[row], l_run = st[TX + row] + st[2 * TX + row];
      __syncwarp();
      if (lane == 0) mbar_arrive(st_free0 + 8 * ob);
      const float inv_l = 1.f / l_run;
      {
        const KindArgs& a = f.k[kind];
        if (eh == 0 && a.lse && xblk * TX + row < a.LX) a.lse[(size_t)b * a.LX + xblk * TX + row] = (m_ref + log2f(l_run)) * LN2;
      }
      mbar_wait(o_full0 + 8 * ob, (n >> 1) & 1);
      tc_fence_after();
      long long t_epi0 = 0;
      if (f.trace && e == 0 && lane == 0) t_epi0 = globaltimer_ns();
      const uint32_t acc = lane_base + COL_O + ob * DPAD;
      if (kind == Q2C) epilogue_item<Q2C>(f, acc, ebuf, b, xblk, q4, eh, lane, inv_l, o_free0 + 8 * ob);
      else if (kind == C2QA) epilogue_item<C2QA>(f, acc, ebuf, b, xblk, q4, eh, lane, inv_l, o_free0 + 8 * ob);
      else epilogue_item<C2QB>(f, acc, ebuf, b, xblk, q4, eh, lane, inv_l, o_free0 + 8 * ob);
      if (kind == Q2C) {                                            // this warp's part of the T rows is in memory: one count per warp
        __syncwarp();
        if (lane == 0) signal_release(f.ready + b);
      }
      if (f.trace && e == 0 && lane == 0) {
        long long* tr = f.trace + ((size_t)(kind == Q2C ? b * f.nq + xblk
                                            : kind == C2QA ? f.n_q2c + b * f.nc + xblk : f.n_q2c + f.n_c2q + b * f.nc + xblk)) * 8;
        tr[1] = t_epi0;
        tr[2] = globaltimer_ns();
        tr[3] = kind;
        uint32_t smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        tr[4] = smid;
        tr[5] = it.w;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp_u == 0) tmem_dealloc(tmem, TMEM_COLS);
}

}  // namespace

// Same contract as bidaf_fwd_tc2_launch (bidaf_fwd_tc2.cu): after bidaf_pack_kernel, one launch for all Q2C, C2QA and C2QB items.
// `text` is the fp32 text input (B, Lc, d); q2c may be null (inference: T is kept only in its packed bf16 form).
int bidaf_fwd_tc4_launch(const BidafPacks& pk, const float* text, const float* bias, float* out, float* q2c, float* bm,
                         float* lse_row, float* lse_col, int B, int Lc, int Lq, int d, cudaStream_t stream) {
  static_assert(PACK_ROWS == TX, "pack padding must match the X tile");
  const int LcP = pk.LcP, LqP = pk.LqP;
  Args f{};
  f.k[Q2C] = KindArgs{pk.qs, pk.cw, pk.cp, pk.c_words, nullptr, q2c, pk.tp, lse_col, nullptr, Lq, LqP, Lc, LcP};
  f.k[C2QA] = KindArgs{pk.cw, pk.qs, pk.qp, pk.q_words, text, out, nullptr, lse_row, nullptr, Lc, LcP, Lq, LqP};
  f.k[C2QB] = KindArgs{pk.cw, pk.qs, pk.tp, pk.q_words, text, out, nullptr, nullptr, bm, Lc, LcP, Lq, LqP};
  f.bias = bias;
  f.ready = pk.ready;
  f.queue = pk.ready + B;
  f.nq = LqP / TX;
  f.nc = LcP / TX;
  f.n_q2c = B * f.nq;
  f.n_c2q = B * f.nc;
  f.d = d;
  static const char* dbg_env = getenv("MMB_TC4_DEBUG");
  f.dbg = dbg_env ? atoi(dbg_env) : 0;
  static const char* trace_env = getenv("MMB_BIDAF_FWD_ITEM_TRACE");     // debugging aid (tools/bidaf_fwd_timeline.py)
  f.trace = trace_env ? reinterpret_cast<long long*>(strtoull(trace_env, nullptr, 0)) : nullptr;
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    MMB_CUDA(cudaGetDevice(&dev));
    MMB_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    MMB_CUDA(cudaFuncSetAttribute(bidaf_tc4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
  }
  MMB_CUDA(cudaMemsetAsync(pk.ready, 0, sizeof(int) * (size_t)(B + 1), stream));
  const int n_items = f.n_q2c + 2 * f.n_c2q;
  bidaf_tc4_kernel<<<n_items < num_sms ? n_items : num_sms, NTHREADS, SMEM_BYTES, stream>>>(f);
  return check_launch("bidaf_tc4_kernel");
}

}  // namespace mmb
