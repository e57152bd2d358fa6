// Persistent length-aware LSTM recurrence, forward and backward (BPTT), fp32.
//
// Replaces the nn.LSTM call of layers/encoding.py:96 (and the sort / pack / unpack / unsort
// gathers around it, encoding.py:91-101) for one layer, both directions.  The input projection
// x W_ih^T + b_ih + b_hh is a plain GEMM done by the caller; this kernel owns the serial part.
//
// One CTA = one direction x NB sequences, alive for the whole sequence.  Thread (j, kp), with
// j = hidden unit and kp = 0..3, keeps in REGISTERS the recurrent weights of the four gate rows of
// unit j restricted to a quarter of the k range (4 x KS floats), so W_hh never leaves the
// register file between time steps.  Per step: 4 x KS FMAs against h (broadcast float4 reads from
// shared memory), a 3-shuffle reduce-scatter over the 4 kp lanes (lane kp ends with gate kp),
// one activation per lane, a 4-shuffle exchange, the cell update, one __syncthreads.
// Input pre-activations are prefetched RING-1 steps ahead with cp.async.
//
// Semantics follow torch.nn.LSTM on a PackedSequence: gate order i,f,g,o; zero initial state;
// the reverse direction starts at each sample's own last valid step; outputs past a sample's
// length are exactly zero (pad_packed_sequence, encoding.py:99); h_n / c_n are the states after
// each sample's last valid step.
#include "common.cuh"

namespace mmb {
namespace {

constexpr int RING = 8;

__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gsrc) {
  const unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

struct LstmArgs {
  float* gates;            // (B, L, ndir, 4H): in = input pre-activations, out (save) = activated gates / d pre-act
  const float* w_hh;       // (ndir, 4H, H)
  const int* lengths;      // (B)
  const int* order;        // (B) permutation (longest first) or nullptr
  float* out;              // fwd: (B, L, ndir*H)
  float* h_n;              // fwd: (B, ndir, H)
  float* c_n;              // fwd: (B, ndir, H)
  float* cell;             // (B, L, ndir, H) saved cell states (fwd writes when save != 0, bwd reads)
  const float* dout;       // bwd: (B, L, ndir*H)
  const float* dh_n;       // bwd: (B, ndir, H) or nullptr
  const float* dc_n;       // bwd: (B, ndir, H) or nullptr
  int B, L, H, ndir, save;
};

template <int KS, int NB>
__global__ void __launch_bounds__(4 * 4 * KS <= 128 ? 128 : 4 * 4 * KS) bilstm_fwd_kernel(const LstmArgs a) {
  constexpr int HP = 4 * KS;                       // padded hidden size
  const int H = a.H, L = a.L, ndir = a.ndir;
  const int dir = blockIdx.y;
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int j = tid >> 2, kp = tid & 3;
  const bool live = j < H;

  extern __shared__ __align__(16) float smem[];
  float* h_s = smem;                               // [2][NB][HP]
  float* ring = h_s + 2 * NB * HP;                 // [RING][NB][nthr]

  int seq[NB], len[NB];
  int max_len = 0;
#pragma unroll
  for (int n = 0; n < NB; ++n) {
    const int slot = blockIdx.x * NB + n;
    seq[n] = slot < a.B ? (a.order ? a.order[slot] : slot) : -1;
    len[n] = seq[n] >= 0 ? min(max(a.lengths[seq[n]], 0), L) : 0;
    max_len = max(max_len, len[n]);
  }

  // recurrent weights of unit j, gates 0..3, k in [kp*KS, kp*KS + KS)
  float w[4][KS];
  {
    const float* wd = a.w_hh + (size_t)dir * 4 * H * H;
#pragma unroll
    for (int g = 0; g < 4; ++g)
#pragma unroll
      for (int kk = 0; kk < KS; ++kk) {
        const int k = kp * KS + kk;
        w[g][kk] = (live && k < H) ? wd[(size_t)(g * H + j) * H + k] : 0.f;
      }
  }
  for (int i = tid; i < 2 * NB * HP; i += nthr) h_s[i] = 0.f;

  auto gate_ptr = [&](int n, int s) -> float* {    // this lane's pre-activation of sequence n at step s
    const int t = dir ? len[n] - 1 - s : s;
    return a.gates + (((size_t)seq[n] * L + t) * ndir + dir) * 4 * H + kp * H + j;
  };
  auto prefetch = [&](int s) {
    if (live) {
#pragma unroll
      for (int n = 0; n < NB; ++n)
        if (s < len[n]) cp_async4(ring + ((s % RING) * NB + n) * nthr + tid, gate_ptr(n, s));
    }
    cp_async_commit();
  };
#pragma unroll 1
  for (int s = 0; s < RING - 1; ++s) prefetch(s);
  __syncthreads();

  float c_reg[NB], h_reg[NB];
#pragma unroll
  for (int n = 0; n < NB; ++n) c_reg[n] = h_reg[n] = 0.f;
  const int quad = (tid & 31) & ~3;

#pragma unroll 1
  for (int s = 0; s < max_len; ++s) {
    prefetch(s + RING - 1);
    cp_async_wait<RING - 1>();
    const float* hc = h_s + (s & 1) * NB * HP;
    float* hn = h_s + ((s + 1) & 1) * NB * HP;

    float acc[NB][4];
#pragma unroll
    for (int n = 0; n < NB; ++n)
#pragma unroll
      for (int g = 0; g < 4; ++g) acc[n][g] = 0.f;
#pragma unroll
    for (int k4 = 0; k4 < KS / 4; ++k4) {
#pragma unroll
      for (int n = 0; n < NB; ++n) {
        const float4 hv = *reinterpret_cast<const float4*>(hc + n * HP + kp * KS + k4 * 4);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          acc[n][g] = fmaf(w[g][k4 * 4 + 0], hv.x, acc[n][g]);
          acc[n][g] = fmaf(w[g][k4 * 4 + 1], hv.y, acc[n][g]);
          acc[n][g] = fmaf(w[g][k4 * 4 + 2], hv.z, acc[n][g]);
          acc[n][g] = fmaf(w[g][k4 * 4 + 3], hv.w, acc[n][g]);
        }
      }
    }
#pragma unroll
    for (int n = 0; n < NB; ++n) {
      // reduce-scatter over the 4 kp lanes: lane kp ends with the full sum of gate kp
      const bool hi = kp & 2, odd = kp & 1;
      float k0 = hi ? acc[n][2] : acc[n][0], k1 = hi ? acc[n][3] : acc[n][1];
      const float s0 = hi ? acc[n][0] : acc[n][2], s1 = hi ? acc[n][1] : acc[n][3];
      k0 += __shfl_xor_sync(0xffffffffu, s0, 2);
      k1 += __shfl_xor_sync(0xffffffffu, s1, 2);
      float mine = odd ? k1 : k0;
      mine += __shfl_xor_sync(0xffffffffu, odd ? k0 : k1, 1);
      const bool active = s < len[n];                               // uniform over the CTA
      const float pre = mine + ((active && live) ? ring[((s % RING) * NB + n) * nthr + tid] : 0.f);
      const float act = kp == 2 ? tanhf(pre) : sigmoidf_acc(pre);
      const float gi = __shfl_sync(0xffffffffu, act, quad + 0);
      const float gf = __shfl_sync(0xffffffffu, act, quad + 1);
      const float gg = __shfl_sync(0xffffffffu, act, quad + 2);
      const float go = __shfl_sync(0xffffffffu, act, quad + 3);
      if (active && live) {
        c_reg[n] = fmaf(gf, c_reg[n], gi * gg);
        h_reg[n] = go * tanhf(c_reg[n]);
        const int t = dir ? len[n] - 1 - s : s;
        const size_t bt = (size_t)seq[n] * L + t;
        if (kp == 0) {
          hn[n * HP + j] = h_reg[n];
          a.out[bt * ndir * H + dir * H + j] = h_reg[n];
        }
        if (a.save) {
          a.gates[(bt * ndir + dir) * 4 * H + kp * H + j] = act;
          if (kp == 1) a.cell[(bt * ndir + dir) * H + j] = c_reg[n];
        }
      }
    }
    __syncthreads();
  }
  cp_async_wait<0>();

#pragma unroll
  for (int n = 0; n < NB; ++n) {
    if (seq[n] < 0) continue;
    if (live && kp == 0) a.h_n[((size_t)seq[n] * ndir + dir) * H + j] = h_reg[n];
    if (live && kp == 1) a.c_n[((size_t)seq[n] * ndir + dir) * H + j] = c_reg[n];
    // pad_packed_sequence: zeros past the sample's length
    for (int i = tid; i < (L - len[n]) * H; i += nthr) {
      const int t = len[n] + i / H, u = i % H;
      a.out[((size_t)seq[n] * L + t) * ndir * H + dir * H + u] = 0.f;
    }
  }
}

// Backward through time.  `gates` holds the activated gates on entry and d(pre-activation) on exit
// (zeros past each length), so dW_ih / dx / db / dW_hh are plain GEMMs for the caller.
template <int KS, int NB>
__global__ void __launch_bounds__(4 * 4 * KS <= 128 ? 128 : 4 * 4 * KS) bilstm_bwd_kernel(const LstmArgs a) {
  constexpr int HP = 4 * KS;
  const int H = a.H, L = a.L, ndir = a.ndir;
  const int dir = blockIdx.y;
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int j = tid >> 2, kp = tid & 3;
  const bool live = j < H;

  extern __shared__ __align__(16) float smem[];
  float* da_s = smem;                              // [2][NB][4][HP]
  float* ring = da_s + 2 * NB * 4 * HP;            // [RING][NB][2][nthr]

  int seq[NB], len[NB];
  int max_len = 0;
#pragma unroll
  for (int n = 0; n < NB; ++n) {
    const int slot = blockIdx.x * NB + n;
    seq[n] = slot < a.B ? (a.order ? a.order[slot] : slot) : -1;
    len[n] = seq[n] >= 0 ? min(max(a.lengths[seq[n]], 0), L) : 0;
    max_len = max(max_len, len[n]);
  }

  // transposed recurrent weights: rows of gate kp, column j
  float wt[HP];
  {
    const float* wd = a.w_hh + (size_t)dir * 4 * H * H;
#pragma unroll
    for (int r = 0; r < HP; ++r) wt[r] = (live && r < H) ? wd[(size_t)(kp * H + r) * H + j] : 0.f;
  }
  for (int i = tid; i < 2 * NB * 4 * HP; i += nthr) da_s[i] = 0.f;

  // backward step s visits forward step fs = len-1-s, i.e. time t = dir ? s : len-1-s
  auto time_of = [&](int n, int s) { return dir ? s : len[n] - 1 - s; };
  auto prefetch = [&](int s) {
    if (live) {
#pragma unroll
      for (int n = 0; n < NB; ++n)
        if (s < len[n]) {
          const int t = time_of(n, s);
          const size_t bt = (size_t)seq[n] * L + t;
          float* slot = ring + (((s % RING) * NB + n) * 2) * nthr + tid;
          cp_async4(slot, a.gates + (bt * ndir + dir) * 4 * H + kp * H + j);
          if (kp == 0) cp_async4(slot + nthr, a.cell + (bt * ndir + dir) * H + j);
          if (kp == 1) cp_async4(slot + nthr, a.dout + bt * ndir * H + dir * H + j);
          if (kp == 2 && s + 1 < len[n]) {           // cell state of the previous forward step
            const int tp = time_of(n, s + 1);
            cp_async4(slot + nthr, a.cell + (((size_t)seq[n] * L + tp) * ndir + dir) * H + j);
          }
        }
    }
    cp_async_commit();
  };
#pragma unroll 1
  for (int s = 0; s < RING - 1; ++s) prefetch(s);

  float dh_rec[NB], dc[NB];
#pragma unroll
  for (int n = 0; n < NB; ++n) {
    dh_rec[n] = (live && seq[n] >= 0 && a.dh_n) ? a.dh_n[((size_t)seq[n] * ndir + dir) * H + j] : 0.f;
    dc[n] = (live && seq[n] >= 0 && a.dc_n) ? a.dc_n[((size_t)seq[n] * ndir + dir) * H + j] : 0.f;
  }
  const int quad = (tid & 31) & ~3;
  __syncthreads();

#pragma unroll 1
  for (int s = 0; s < max_len; ++s) {
    prefetch(s + RING - 1);
    cp_async_wait<RING - 1>();
    float* dcur = da_s + (s & 1) * NB * 4 * HP;

#pragma unroll
    for (int n = 0; n < NB; ++n) {
      const bool active = s < len[n];
      const float* slot = ring + (((s % RING) * NB + n) * 2) * nthr + tid;
      const float v1 = (active && live) ? slot[0] : 0.f;
      const float v2 = (active && live && (kp < 2 || (kp == 2 && s + 1 < len[n]))) ? slot[nthr] : 0.f;
      const float gi = __shfl_sync(0xffffffffu, v1, quad + 0);
      const float gf = __shfl_sync(0xffffffffu, v1, quad + 1);
      const float gg = __shfl_sync(0xffffffffu, v1, quad + 2);
      const float go = __shfl_sync(0xffffffffu, v1, quad + 3);
      const float ct = __shfl_sync(0xffffffffu, v2, quad + 0);
      const float dy = __shfl_sync(0xffffffffu, v2, quad + 1);
      const float cp = __shfl_sync(0xffffffffu, v2, quad + 2);
      const float dh = dy + dh_rec[n];
      const float tc = tanhf(ct);
      const float dct = fmaf(dh * go, 1.f - tc * tc, dc[n]);
      float da;
      if (kp == 0) da = dct * gg * gi * (1.f - gi);
      else if (kp == 1) da = dct * cp * gf * (1.f - gf);
      else if (kp == 2) da = dct * gi * (1.f - gg * gg);
      else da = dh * tc * go * (1.f - go);
      if (active && live) {
        dc[n] = dct * gf;
        dcur[(n * 4 + kp) * HP + j] = da;
        const size_t bt = (size_t)seq[n] * L + time_of(n, s);
        a.gates[(bt * ndir + dir) * 4 * H + kp * H + j] = da;
      } else if (live) {
        dcur[(n * 4 + kp) * HP + j] = 0.f;
      }
    }
    __syncthreads();
    // dh_rec[j] = sum_r W_hh[r][j] da[r]: this lane covers the rows of gate kp
    float part[NB];
#pragma unroll
    for (int n = 0; n < NB; ++n) part[n] = 0.f;
#pragma unroll
    for (int r4 = 0; r4 < HP / 4; ++r4) {
#pragma unroll
      for (int n = 0; n < NB; ++n) {
        const float4 dv = *reinterpret_cast<const float4*>(dcur + (n * 4 + kp) * HP + r4 * 4);
        part[n] = fmaf(wt[r4 * 4 + 0], dv.x, part[n]);
        part[n] = fmaf(wt[r4 * 4 + 1], dv.y, part[n]);
        part[n] = fmaf(wt[r4 * 4 + 2], dv.z, part[n]);
        part[n] = fmaf(wt[r4 * 4 + 3], dv.w, part[n]);
      }
    }
#pragma unroll
    for (int n = 0; n < NB; ++n) {
      part[n] += __shfl_xor_sync(0xffffffffu, part[n], 1);
      part[n] += __shfl_xor_sync(0xffffffffu, part[n], 2);
      dh_rec[n] = part[n];
    }
  }
  cp_async_wait<0>();

  // d(pre-activation) is zero past each sample's length
#pragma unroll
  for (int n = 0; n < NB; ++n) {
    if (seq[n] < 0) continue;
    for (int i = tid; i < (L - len[n]) * 4 * H; i += nthr) {
      const int t = len[n] + i / (4 * H), u = i % (4 * H);
      a.gates[(((size_t)seq[n] * L + t) * ndir + dir) * 4 * H + u] = 0.f;
    }
  }
}

template <int KS, int NB>
int launch(const LstmArgs& a, bool backward, cudaStream_t stream) {
  constexpr int HP = 4 * KS;
  const int nthr = ((4 * a.H + 31) / 32) * 32;
  dim3 grid((a.B + NB - 1) / NB, a.ndir), block(nthr);
  if (!backward) {
    const size_t smem = sizeof(float) * (2 * NB * HP + (size_t)RING * NB * nthr);
    MMB_CUDA(cudaFuncSetAttribute(bilstm_fwd_kernel<KS, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    bilstm_fwd_kernel<KS, NB><<<grid, block, smem, stream>>>(a);
    return check_launch("bilstm_fwd_kernel");
  }
  const size_t smem = sizeof(float) * (2 * NB * 4 * HP + (size_t)RING * NB * 2 * nthr);
  MMB_CUDA(cudaFuncSetAttribute(bilstm_bwd_kernel<KS, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  bilstm_bwd_kernel<KS, NB><<<grid, block, smem, stream>>>(a);
  return check_launch("bilstm_bwd_kernel");
}

int pick_nb(int B, int ndir) {
  // Fewest sequences per CTA that still fits one wave of 148 SMs (latency-bound recurrence).
  return ((B + 0) * ndir <= 148) ? 1 : 2;
}

int dispatch(const LstmArgs& a, bool backward, cudaStream_t stream) {
  const int nb = pick_nb(a.B, a.ndir);
#define MMB_LSTM_CASE(KS)                                                       \
  return nb == 1 ? launch<KS, 1>(a, backward, stream) : launch<KS, 2>(a, backward, stream)
  if (a.H <= 16) { MMB_LSTM_CASE(4); }
  if (a.H <= 64) { MMB_LSTM_CASE(16); }
  if (a.H <= 112) { MMB_LSTM_CASE(28); }
  if (a.H <= 128) { MMB_LSTM_CASE(32); }
#undef MMB_LSTM_CASE
  set_error("bilstm: hidden size %d > 128 unsupported", a.H);
  return MMB_ERR_UNSUPPORTED;
}

}  // namespace
}  // namespace mmb

extern "C" int mmb_bilstm_fwd(float* gates, const float* w_hh, const int32_t* lengths, const int32_t* order, float* out,
                              float* h_n, float* c_n, float* cell, int B, int L, int H, int ndir, int save,
                              mmb_stream_t stream) {
  MMB_REQUIRE(gates && w_hh && lengths && out && h_n && c_n, MMB_ERR_INVALID, "mmb_bilstm_fwd: null pointer");
  MMB_REQUIRE(!save || cell, MMB_ERR_INVALID, "mmb_bilstm_fwd: save=1 needs a cell buffer");
  MMB_REQUIRE(B > 0 && L > 0 && H > 0 && (ndir == 1 || ndir == 2), MMB_ERR_INVALID,
              "mmb_bilstm_fwd: B=%d L=%d H=%d ndir=%d", B, L, H, ndir);
  mmb::LstmArgs a{gates, w_hh, lengths, order, out, h_n, c_n, cell, nullptr, nullptr, nullptr, B, L, H, ndir, save};
  return mmb::dispatch(a, false, static_cast<cudaStream_t>(stream));
}

extern "C" int mmb_bilstm_bwd(float* gates, const float* cell, const float* w_hh, const int32_t* lengths,
                              const int32_t* order, const float* dout, const float* dh_n, const float* dc_n, int B, int L,
                              int H, int ndir, mmb_stream_t stream) {
  MMB_REQUIRE(gates && cell && w_hh && lengths && dout, MMB_ERR_INVALID, "mmb_bilstm_bwd: null pointer");
  MMB_REQUIRE(B > 0 && L > 0 && H > 0 && (ndir == 1 || ndir == 2), MMB_ERR_INVALID,
              "mmb_bilstm_bwd: B=%d L=%d H=%d ndir=%d", B, L, H, ndir);
  mmb::LstmArgs a{gates, w_hh, lengths, order, nullptr, nullptr, nullptr, const_cast<float*>(cell), dout, dh_n, dc_n,
                  B, L, H, ndir, 1};
  return mmb::dispatch(a, true, static_cast<cudaStream_t>(stream));
}
