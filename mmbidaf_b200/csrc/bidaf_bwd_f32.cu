// Fused BiDAF attention backward, fp32 tier (FFMA contractions; the 1e-5 companion of bidaf_fwd_f32.cu).
//
// Same closed form and the same three passes as the tensor-core tier (bidaf_bwd_tc.cu has the derivation):
//   prep   per text row: dA = G1 + c o G2, dBm = c o G3, Drow = dA.A + dBm.Bm, dc <- G0 + A o G2 + Bm o G3
//   PT     X = 32 modality rows, streams text tiles:  P^T from lse_row;  dq <- P^T dA,  dT = P^T dBm,  Dcol = dT.T
//   DC     X = 32 text rows, streams modality tiles:  S, dP = dA q^T + dBm T^T, dR = c dT^T -> dS;
//          acc0 = dS q~, acc1 = R dT;  dc += keep o (rowsum(dS) w_c + acc0 o w_cq) + acc1;  partials of dw_c, dw_cq, dbias
//   DQ     the transposed pass for dq and dw_q
//   reduce fixed-order sum of the per-CTA partials
// S, the soft-maxes and dS are rebuilt per 32 x 32 tile in shared memory from the saved log-sum-exp vectors; nothing of
// size Lc x Lq is stored.  Operand tiles are fp32 rows in shared memory (row stride d + 4: conflict-free float4 reads).
#include "common.cuh"

namespace mmb {
namespace {

constexpr int TR = 32;          // tile rows, both sides
constexpr int NT = 256;
constexpr int PS = TR + 1;

enum Mode { PT = 0, DC = 1, DQ = 2 };

struct Operand {
  const float* p;               // (B, L, d)
  const uint8_t* keep;          // nullable (B, L, d): dropout keep mask of the similarity inputs
  const float* fold;            // nullable (d): text_modality_weight folded into the text-side S operand
  const float* w_term;          // nullable (d): weight whose dot with the (dropped) row is the additive term
};

struct F32Args {
  Operand x[4], y[4];
  const uint8_t* x_mask;        // (B, LX)
  const uint8_t* y_mask;        // (B, LY)
  const float* bias;
  const float* norm_x;          // DC/DQ: lse of the soft-max that runs along Y, per X row
  const float* norm_y;          // lse of the soft-max that runs along X, per Y row
  const float* dlt_x;           // DC/DQ: Drow / Dcol per X row
  const float* dlt_y;
  const float* w_term;          // DC/DQ epilogue: (d)
  const float* w_fold;          // DC: (d); DQ: null
  const float* t_feat;          // PT: T (B, LX, d)
  float* dx;                    // PT: d_modality written; DC/DQ: accumulated
  float* dt;                    // PT: dT (B, LX, d)
  float* d_col;                 // PT: (B, LX)
  float* part;                  // DC/DQ: (B, nxb, 3, d) partials [sum x~ rs | sum x~ acc0 | sum rs]
  float keep_scale;
  int LX, LY, d;
};

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// Warp-per-row loader of TR rows; rows past `L` are zero.  term[r] = dot(dropped row, w_term).
__device__ __forceinline__ void load_tile(float* __restrict__ dst, float* __restrict__ term, const Operand& op, size_t batch_off,
                                          float keep_scale, int row0, int L, int d, int DS) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, dv4 = d >> 2;
  for (int r = warp; r < TR; r += NT / 32) {
    const int g = row0 + r;
    float dot = 0.f;
    for (int c4 = lane; c4 < dv4; c4 += 32) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (g < L) {
        v = ld4(op.p + batch_off + (size_t)g * d + c4 * 4);
        if (op.keep) {
          const uchar4 k = *reinterpret_cast<const uchar4*>(op.keep + batch_off + (size_t)g * d + c4 * 4);
          v.x = k.x ? v.x * keep_scale : 0.f;
          v.y = k.y ? v.y * keep_scale : 0.f;
          v.z = k.z ? v.z * keep_scale : 0.f;
          v.w = k.w ? v.w * keep_scale : 0.f;
        }
        if (op.w_term) {
          const float4 w = ld4(op.w_term + c4 * 4);
          dot += v.x * w.x + v.y * w.y + v.z * w.z + v.w * w.w;
        }
        if (op.fold) {
          const float4 f = ld4(op.fold + c4 * 4);
          v.x *= f.x; v.y *= f.y; v.z *= f.z; v.w *= f.w;
        }
      }
      *reinterpret_cast<float4*>(dst + r * DS + c4 * 4) = v;
    }
    if (term && op.w_term) {
      dot = warp_sum(dot);
      if (lane == 0) term[r] = dot;
    }
  }
}

// out[c] (+)= sum_k A[ty][k] B[tx + 8c][k]
__device__ __forceinline__ void tile_dot(const float* __restrict__ A, const float* __restrict__ Bt, int DS, int dv4, int ty, int tx,
                                         float out[4]) {
  const float* ar = A + ty * DS;
  for (int k4 = 0; k4 < dv4; ++k4) {
    const float4 a = ld4(ar + k4 * 4);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float4 b = ld4(Bt + (tx + 8 * c) * DS + k4 * 4);
      out[c] += a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
    }
  }
}

template <int MODE>
__global__ void __launch_bounds__(NT) bidaf_bwd_f32_kernel(const F32Args a) {
  constexpr bool IS_PT = MODE == PT;
  constexpr int NX = IS_PT ? 1 : 4, NY = IS_PT ? 3 : 4;
  const int d = a.d, DS = d + 4, dv4 = d >> 2;
  extern __shared__ __align__(16) float smem[];
  float* Xs = smem;                                   // [NX][TR][DS]
  float* Ys = Xs + NX * TR * DS;                      // [NY][TR][DS]
  float* Ps = Ys + NY * TR * DS;                      // [TR][PS]  P^T (PT) or dS (DC/DQ)
  float* Rs = Ps + TR * PS;                           // [TR][PS]  R (DC)
  float* xterm = Rs + TR * PS;                        // [TR]
  float* yterm = xterm + TR;                          // [TR]
  float* rowsum = yterm + TR;                         // [TR]  DC/DQ: sum over y of the non-vanishing part of dS
  float* red = rowsum + TR;                           // [8][2][d] + [8]: epilogue partials

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.y, x0 = blockIdx.x * TR;
  const size_t xoff = (size_t)b * a.LX * d, yoff = (size_t)b * a.LY * d;
  const float bias = a.bias[0];

#pragma unroll
  for (int p = 0; p < NX; ++p) load_tile(Xs + p * TR * DS, p == 0 ? xterm : nullptr, a.x[p], xoff, a.keep_scale, x0, a.LX, d, DS);
  if (tid < TR) rowsum[tid] = 0.f;

  // tile-product mapping: row ty, columns tx + 8c
  const int ty = tid >> 3, tx = tid & 7;
  const int gx = x0 + ty;
  const bool valid_x = gx < a.LX;
  const bool open_x = valid_x && a.x_mask[(size_t)b * a.LX + gx] != 0;
  float n1 = 0.f, d1 = 0.f;
  if (!IS_PT && valid_x) {
    n1 = a.norm_x[(size_t)b * a.LX + gx];
    d1 = a.dlt_x[(size_t)b * a.LX + gx];
  }
  const float log_lx = logf((float)a.LX);
  // second-product mapping: float4 column pc, rows pr0 + 4r
  const int pc = tid & 63, pr0 = tid >> 6;
  float4 acc0[8], acc1[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) acc0[r] = acc1[r] = make_float4(0.f, 0.f, 0.f, 0.f);
  float rs_part = 0.f;

  for (int y0 = 0; y0 < a.LY; y0 += TR) {
    __syncthreads();                                   // previous tile's readers are done
#pragma unroll
    for (int p = 0; p < NY; ++p) load_tile(Ys + p * TR * DS, p == 0 ? yterm : nullptr, a.y[p], yoff, a.keep_scale, y0, a.LY, d, DS);
    __syncthreads();

    float s[4] = {0.f, 0.f, 0.f, 0.f}, ga[4] = {0.f, 0.f, 0.f, 0.f}, gb[4] = {0.f, 0.f, 0.f, 0.f};
    tile_dot(Xs, Ys, DS, dv4, ty, tx, s);
    if (!IS_PT) {
      // DC: GA = dA q^T + dBm T^T (x1 y1 + x2 y2), GB = c dT^T (x3 y3);  DQ: GA = dT c^T (x1 y1), GB = q dA^T + T dBm^T (x2 y2 + x3 y3)
      tile_dot(Xs + 1 * TR * DS, Ys + 1 * TR * DS, DS, dv4, ty, tx, ga);
      tile_dot(Xs + 2 * TR * DS, Ys + 2 * TR * DS, DS, dv4, ty, tx, MODE == DC ? ga : gb);
      tile_dot(Xs + 3 * TR * DS, Ys + 3 * TR * DS, DS, dv4, ty, tx, gb);
    }
    const float xt = xterm[ty];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int yl = tx + 8 * c, gy = y0 + yl;
      const bool valid_y = gy < a.LY;
      const bool open_y = valid_y && a.y_mask[(size_t)b * a.LY + gy] != 0;
      const float logit = ((xt + yterm[yl]) + s[c]) + bias;            // grouping of attention.py:73
      float lse_y = 0.f, fill = kNegFill;
      if (valid_y) {
        lse_y = a.norm_y[(size_t)b * a.LY + gy];
        // a fully masked soft-max is uniform (attention.py:94): every logit is the fill value and logit - lse = -log(LX),
        // which fp32 cannot hold next to 1e30 -- re-base both to 0
        if (lse_y < -5e29f) {
          lse_y = log_lx;
          fill = 0.f;
        }
      }
      if (IS_PT) {
        Ps[ty * PS + yl] = (valid_x && valid_y) ? expf((open_x ? logit : fill) - lse_y) : 0.f;
      } else {
        const float w2 = (valid_x && valid_y) ? expf((open_x ? logit : fill) - lse_y) : 0.f;     // R (column soft-max weight)
        const float t1 = (valid_x && open_y) ? expf(logit - n1) * (ga[c] - d1) : 0.f;
        const float t2 = (open_x && valid_y) ? w2 * (gb[c] - a.dlt_y[(size_t)b * a.LY + gy]) : 0.f;
        Ps[ty * PS + yl] = t1 + t2;
        if (MODE == DC) Rs[ty * PS + yl] = w2;
        rs_part += t2;                 // the W1 part sums to zero along y analytically
      }
    }
    __syncthreads();
    // acc0 += tile0 . V0,  acc1 += tile1 . V1     (PT: V0 = dA (y1), V1 = dBm (y2); DC: V0 = q~ (y0), V1 = dT (y3) with R; DQ: V0 = y0)
    if (pc < dv4) {
      const float* v0 = Ys + (IS_PT ? 1 : 0) * TR * DS + pc * 4;
      const float* v1 = Ys + (IS_PT ? 2 : 3) * TR * DS + pc * 4;
      const float* t1p = IS_PT ? Ps : Rs;
#pragma unroll 4
      for (int y = 0; y < TR; ++y) {
        const float4 va = ld4(v0 + y * DS);
        float4 vb = make_float4(0.f, 0.f, 0.f, 0.f);
        if (MODE != DQ) vb = ld4(v1 + y * DS);
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const float p = Ps[(pr0 + 4 * r) * PS + y];
          acc0[r].x = fmaf(p, va.x, acc0[r].x); acc0[r].y = fmaf(p, va.y, acc0[r].y);
          acc0[r].z = fmaf(p, va.z, acc0[r].z); acc0[r].w = fmaf(p, va.w, acc0[r].w);
          if (MODE != DQ) {
            const float q = t1p[(pr0 + 4 * r) * PS + y];
            acc1[r].x = fmaf(q, vb.x, acc1[r].x); acc1[r].y = fmaf(q, vb.y, acc1[r].y);
            acc1[r].z = fmaf(q, vb.z, acc1[r].z); acc1[r].w = fmaf(q, vb.w, acc1[r].w);
          }
        }
      }
    }
  }

  // ---- epilogue ---------------------------------------------------------------------------------------------------------
  if (IS_PT) {
    if (pc < dv4) {
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int g = x0 + pr0 + 4 * r;
        if (g >= a.LX) continue;
        *reinterpret_cast<float4*>(a.dx + xoff + (size_t)g * d + pc * 4) = acc0[r];
        *reinterpret_cast<float4*>(a.dt + xoff + (size_t)g * d + pc * 4) = acc1[r];
      }
    }
    __syncthreads();                                   // this CTA's dT rows are visible to its own threads below
    for (int r = warp; r < TR; r += NT / 32) {         // Dcol_j = dT_j . T_j
      const int g = x0 + r;
      if (g >= a.LX) break;
      float dot = 0.f;
      for (int c4 = lane; c4 < dv4; c4 += 32) {
        const float4 t = ld4(a.t_feat + xoff + (size_t)g * d + c4 * 4), v = ld4(a.dt + xoff + (size_t)g * d + c4 * 4);
        dot += t.x * v.x + t.y * v.y + t.z * v.z + t.w * v.w;
      }
      dot = warp_sum(dot);
      if (lane == 0) a.d_col[(size_t)b * a.LX + g] = dot;
    }
  } else {
    // rowsum over the 8 column-threads of a row, accumulated over all tiles
    rs_part += __shfl_xor_sync(0xffffffffu, rs_part, 1);
    rs_part += __shfl_xor_sync(0xffffffffu, rs_part, 2);
    rs_part += __shfl_xor_sync(0xffffffffu, rs_part, 4);
    __syncthreads();
    if (tx == 0) rowsum[ty] = rs_part;
    __syncthreads();
    float pt[4] = {0.f, 0.f, 0.f, 0.f}, pf[4] = {0.f, 0.f, 0.f, 0.f}, p_sum = 0.f;
    if (pc < dv4) {
      float wt[4], wf[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        wt[e] = a.w_term[pc * 4 + e];
        wf[e] = a.w_fold ? a.w_fold[pc * 4 + e] : 1.f;
      }
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int row = pr0 + 4 * r, g = x0 + row;
        if (g >= a.LX) continue;
        const size_t gi = xoff + (size_t)g * d + pc * 4;
        const float rs = rowsum[row];
        const float4 xv = ld4(a.x[0].p + gi);
        float4 o = ld4(a.dx + gi);
        uchar4 kv = make_uchar4(1, 1, 1, 1);
        if (a.x[0].keep) kv = *reinterpret_cast<const uchar4*>(a.x[0].keep + gi);
        const float x[4] = {xv.x, xv.y, xv.z, xv.w}, v0[4] = {acc0[r].x, acc0[r].y, acc0[r].z, acc0[r].w};
        const float v1[4] = {acc1[r].x, acc1[r].y, acc1[r].z, acc1[r].w};
        const unsigned char kk[4] = {kv.x, kv.y, kv.z, kv.w};
        float ov[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float ks = a.x[0].keep ? (kk[e] ? a.keep_scale : 0.f) : 1.f;
          const float xd = x[e] * ks;
          ov[e] += ks * fmaf(rs, wt[e], v0[e] * wf[e]) + (MODE == DC ? v1[e] : 0.f);
          pt[e] = fmaf(xd, rs, pt[e]);
          pf[e] = fmaf(xd, v0[e], pf[e]);
        }
        if (pc == 0) p_sum += rs;
        *reinterpret_cast<float4*>(a.dx + gi) = make_float4(ov[0], ov[1], ov[2], ov[3]);
      }
      *reinterpret_cast<float4*>(red + (pr0 * 2 + 0) * d + pc * 4) = make_float4(pt[0], pt[1], pt[2], pt[3]);
      *reinterpret_cast<float4*>(red + (pr0 * 2 + 1) * d + pc * 4) = make_float4(pf[0], pf[1], pf[2], pf[3]);
      if (pc == 0) red[4 * 2 * d + pr0] = p_sum;
    }
    __syncthreads();
    if (tid < d) {                                     // fixed-order sum over the four row groups
      float s_term = 0.f, s_fold = 0.f;
      for (int g = 0; g < 4; ++g) {
        s_term += red[(g * 2 + 0) * d + tid];
        s_fold += red[(g * 2 + 1) * d + tid];
      }
      float* part = a.part + ((size_t)b * gridDim.x + blockIdx.x) * 3 * d;
      part[tid] = s_term;
      part[d + tid] = s_fold;
      if (tid == 0) part[2 * d] = red[8 * d + 0] + red[8 * d + 1] + red[8 * d + 2] + red[8 * d + 3];
    }
  }
}

// prep: one warp per text row
struct PrepArgs {
  const float *grad, *text, *out, *bm;
  float *da, *dbm, *d_text, *d_row;
  long long rows;             // B * Lc
  int d;
};

__global__ void __launch_bounds__(NT) bidaf_bwd_f32_prep_kernel(const PrepArgs a) {
  const long long row = (long long)blockIdx.x * (NT / 32) + (threadIdx.x >> 5);
  if (row >= a.rows) return;
  const int lane = threadIdx.x & 31, d = a.d, dv4 = d >> 2;
  const float* g = a.grad + row * 4 * d;
  float dot = 0.f;
  for (int c4 = lane; c4 < dv4; c4 += 32) {
    const float4 g0 = ld4(g + c4 * 4), g1 = ld4(g + d + c4 * 4), g2 = ld4(g + 2 * d + c4 * 4), g3 = ld4(g + 3 * d + c4 * 4);
    const float4 c = ld4(a.text + row * d + c4 * 4), av = ld4(a.out + row * 4 * d + d + c4 * 4), bv = ld4(a.bm + row * d + c4 * 4);
    const float4 da = make_float4(fmaf(c.x, g2.x, g1.x), fmaf(c.y, g2.y, g1.y), fmaf(c.z, g2.z, g1.z), fmaf(c.w, g2.w, g1.w));
    const float4 db = make_float4(c.x * g3.x, c.y * g3.y, c.z * g3.z, c.w * g3.w);
    const float4 dc = make_float4(fmaf(bv.x, g3.x, fmaf(av.x, g2.x, g0.x)), fmaf(bv.y, g3.y, fmaf(av.y, g2.y, g0.y)),
                                  fmaf(bv.z, g3.z, fmaf(av.z, g2.z, g0.z)), fmaf(bv.w, g3.w, fmaf(av.w, g2.w, g0.w)));
    dot += da.x * av.x + da.y * av.y + da.z * av.z + da.w * av.w + db.x * bv.x + db.y * bv.y + db.z * bv.z + db.w * bv.w;
    *reinterpret_cast<float4*>(a.da + row * d + c4 * 4) = da;
    *reinterpret_cast<float4*>(a.dbm + row * d + c4 * 4) = db;
    *reinterpret_cast<float4*>(a.d_text + row * d + c4 * 4) = dc;
  }
  dot = warp_sum(dot);
  if (lane == 0) a.d_row[row] = dot;
}

struct ReduceArgs {
  const float *part_c, *part_q;   // (B * nxb, 3, d)
  float *d_w_text, *d_w_cross, *d_w_modality, *d_bias;
  int n_c, n_q, d;
};

__global__ void __launch_bounds__(NT) bidaf_bwd_f32_reduce_kernel(const ReduceArgs a) {
  const int which = blockIdx.x, k = threadIdx.x;     // 0: dw_text, 1: dw_cross, 2: dw_modality, 3: dbias
  if (which < 3) {
    if (k >= a.d) return;
    const float* part = which == 2 ? a.part_q : a.part_c;
    const int n = which == 2 ? a.n_q : a.n_c, off = which == 1 ? a.d : 0;
    float acc = 0.f;
    for (int i = 0; i < n; ++i) acc += part[(size_t)i * 3 * a.d + off + k];
    (which == 0 ? a.d_w_text : which == 1 ? a.d_w_cross : a.d_w_modality)[k] = acc;
  } else {
    __shared__ float red[NT];
    float acc = 0.f;
    for (int i = k; i < a.n_c; i += NT) acc += a.part_c[(size_t)i * 3 * a.d + 2 * a.d];
    red[k] = acc;
    __syncthreads();
    for (int s = NT / 2; s > 0; s >>= 1) {
      if (k < s) red[k] += red[k + s];
      __syncthreads();
    }
    if (k == 0) a.d_bias[0] = red[0];
  }
}

template <int MODE>
size_t smem_bytes(int d) {
  const int NX = MODE == PT ? 1 : 4, NY = MODE == PT ? 3 : 4;
  return sizeof(float) * ((size_t)(NX + NY) * TR * (d + 4) + 2 * TR * PS + 3 * TR + 8 * d + 8);
}

struct Workspace {
  float *da, *dbm, *dt, *d_row, *d_col, *part_c, *part_q;
  size_t bytes;
};
Workspace workspace(void* ws, int B, int Lc, int Lq, int d) {
  char* base = static_cast<char*>(ws);
  size_t off = 0;
  auto take = [&](size_t n_floats) {
    float* p = reinterpret_cast<float*>(base + off);
    off += (n_floats * 4 + 255) / 256 * 256;
    return p;
  };
  Workspace w;
  w.da = take((size_t)B * Lc * d);
  w.dbm = take((size_t)B * Lc * d);
  w.dt = take((size_t)B * Lq * d);
  w.d_row = take((size_t)B * Lc);
  w.d_col = take((size_t)B * Lq);
  w.part_c = take((size_t)B * ((Lc + TR - 1) / TR) * 3 * d);
  w.part_q = take((size_t)B * ((Lq + TR - 1) / TR) * 3 * d);
  w.bytes = off;
  return w;
}

template <int MODE>
int launch(const F32Args& a, int B, cudaStream_t stream, const char* what) {
  const size_t smem = smem_bytes<MODE>(a.d);
  MMB_REQUIRE(smem <= 227 * 1024, MMB_ERR_UNSUPPORTED, "mmb_bidaf_bwd (fp32 tier): d=%d needs %zu B of shared memory", a.d, smem);
  MMB_CUDA(cudaFuncSetAttribute(bidaf_bwd_f32_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  bidaf_bwd_f32_kernel<MODE><<<dim3((a.LX + TR - 1) / TR, B), NT, smem, stream>>>(a);
  return check_launch(what);
}

}  // namespace

size_t bidaf_bwd_f32_workspace_bytes(int B, int Lc, int Lq, int d) { return workspace(nullptr, B, Lc, Lq, d).bytes; }

int bidaf_bwd_f32(const float* grad_out, const float* text, const float* modality, const uint8_t* text_mask,
                  const uint8_t* modality_mask, const float* w_text, const float* w_modality, const float* w_cross,
                  const float* bias, const uint8_t* keep_text, const uint8_t* keep_modality, float keep_scale, const float* out,
                  const float* bm, const float* q2c, const float* lse_row, const float* lse_col, void* ws, float* d_text,
                  float* d_modality, float* d_w_text, float* d_w_modality, float* d_w_cross, float* d_bias, int B, int Lc, int Lq,
                  int d, cudaStream_t stream) {
  MMB_REQUIRE(d % 4 == 0 && d <= 256, MMB_ERR_UNSUPPORTED, "mmb_bidaf_bwd (fp32 tier): d=%d (need d %% 4 == 0, d <= 256)", d);
  MMB_REQUIRE(ws && text_mask && modality_mask, MMB_ERR_INVALID, "mmb_bidaf_bwd (fp32 tier): workspace or mask is null");
  const Workspace w = workspace(ws, B, Lc, Lq, d);
  const long long rows = (long long)B * Lc;
  PrepArgs pa{grad_out, text, out, bm, w.da, w.dbm, d_text, w.d_row, rows, d};
  bidaf_bwd_f32_prep_kernel<<<(unsigned)((rows + NT / 32 - 1) / (NT / 32)), NT, 0, stream>>>(pa);
  if (int rc = check_launch("bidaf_bwd_f32_prep_kernel")) return rc;

  const Operand text_s{text, keep_text, w_cross, w_text}, text_p{text, nullptr, nullptr, nullptr};
  const Operand mod_s{modality, keep_modality, nullptr, w_modality}, mod_p{modality, nullptr, nullptr, nullptr};
  const Operand dA{w.da, nullptr, nullptr, nullptr}, dBm{w.dbm, nullptr, nullptr, nullptr};
  const Operand T{q2c, nullptr, nullptr, nullptr}, dT{w.dt, nullptr, nullptr, nullptr};
  {   // PT: X = modality rows, Y = text rows
    F32Args a{};
    a.x[0] = mod_s;
    a.y[0] = text_s; a.y[1] = dA; a.y[2] = dBm;
    a.x_mask = modality_mask; a.y_mask = text_mask; a.bias = bias;
    a.norm_y = lse_row;
    a.t_feat = q2c; a.dx = d_modality; a.dt = w.dt; a.d_col = w.d_col;
    a.keep_scale = keep_scale; a.LX = Lq; a.LY = Lc; a.d = d;
    if (int rc = launch<PT>(a, B, stream, "bidaf_bwd_f32_kernel<PT>")) return rc;
  }
  {   // DC: X = text rows, Y = modality rows
    F32Args a{};
    a.x[0] = text_s; a.x[1] = dA; a.x[2] = dBm; a.x[3] = text_p;
    a.y[0] = mod_s; a.y[1] = mod_p; a.y[2] = T; a.y[3] = dT;
    a.x_mask = text_mask; a.y_mask = modality_mask; a.bias = bias;
    a.norm_x = lse_row; a.norm_y = lse_col; a.dlt_x = w.d_row; a.dlt_y = w.d_col;
    a.w_term = w_text; a.w_fold = w_cross; a.dx = d_text; a.part = w.part_c;
    a.keep_scale = keep_scale; a.LX = Lc; a.LY = Lq; a.d = d;
    if (int rc = launch<DC>(a, B, stream, "bidaf_bwd_f32_kernel<DC>")) return rc;
  }
  {   // DQ: X = modality rows, Y = text rows
    F32Args a{};
    a.x[0] = mod_s; a.x[1] = dT; a.x[2] = mod_p; a.x[3] = T;
    a.y[0] = text_s; a.y[1] = text_p; a.y[2] = dA; a.y[3] = dBm;
    a.x_mask = modality_mask; a.y_mask = text_mask; a.bias = bias;
    a.norm_x = lse_col; a.norm_y = lse_row; a.dlt_x = w.d_col; a.dlt_y = w.d_row;
    a.w_term = w_modality; a.w_fold = nullptr; a.dx = d_modality; a.part = w.part_q;
    a.keep_scale = keep_scale; a.LX = Lq; a.LY = Lc; a.d = d;
    if (int rc = launch<DQ>(a, B, stream, "bidaf_bwd_f32_kernel<DQ>")) return rc;
  }
  ReduceArgs ra{w.part_c, w.part_q, d_w_text, d_w_cross, d_w_modality, d_bias, B * ((Lc + TR - 1) / TR), B * ((Lq + TR - 1) / TR), d};
  bidaf_bwd_f32_reduce_kernel<<<4, NT, 0, stream>>>(ra);
  return check_launch("bidaf_bwd_f32_reduce_kernel");
}

}  // namespace mmb
