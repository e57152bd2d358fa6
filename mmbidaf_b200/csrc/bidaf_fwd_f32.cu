// Fused BiDAF attention forward, fp32 tier (FFMA contractions, rel <= 1e-5 against the oracle).
//
// Replaces layers/attention.py:37-75 of the reference.  The similarity matrix
//     S[i][j] = c~_i.w_c + q~_j.w_q + (c~_i o w_cq).q~_j + bias          (attention.py:70-73)
// is produced tile by tile in shared memory and consumed at once; it never reaches HBM.
// Both soft-maxes are "flash" style streaming soft-maxes (running max / running sum), so one
// generic pass serves all three contractions:
//
//   pass Q2C  X = modality rows j, Y = text rows i   : T[j]  = sum_i softmax_i(S)[i][j] c_i
//   pass C2Q  X = text rows i,     Y = modality rows : a[i]  = sum_j softmax_j(S)[i][j] q_j
//   pass C2QB X = text rows i,     Y = modality rows : b[i]  = sum_j softmax_j(S)[i][j] T_j
//
// (b = s1 (s2^T c) is the re-associated form of attention.py:50; it avoids the (B,Lc,Lc) matrix.)
// A CTA owns TX rows of X and streams TY-row tiles of Y.  w_cq is folded into the X tile, so in
// eval mode the Y tile serves both as the S operand and as the value operand.
//
// Masking follows attention.py:94 literally: a masked logit is the literal -1e30, so a fully
// masked soft-max degenerates to a uniform distribution exactly like the reference.
#include "common.cuh"

namespace mmb {
namespace {

constexpr int TX = 64;    // X rows per CTA
constexpr int TY = 32;    // Y rows per streamed tile
constexpr int NT = 256;   // threads per CTA
constexpr int PS = TY + 4;

enum PassKind { kQ2C = 0, kC2Q = 1, kC2QB = 2 };

struct PassArgs {
  const float* x_feat;      // (B, LX, D)
  const float* y_feat;      // (B, LY, D)  S operand (before dropout)
  const float* v_feat;      // (B, LY, D)  value operand for kC2QB (T); unused otherwise
  const uint8_t* x_keep;    // nullable (B, LX, D)
  const uint8_t* y_keep;    // nullable (B, LY, D)
  const uint8_t* y_mask;    // (B, LY)
  const float* w_x;         // (D) weight whose dot with x~ gives the x-side additive term
  const float* w_y;         // (D)
  const float* w_cross;     // (D)
  const float* bias;        // (1)
  float keep_scale;
  float* out;               // kQ2C: T (B,LX,D);  kC2Q/kC2QB: out (B,LX,4D)
  float* lse;               // kQ2C: lse_col (B,LX); kC2Q: lse_row (B,LX); kC2QB: unused
  float* bm;                // kC2QB: optional (B,LX,D) copy of b = s1 T for the backward pass
  int LX, LY, D;
};

struct PassPair {           // blockIdx.z selects one of up to two passes sharing a launch
  PassArgs p[2];
  int kind[2];
};

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// Warp-per-row tile loader.  Stores x~ (optionally scaled by `scale_vec` afterwards) and returns
// the dot of the *dropped* row with `w` through `term`.  Rows past `rows_valid` are zero filled.
template <int ROWS>
__device__ __forceinline__ void load_tile(float* __restrict__ dst, float* __restrict__ dst_raw, float* __restrict__ term,
                                          const float* __restrict__ src, const uint8_t* __restrict__ keep,
                                          float keep_scale, const float* __restrict__ w,
                                          const float* __restrict__ fold, int row0, int rows_total, int D, int DS) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int dv4 = D >> 2;
  for (int r = warp; r < ROWS; r += NT / 32) {
    const int g = row0 + r;
    float dot = 0.f;
    for (int c4 = lane; c4 < dv4; c4 += 32) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (g < rows_total) {
        v = ld4(src + (size_t)g * D + c4 * 4);
        if (dst_raw) *reinterpret_cast<float4*>(dst_raw + r * DS + c4 * 4) = v;
        if (keep) {
          const uchar4 k = *reinterpret_cast<const uchar4*>(keep + (size_t)g * D + c4 * 4);
          v.x = k.x ? v.x * keep_scale : 0.f;
          v.y = k.y ? v.y * keep_scale : 0.f;
          v.z = k.z ? v.z * keep_scale : 0.f;
          v.w = k.w ? v.w * keep_scale : 0.f;
        }
      } else if (dst_raw) {
        *reinterpret_cast<float4*>(dst_raw + r * DS + c4 * 4) = v;
      }
      const float4 ww = ld4(w + c4 * 4);
      dot += v.x * ww.x + v.y * ww.y + v.z * ww.z + v.w * ww.w;
      if (fold) {
        const float4 f = ld4(fold + c4 * 4);
        v.x *= f.x; v.y *= f.y; v.z *= f.z; v.w *= f.w;
      }
      *reinterpret_cast<float4*>(dst + r * DS + c4 * 4) = v;
    }
    dot = warp_sum(dot);
    if (lane == 0) term[r] = dot;
  }
}

template <int NS>
__global__ void __launch_bounds__(NT) bidaf_pass_f32(const PassPair pp) {
  const PassArgs& a = pp.p[blockIdx.z];
  const int kind = pp.kind[blockIdx.z];
  const int D = a.D, DS = D + 4, dv4 = D >> 2;
  const bool sep_v = (kind == kC2QB) || (a.y_keep != nullptr);

  extern __shared__ __align__(16) float smem[];
  float* Xs = smem;
  float* Ys = Xs + TX * DS;
  float* Vs = sep_v ? Ys + TY * DS : Ys;
  float* Ps = Vs + TY * DS;
  float* xterm = Ps + TX * PS;
  float* yterm = xterm + TX;
  float* alpha_s = yterm + TY;
  float* l_s = alpha_s + TX;

  const int b = blockIdx.y;
  const int x0 = blockIdx.x * TX;
  const int tid = threadIdx.x;
  const float* xg = a.x_feat + (size_t)b * a.LX * D;
  const float* yg = a.y_feat + (size_t)b * a.LY * D;
  const float* vg = (kind == kC2QB) ? a.v_feat + (size_t)b * a.LY * D : nullptr;
  const uint8_t* xk = a.x_keep ? a.x_keep + (size_t)b * a.LX * D : nullptr;
  const uint8_t* yk = a.y_keep ? a.y_keep + (size_t)b * a.LY * D : nullptr;
  const uint8_t* ym = a.y_mask + (size_t)b * a.LY;
  const float bias = a.bias[0];

  load_tile<TX>(Xs, nullptr, xterm, xg, xk, a.keep_scale, a.w_x, a.w_cross, x0, a.LX, D, DS);

  // S-phase mapping: 8 lanes share an X row pair, each lane owns 4 Y columns.
  const int sx = tid >> 3;          // rows sx, sx + 32
  const int sy = tid & 7;           // cols sy + 8 c
  float run_m[2] = {-INFINITY, -INFINITY};
  float run_l[2] = {0.f, 0.f};
  // PV-phase mapping: 16 row threads x 16 column threads; rows pr + 16 r, vec4 columns pc + 16 s.
  const int pr = tid >> 4, pc = tid & 15;
  float4 acc[4][NS];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int s = 0; s < NS; ++s) acc[r][s] = make_float4(0.f, 0.f, 0.f, 0.f);

  for (int y0 = 0; y0 < a.LY; y0 += TY) {
    if (kind == kC2QB) {
      load_tile<TY>(Ys, nullptr, yterm, yg, yk, a.keep_scale, a.w_y, nullptr, y0, a.LY, D, DS);
      // value tile T: plain copy
      const int warp = tid >> 5, lane = tid & 31;
      for (int r = warp; r < TY; r += NT / 32)
        for (int c4 = lane; c4 < dv4; c4 += 32) {
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (y0 + r < a.LY) v = ld4(vg + (size_t)(y0 + r) * D + c4 * 4);
          *reinterpret_cast<float4*>(Vs + r * DS + c4 * 4) = v;
        }
    } else {
      load_tile<TY>(Ys, sep_v ? Vs : nullptr, yterm, yg, yk, a.keep_scale, a.w_y, nullptr, y0, a.LY, D, DS);
    }
    __syncthreads();

    // ---- S tile (64 x 32) and streaming soft-max over y -------------------------------------
    float s[2][4];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) s[r][c] = 0.f;
    {
      const float* xr0 = Xs + sx * DS;
      const float* xr1 = Xs + (sx + 32) * DS;
      const float* yr = Ys + sy * DS;
      for (int k4 = 0; k4 < dv4; ++k4) {
        const float4 xa = ld4(xr0 + k4 * 4), xb = ld4(xr1 + k4 * 4);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float4 yv = ld4(yr + c * 8 * DS + k4 * 4);
          s[0][c] += xa.x * yv.x + xa.y * yv.y + xa.z * yv.z + xa.w * yv.w;
          s[1][c] += xb.x * yv.x + xb.y * yv.y + xb.z * yv.z + xb.w * yv.w;
        }
      }
    }
    bool in_range[4], unmasked[4];
    float yt[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int y = y0 + sy + 8 * c;
      in_range[c] = y < a.LY;
      unmasked[c] = in_range[c] && ym[y] != 0;
      yt[c] = yterm[sy + 8 * c];
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int row = sx + 32 * r;
      const float xt = xterm[row];
      float tile_m = -INFINITY;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        // grouping mirrors attention.py:73 (s0 + s1 + s2 + bias); masked -> literal -1e30
        const float v = unmasked[c] ? ((xt + yt[c]) + s[r][c]) + bias : kNegFill;
        s[r][c] = v;
        if (in_range[c]) tile_m = fmaxf(tile_m, v);
      }
      tile_m = fmaxf(tile_m, __shfl_xor_sync(0xffffffffu, tile_m, 1));
      tile_m = fmaxf(tile_m, __shfl_xor_sync(0xffffffffu, tile_m, 2));
      tile_m = fmaxf(tile_m, __shfl_xor_sync(0xffffffffu, tile_m, 4));
      const float new_m = fmaxf(run_m[r], tile_m);
      const float alpha = expf(run_m[r] - new_m);          // 0 on the first tile (run_m = -inf)
      float tile_l = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float p = in_range[c] ? expf(s[r][c] - new_m) : 0.f;
        tile_l += p;
        Ps[row * PS + sy + 8 * c] = p;
      }
      tile_l += __shfl_xor_sync(0xffffffffu, tile_l, 1);
      tile_l += __shfl_xor_sync(0xffffffffu, tile_l, 2);
      tile_l += __shfl_xor_sync(0xffffffffu, tile_l, 4);
      run_l[r] = run_l[r] * alpha + tile_l;
      run_m[r] = new_m;
      if (sy == 0) alpha_s[row] = alpha;
    }
    __syncthreads();

    // ---- acc[x][:] = acc * alpha + P[x][y] V[y][:] -----------------------------------------------
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const float al = alpha_s[pr + 16 * r];
#pragma unroll
      for (int s2 = 0; s2 < NS; ++s2) {
        acc[r][s2].x *= al; acc[r][s2].y *= al; acc[r][s2].z *= al; acc[r][s2].w *= al;
      }
    }
#pragma unroll 2
    for (int y4 = 0; y4 < TY / 4; ++y4) {
      float4 p[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) p[r] = ld4(Ps + (pr + 16 * r) * PS + y4 * 4);
#pragma unroll
      for (int yy = 0; yy < 4; ++yy) {
        const float* vrow = Vs + (y4 * 4 + yy) * DS;
#pragma unroll
        for (int s2 = 0; s2 < NS; ++s2) {
          const int c4 = pc + 16 * s2;
          if (s2 < NS - 1 || c4 < dv4) {
            const float4 v = ld4(vrow + c4 * 4);
#pragma unroll
            for (int r = 0; r < 4; ++r) {
              const float pv = yy == 0 ? p[r].x : yy == 1 ? p[r].y : yy == 2 ? p[r].z : p[r].w;
              acc[r][s2].x += pv * v.x; acc[r][s2].y += pv * v.y; acc[r][s2].z += pv * v.z; acc[r][s2].w += pv * v.w;
            }
          }
        }
      }
    }
    __syncthreads();
  }

  // ---- epilogue -----------------------------------------------------------------------------------
  if (sy == 0) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int row = sx + 32 * r;
      l_s[row] = run_l[r];
      if (kind != kC2QB && x0 + row < a.LX) a.lse[(size_t)b * a.LX + x0 + row] = run_m[r] + logf(run_l[r]);
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int row = pr + 16 * r;
    const int gx = x0 + row;
    if (gx >= a.LX) continue;
    const float inv = 1.0f / l_s[row];
#pragma unroll
    for (int s2 = 0; s2 < NS; ++s2) {
      const int c4 = pc + 16 * s2;
      if (c4 >= dv4) continue;
      float4 v = acc[r][s2];
      v.x *= inv; v.y *= inv; v.z *= inv; v.w *= inv;
      if (kind == kQ2C) {
        *reinterpret_cast<float4*>(a.out + ((size_t)b * a.LX + gx) * D + c4 * 4) = v;
      } else {
        const float4 c = ld4(xg + (size_t)gx * D + c4 * 4);            // un-dropped text row (attention.py:52)
        float* o = a.out + ((size_t)b * a.LX + gx) * 4 * D + c4 * 4;
        const float4 cv = make_float4(c.x * v.x, c.y * v.y, c.z * v.z, c.w * v.w);
        if (kind == kC2Q) {
          *reinterpret_cast<float4*>(o) = c;
          *reinterpret_cast<float4*>(o + D) = v;
          *reinterpret_cast<float4*>(o + 2 * D) = cv;
        } else {
          *reinterpret_cast<float4*>(o + 3 * D) = cv;
          if (a.bm) *reinterpret_cast<float4*>(a.bm + ((size_t)b * a.LX + gx) * D + c4 * 4) = v;
        }
      }
    }
  }
}

size_t pass_smem_bytes(int D, bool sep_v) {
  const int DS = D + 4;
  return sizeof(float) * ((size_t)TX * DS + (size_t)TY * DS * (sep_v ? 2 : 1) + TX * PS + TX + TY + TX + TX);
}

int launch_pass(const PassPair& pp, int npass, int B, cudaStream_t stream) {
  const int D = pp.p[0].D;
  bool sep_v = false;
  for (int i = 0; i < npass; ++i) sep_v = sep_v || pp.kind[i] == kC2QB || pp.p[i].y_keep != nullptr;
  const size_t smem = pass_smem_bytes(D, sep_v);
  const int ns = (D / 4 + 15) / 16;
  dim3 grid((pp.p[0].LX + TX - 1) / TX, B, npass), block(NT);
  auto go = [&](auto kernel) -> int {
    MMB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kernel<<<grid, block, smem, stream>>>(pp);
    return check_launch("bidaf_pass_f32");
  };
  switch (ns) {
    case 1: return go(bidaf_pass_f32<1>);
    case 2: return go(bidaf_pass_f32<2>);
    case 3: return go(bidaf_pass_f32<3>);
    case 4: return go(bidaf_pass_f32<4>);
  }
  set_error("bidaf: d=%d unsupported", D);
  return MMB_ERR_UNSUPPORTED;
}

}  // namespace

int bidaf_fwd_f32(const float* text, const float* modality, const uint8_t* text_mask, const uint8_t* modality_mask,
                  const float* w_text, const float* w_modality, const float* w_cross, const float* bias,
                  const uint8_t* keep_text, const uint8_t* keep_modality, float keep_scale, float* out, float* q2c,
                  float* bm, float* lse_row, float* lse_col, int B, int Lc, int Lq, int d, cudaStream_t stream) {
  PassPair q{};
  q.kind[0] = kQ2C;
  q.p[0] = PassArgs{modality, text, nullptr, keep_modality, keep_text, text_mask, w_modality, w_text, w_cross, bias,
                    keep_scale, q2c, lse_col, nullptr, Lq, Lc, d};
  int rc = launch_pass(q, 1, B, stream);
  if (rc) return rc;
  PassPair c{};
  c.kind[0] = kC2Q;
  c.p[0] = PassArgs{text, modality, nullptr, keep_text, keep_modality, modality_mask, w_text, w_modality, w_cross, bias,
                    keep_scale, out, lse_row, nullptr, Lc, Lq, d};
  c.kind[1] = kC2QB;
  c.p[1] = c.p[0];
  c.p[1].v_feat = q2c;
  c.p[1].bm = bm;
  return launch_pass(c, 2, B, stream);
}

}  // namespace mmb
