// Shared helpers for the mmbidaf_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/mmbidaf_b200.h"

namespace mmb {

void set_error(const char* fmt, ...);

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return MMB_ERR_CUDA;
  }
  return MMB_OK;
}

#define MMB_REQUIRE(cond, code, ...)      \
  do {                                    \
    if (!(cond)) {                        \
      ::mmb::set_error(__VA_ARGS__);      \
      return code;                        \
    }                                     \
  } while (0)

#define MMB_CUDA(call)                                                        \
  do {                                                                        \
    cudaError_t e_ = (call);                                                  \
    if (e_ != cudaSuccess) {                                                  \
      ::mmb::set_error("%s failed: %s", #call, cudaGetErrorString(e_));       \
      return MMB_ERR_CUDA;                                                    \
    }                                                                         \
  } while (0)

constexpr float kNegFill = -1e30f;   // layers/attention.py:94

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.0f / (1.0f + expf(-x)); }
// Branch-free gate non-linearities for the serial LSTM chain: act_k(x) = 1 - k / (1 + e^{kx}) is tanh for
// k = 2 and the logistic sigmoid for k = 1.  __expf / __fdividef keep the absolute error near 1e-7
// (measured against the fp64 oracle in tests/test_lstm_gpu.py), which is what the gates need.
__device__ __forceinline__ float gate_act(float x, float k) { return 1.0f - __fdividef(k, 1.0f + __expf(k * x)); }
__device__ __forceinline__ float tanh_fast(float x) { return gate_act(x, 2.0f); }

}  // namespace mmb
