// Mask / length plumbing on the device: ONE launch turns an int32 lengths vector into
//   * the boolean position mask (B, L): pos < len            -- models.py:86-92 `get_mask` (built on the CPU there, then copied)
//   * the decoder mask (B, M): the same mask zero-padded to M  -- models.py:119-123 (zeros + type cast + cat on the CPU)
//   * the longest-first scheduling order of the recurrences    -- stable: rank = #{longer} + #{equal, earlier}
// The hidden-state row permutation of quirk Q3 (encoding.py:91: torch.sort on a CPU float tensor) is NOT taken from this order:
// that sort is not stable and its tie order comes from the host's SIMD sorting network (x86-simd-sort; measured here: it differs from
// the stable order on 51 % of random length lists with ties), so the only bit-exact source is the reference's own call on the host.
#include "common.cuh"

namespace mmb {
namespace {

__global__ void __launch_bounds__(256) length_plan_kernel(const int32_t* __restrict__ lengths, uint8_t* __restrict__ mask,
                                                          uint8_t* __restrict__ dec_mask, int32_t* __restrict__ order, int B, int L,
                                                          int M) {
  const long long n_mask = (long long)B * L, n_dec = dec_mask ? (long long)B * M : 0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_mask + n_dec; i += stride) {
    if (i < n_mask) {
      const int b = (int)(i / L), pos = (int)(i - (long long)b * L);
      mask[i] = pos < lengths[b];
    } else {
      const long long j = i - n_mask;
      const int b = (int)(j / M), pos = (int)(j - (long long)b * M);
      dec_mask[j] = pos < L && pos < lengths[b];
    }
  }
  if (order && blockIdx.x == 0) {
    for (int i = threadIdx.x; i < B; i += blockDim.x) {
      const int li = lengths[i];
      int rank = 0;
      for (int j = 0; j < B; ++j) {
        const int lj = lengths[j];
        rank += (lj > li) || (lj == li && j < i);
      }
      order[rank] = i;
    }
  }
}

}  // namespace
}  // namespace mmb

extern "C" int mmb_length_plan(const int32_t* lengths, uint8_t* mask, uint8_t* dec_mask, int32_t* order, int B, int L, int M,
                               mmb_stream_t stream) {
  MMB_REQUIRE(lengths && B > 0 && L >= 0 && (mask || L == 0), MMB_ERR_INVALID, "mmb_length_plan: bad arguments");
  MMB_REQUIRE(!dec_mask || M >= L, MMB_ERR_INVALID, "mmb_length_plan: decoder mask width %d < mask width %d", M, L);
  const long long work = (long long)B * L + (dec_mask ? (long long)B * M : 0);
  long long blocks = (work + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > 1184) blocks = 1184;
  mmb::length_plan_kernel<<<(int)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(lengths, mask, dec_mask, order, B, L, M);
  return mmb::check_launch("length_plan_kernel");
}
