// K0 (fast): device negative sampler.  Same job as the rejection loop at
// base/deepRecommender.py:47-49 / model/ranking/BPR.py:35-37, but drawn from a counter-based
// Philox4x32-10 stream so that every triple is independent (the reference's MT19937 stream is
// inherently serial; the bit-exact clone lives in host_sampler.cpp).
//
//   j = (philox(key=(seed_lo, seed_hi); ctr=(k_lo, k_hi, attempt, epoch)).x * num_items) >> 32
// attempt = 0, 1, ... until j is not in user u[k]'s sorted rated-item row (binary search).
#include "common.h"

namespace {

__device__ __forceinline__ uint32_t philox4x32_10_x(uint32_t c0, uint32_t c1, uint32_t c2,
                                                    uint32_t c3, uint32_t k0, uint32_t k1) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  return c0;
}

__global__ void __launch_bounds__(256)
sample_neg_philox_kernel(long long n, int num_items, const int* __restrict__ u,
                         const long long* __restrict__ rowptr, const int* __restrict__ cols,
                         uint32_t seed_lo, uint32_t seed_hi, uint32_t epoch, int* __restrict__ out_j) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
    const int uu = __ldg(u + k);
    const long long lo0 = __ldg(rowptr + uu), hi0 = __ldg(rowptr + uu + 1);
    uint32_t attempt = 0;
    int j;
    while (true) {
      const uint32_t r = philox4x32_10_x((uint32_t)k, (uint32_t)((unsigned long long)k >> 32),
                                         attempt, epoch, seed_lo, seed_hi);
      j = (int)(((unsigned long long)r * (unsigned long long)(uint32_t)num_items) >> 32);
      long long lo = lo0, hi = hi0;
      bool hit = false;
      while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        const int c = __ldg(cols + mid);
        if (c < j) lo = mid + 1;
        else if (c > j) hi = mid;
        else { hit = true; break; }
      }
      if (!hit) break;
      ++attempt;
    }
    out_j[k] = j;
  }
}

}  // namespace

extern "C" int qrec_sample_neg_philox(int64_t n, int32_t num_items, const int32_t* u,
                                      const int64_t* rowptr, const int32_t* cols, uint64_t seed,
                                      uint32_t epoch, int32_t* out_j, void* stream) {
  QREC_REQUIRE(n >= 0 && num_items >= 1, "qrec_sample_neg_philox: bad sizes");
  if (n == 0) return QREC_OK;
  QREC_REQUIRE(u && rowptr && cols && out_j, "qrec_sample_neg_philox: null pointer");
  const long long blocks = (n + 255) / 256;
  const long long cap = 148LL * 16;
  sample_neg_philox_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream>>>(
      n, num_items, u, reinterpret_cast<const long long*>(rowptr), cols, (uint32_t)seed,
      (uint32_t)(seed >> 32), epoch, out_j);
  QREC_LAUNCH_CHECK();
  return QREC_OK;
}
