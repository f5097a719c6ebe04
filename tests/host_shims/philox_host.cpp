// Host build of qrec_b200/csrc/philox.cuh for the CPU suite: the device functions are plain scalar
// code apart from three intrinsics, which are mapped to their C++ meaning here.  Lets the tests run
// the REAL sampler source (not a restatement) against the numpy oracle without a GPU.
#include <cstddef>
#include <cstdint>
#define __device__
#define __forceinline__ inline
#define __restrict__
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32); }
template <typename T>
static inline T __ldg(const T* p) { return *p; }
#include "philox.cuh"

extern "C" {

void host_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* out) {
  qrec::philox4x32_10(c0, c1, c2, c3, k0, k1, out);
}

// negatives of triples k0 .. k0+n-1 whose users are u[0..n): plain bisection and signature pre-test
void host_sample_negatives(int64_t n, int64_t k0, const int32_t* u, const int64_t* rowptr, const int32_t* cols,
                           const uint32_t* sig, int32_t num_items, uint64_t seed, uint32_t epoch, int32_t* plain,
                           int32_t* with_sig) {
  for (int64_t t = 0; t < n; ++t) {
    const long long lo = rowptr[u[t]], hi = rowptr[u[t] + 1];
    plain[t] = qrec::sample_negative(k0 + t, epoch, (uint32_t)seed, (uint32_t)(seed >> 32), num_items, cols, lo, hi);
    with_sig[t] = qrec::sample_negative_sig(k0 + t, epoch, (uint32_t)seed, (uint32_t)(seed >> 32), num_items, cols, lo,
                                            hi, sig + (size_t)u[t] * qrec::RATED_SIG_WORDS);
  }
}

}  // extern "C"
