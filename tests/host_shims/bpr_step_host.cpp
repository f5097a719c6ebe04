// Host build of qrec_b200/csrc/bpr_step.cuh: replays bpr_sgd_ordered_kernel's arithmetic triple by triple
// on the CPU (per-lane partial dots, the warp's xor-shuffle reduction order, then the header's own update),
// and exposes the throughput kernels' 4-wide step functions, so that the CPU suite pins the device source
// to the reference's golden run.  Built with -ffp-contract=off: the *_rn intrinsics never fuse.
#include <cmath>
#include <cstddef>
#include <cstdint>
#define __device__
#define __forceinline__ inline
struct float4 { float x, y, z, w; };
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
#include <math.h>   // expf, logf, fmaf, exp, log in the global namespace, as in CUDA
#include "bpr_step.cuh"

namespace {

template <typename T>
T warp_dot(const T* p, const T* q, int d) {
  T v[32];
  for (int lane = 0; lane < 32; ++lane) {
    T dot = 0;
    for (int c = lane; c < d; c += 32) dot += p[c] * q[c];       // d <= 32 in the tests: one product per lane
    v[lane] = dot;
  }
  for (int o = 16; o > 0; o >>= 1) {
    T w[32];
    for (int lane = 0; lane < 32; ++lane) w[lane] = v[lane] + v[lane ^ o];
    for (int lane = 0; lane < 32; ++lane) v[lane] = w[lane];
  }
  return v[0];
}

template <typename T>
double epoch(T* P, T* Q, int d, int64_t n, const int32_t* u, const int32_t* i, const int32_t* j, T lr, T reg_u, T reg_i) {
  using namespace qrec::bpr;
  double loss = 0.0;
  const T a_u = mul_rn(lr, reg_u), a_i = mul_rn(lr, reg_i);
  for (int64_t k = 0; k < n; ++k) {
    T* p = P + (size_t)u[k] * d;
    T* qi = Q + (size_t)i[k] * d;
    T* qj = Q + (size_t)j[k] * d;
    const T di = warp_dot<T>(p, qi, d), dj = warp_dot<T>(p, qj, d);
    const T s = sigmoid_full(sub_rn(di, dj));
    const T g = mul_rn(lr, sub_rn((T)1, s));
    for (int c = 0; c < d; ++c) bpr_update_parity<T>(p[c], qi[c], qj[c], g, a_u, a_i, p[c], qi[c], qj[c]);
    loss += neg_log(s);
  }
  return loss;
}

}  // namespace

extern "C" {

double host_bpr_ordered_f64(double* P, double* Q, int d, int64_t n, const int32_t* u, const int32_t* i,
                            const int32_t* j, double lr, double reg_u, double reg_i) {
  return epoch<double>(P, Q, d, n, u, i, j, lr, reg_u, reg_i);
}
double host_bpr_ordered_f32(float* P, float* Q, int d, int64_t n, const int32_t* u, const int32_t* i,
                            const int32_t* j, float lr, float reg_u, float reg_i) {
  return epoch<float>(P, Q, d, n, u, i, j, lr, reg_u, reg_i);
}

// the throughput kernels' 4-wide steps on one slice: deltas of bpr_step4, and bpr_step4_inplace's outputs
void host_bpr_step4(const float* p, const float* qi, const float* qj, float g, float a_u, float a_i, float* dp,
                    float* dqi, float* dqj) {
  float4 P4{p[0], p[1], p[2], p[3]}, I4{qi[0], qi[1], qi[2], qi[3]}, J4{qj[0], qj[1], qj[2], qj[3]}, a, b, c;
  qrec::bpr::bpr_step4(P4, I4, J4, g, a_u, a_i, a, b, c);
  dp[0] = a.x; dp[1] = a.y; dp[2] = a.z; dp[3] = a.w;
  dqi[0] = b.x; dqi[1] = b.y; dqi[2] = b.z; dqi[3] = b.w;
  dqj[0] = c.x; dqj[1] = c.y; dqj[2] = c.z; dqj[3] = c.w;
}
void host_bpr_step4_inplace(float* p, const float* qi, const float* qj, float g, float a_u, float a_i, float* dqi,
                            float* dqj) {
  float4 P4{p[0], p[1], p[2], p[3]}, I4{qi[0], qi[1], qi[2], qi[3]}, J4{qj[0], qj[1], qj[2], qj[3]}, b, c;
  qrec::bpr::bpr_step4_inplace(P4, I4, J4, g, 1.0f - a_u, g * (1.0f - a_i), a_i, b, c);
  p[0] = P4.x; p[1] = P4.y; p[2] = P4.z; p[3] = P4.w;
  dqi[0] = b.x; dqi[1] = b.y; dqi[2] = b.z; dqi[3] = b.w;
  dqj[0] = c.x; dqj[1] = c.y; dqj[2] = c.z; dqj[3] = c.w;
}

}  // extern "C"
