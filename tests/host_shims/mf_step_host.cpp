// Host build of qrec_b200/csrc/mf_step.cuh: replays mf_sgd_ordered_kernel's arithmetic entry by entry on
// the CPU -- per-lane partial dot products, the xor-shuffle reduction tree in the warp's order, then the
// header's own step functions -- so the CPU suite pins the device source to the reference's golden runs.
// (The kernel's scheduling -- tickets and row versions -- only decides WHEN an entry runs; the values
// it computes are these.)
#include <cstddef>
#include <cstdint>
#define __device__
#define __forceinline__ inline
static inline float __fmul_rn(float a, float b) { return a * b; }      // built with -ffp-contract=off
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
#include "mf_step.cuh"

namespace {

template <typename T>
T warp_dot(const T* p, const T* q, int d) {
  T v[32];
  for (int lane = 0; lane < 32; ++lane) {
    T dot = 0;
    for (int c = lane; c < d; c += 32) dot += p[c] * q[c];
    v[lane] = dot;
  }
  for (int o = 16; o > 0; o >>= 1) {          // v += __shfl_xor_sync(v, o) on all lanes at once
    T w[32];
    for (int lane = 0; lane < 32; ++lane) w[lane] = v[lane] + v[lane ^ o];
    for (int lane = 0; lane < 32; ++lane) v[lane] = w[lane];
  }
  return v[0];
}

template <typename T, int KIND>
double epoch(T* P, T* Q, int d, int64_t n, const int32_t* u, const int32_t* i, const T* r, T lr, T reg_u, T reg_i,
             T* Bu, T* Bi, T reg_b, T gm) {
  double loss = 0.0;
  for (int64_t k = 0; k < n; ++k) {
    T* p = P + (size_t)u[k] * d;
    T* q = Q + (size_t)i[k] * d;
    const T dot = warp_dot<T>(p, q, d);
    const T bu = KIND == 2 ? Bu[u[k]] : (T)0, bi = KIND == 2 ? Bi[i[k]] : (T)0;
    const T err = qrec::mf_sub(r[k], qrec::mf_prediction<T, KIND>(dot, gm, bi, bu));
    const T g = qrec::mf_mul(lr, err);
    for (int c = 0; c < d; ++c) {
      T pn, qn;
      qrec::mf_update_parity<T, KIND>(p[c], q[c], err, g, lr, reg_u, reg_i, pn, qn);
      p[c] = pn;
      q[c] = qn;
    }
    if (KIND == 2) {
      Bu[u[k]] = qrec::mf_bias_parity<T>(bu, err, lr, reg_b);
      Bi[i[k]] = qrec::mf_bias_parity<T>(bi, err, lr, reg_b);
    }
    loss += (double)err * (double)err;
  }
  return loss;
}

template <typename T>
double dispatch(int kind, T* P, T* Q, int d, int64_t n, const int32_t* u, const int32_t* i, const T* r, T lr, T reg_u,
                T reg_i, T* Bu, T* Bi, T reg_b, T gm) {
  if (kind == 0) return epoch<T, 0>(P, Q, d, n, u, i, r, lr, reg_u, reg_i, Bu, Bi, reg_b, gm);
  if (kind == 1) return epoch<T, 1>(P, Q, d, n, u, i, r, lr, reg_u, reg_i, Bu, Bi, reg_b, gm);
  return epoch<T, 2>(P, Q, d, n, u, i, r, lr, reg_u, reg_i, Bu, Bi, reg_b, gm);
}

}  // namespace

extern "C" {

double host_mf_ordered_f64(int kind, double* P, double* Q, int d, int64_t n, const int32_t* u, const int32_t* i,
                           const double* r, double lr, double reg_u, double reg_i, double* Bu, double* Bi,
                           double reg_b, double gm) {
  return dispatch<double>(kind, P, Q, d, n, u, i, r, lr, reg_u, reg_i, Bu, Bi, reg_b, gm);
}

double host_mf_ordered_f32(int kind, float* P, float* Q, int d, int64_t n, const int32_t* u, const int32_t* i,
                           const float* r, float lr, float reg_u, float reg_i, float* Bu, float* Bi, float reg_b,
                           float gm) {
  return dispatch<float>(kind, P, Q, d, n, u, i, r, lr, reg_u, reg_i, Bu, Bi, reg_b, gm);
}

// the fast kernel's per-component deltas for one entry (rows of length d)
void host_mf_delta_fast(int kind, const float* p, const float* q, int d, float e, float lr, float reg_u, float reg_i,
                        float* dp, float* dq) {
  for (int c = 0; c < d; ++c) {
    if (kind == 0) qrec::mf_delta_fast<0>(p[c], q[c], e, lr, reg_u, reg_i, dp[c], dq[c]);
    else qrec::mf_delta_fast<1>(p[c], q[c], e, lr, reg_u, reg_i, dp[c], dq[c]);
  }
}

}  // extern "C"
