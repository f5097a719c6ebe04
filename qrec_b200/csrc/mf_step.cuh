// Scalar pieces of the K9 step (mf_kernels.cu), kept apart so that the CPU suite can compile and run
// the very same source (tests/host_shims/mf_step_host.cpp).  Parity flavour: numpy's evaluation order,
// every product and sum rounded separately (__fmul_rn / __fadd_rn never contract into an FMA).
//   kind 0  BasicMF.py:22-23   P[u] += (lr*e)*q ;           Q[i] += (lr*e)*P[u]
//   kind 1  PMF.py:21-22       P[u] += lr*(e*q - regU*p) ;  Q[i] += lr*(e*P[u] - regI*q)
//   kind 2  SVD.py:27-30,88    kind 1 + biases; prediction = ((dot + mean) + Bi[i]) + Bu[u]
#pragma once

namespace qrec {

__device__ __forceinline__ float mf_mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float mf_add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float mf_sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ double mf_mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double mf_add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double mf_sub(double a, double b) { return __dsub_rn(a, b); }

template <typename T, int KIND>
__device__ __forceinline__ T mf_prediction(T dot, T global_mean, T bi, T bu) {
  return KIND == 2 ? mf_add(mf_add(mf_add(dot, global_mean), bi), bu) : dot;
}

// one component of both rows; g = lr*err (kind 0 only)
template <typename T, int KIND>
__device__ __forceinline__ void mf_update_parity(T p, T q, T err, T g, T lr, T reg_u, T reg_i, T& pn, T& qn) {
  if (KIND == 0) {
    pn = mf_add(p, mf_mul(g, q));
    qn = mf_add(q, mf_mul(g, pn));
  } else {
    pn = mf_add(p, mf_mul(lr, mf_sub(mf_mul(err, q), mf_mul(reg_u, p))));
    qn = mf_add(q, mf_mul(lr, mf_sub(mf_mul(err, pn), mf_mul(reg_i, q))));
  }
}

template <typename T>
__device__ __forceinline__ T mf_bias_parity(T b, T err, T lr, T reg_b) {
  return mf_add(b, mf_mul(lr, mf_sub(err, mf_mul(reg_b, b))));
}

// throughput flavour: the row DELTAS of one component (fp32, contraction allowed)
template <int KIND>
__device__ __forceinline__ void mf_delta_fast(float p, float q, float e, float lr, float reg_u, float reg_i,
                                              float& dp, float& dq) {
  if (KIND == 0) {
    dp = (lr * e) * q;
    dq = (lr * e) * (p + dp);
  } else {
    dp = lr * (e * q - reg_u * p);
    dq = lr * (e * (p + dp) - reg_i * q);
  }
}

}  // namespace qrec
