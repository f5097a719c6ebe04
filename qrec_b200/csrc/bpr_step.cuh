// Scalar pieces of the K1 step (bpr_kernels.cu), kept apart so that the CPU suite can compile and run the
// very same source (tests/host_shims/bpr_step_host.cpp).   reference: model/ranking/BPR.py:45-53
//   x = P[u].Q[i] - P[u].Q[j];  s = 1/(1+exp(-x));  g = lr*(1-s)
//   P[u] += g*(Q[i]-Q[j]);  Q[i] += g*P[u](new);  Q[j] -= g*P[u](new)
//   P[u] -= lr*regU*P[u];   Q[i] -= lr*regI*Q[i];  Q[j] -= lr*regI*Q[j]
#pragma once

namespace qrec {
namespace bpr {

// Round-to-nearest mul/add/sub that ptxas never contracts into an FMA: numpy evaluates
// `P[u] += g*(Q[i]-Q[j])` as separate multiply and add, and parity mode follows it.
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }

__device__ __forceinline__ float sigmoid_full(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ double sigmoid_full(double x) { return 1.0 / (1.0 + exp(-x)); }
__device__ __forceinline__ double neg_log(float s) { return -(double)logf(s); }
__device__ __forceinline__ double neg_log(double s) { return -log(s); }

// parity flavour: one component of the three rows, the reference's statements in order
template <typename T>
__device__ __forceinline__ void bpr_update_parity(T p, T qi, T qj, T g, T a_u, T a_i, T& pn_out, T& qin_out, T& qjn_out) {
  T pn = add_rn(p, mul_rn(g, sub_rn(qi, qj)));
  T qin = add_rn(qi, mul_rn(g, pn));
  T qjn = sub_rn(qj, mul_rn(g, pn));
  pn_out = sub_rn(pn, mul_rn(a_u, pn));
  qin_out = sub_rn(qin, mul_rn(a_i, qin));
  qjn_out = sub_rn(qjn, mul_rn(a_i, qjn));
}

// One BPR step on a 4-wide slice; returns the three deltas.
__device__ __forceinline__ void bpr_step4(float4 p, float4 qi, float4 qj, float g, float a_u,
                                          float a_i, float4& dp, float4& dqi, float4& dqj) {
#define QREC_STEP(c)                                  \
  {                                                   \
    float pn = p.c + g * (qi.c - qj.c);               \
    float qin = qi.c + g * pn;                        \
    float qjn = qj.c - g * pn;                        \
    dp.c = (pn - a_u * pn) - p.c;                     \
    dqi.c = (qin - a_i * qin) - qi.c;                 \
    dqj.c = (qjn - a_i * qjn) - qj.c;                 \
  }
  QREC_STEP(x) QREC_STEP(y) QREC_STEP(z) QREC_STEP(w)
#undef QREC_STEP
}

// The same step with the decay folded into the coefficients (7 instead of 11 flops per component):
//   pn = p + g (qi - qj);  p' = (1 - a_u) pn;  dqi = g (1 - a_i) pn - a_i qi;  dqj = -g (1 - a_i) pn - a_i qj
// (identical algebra to BPR.py:46-52; differs from bpr_step4 by one rounding of the (1-a) product).
__device__ __forceinline__ void bpr_step4_inplace(float4& p, float4 qi, float4 qj, float g, float one_m_au,
                                                  float c1, float a_i, float4& dqi, float4& dqj) {
#define QREC_STEP(c)                                   \
  {                                                    \
    const float pn = fmaf(g, qi.c - qj.c, p.c);        \
    dqi.c = fmaf(c1, pn, -a_i * qi.c);                 \
    dqj.c = fmaf(-c1, pn, -a_i * qj.c);                \
    p.c = one_m_au * pn;                               \
  }
  QREC_STEP(x) QREC_STEP(y) QREC_STEP(z) QREC_STEP(w)
#undef QREC_STEP
}

}  // namespace bpr
}  // namespace qrec
