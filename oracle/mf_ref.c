/*
 * mf_ref.c -- plain-C restatement of the reference's pointwise MF loops.
 * TEST INFRASTRUCTURE / CPU BASELINE ONLY, like bpr_ref.c: never linked by the product path.
 *
 * Follows model/rating/BasicMF.py:13-23 (kind 0), model/rating/PMF.py:13-22 (kind 1) and
 * model/rating/SVD.py:17-32 with predictForRating SVD.py:84-90 (kind 2), one entry (u,i,r) at a
 * time in the caller's order.  `p = self.P[u]` is a view in the reference: the item row is
 * updated from the ALREADY UPDATED user row, the user row from the old item row.
 * Pinned by tests/test_oracle_mf_golden.py against the golden runs of the unmodified reference
 * (tables after three epochs; the dot product here is a plain left-to-right loop, numpy's ddot
 * may group differently, hence a 1e-11 tolerance instead of bit equality for this C version --
 * the numpy restatement in mf_oracle.py is the bit-exact one).
 */
#include <stdint.h>

#define MF_BODY(T)                                                                               \
  double loss = 0.0;                                                                             \
  for (int64_t k = 0; k < n; ++k) {                                                              \
    T* p = P + (int64_t)u[k] * d;                                                                \
    T* q = Q + (int64_t)i[k] * d;                                                                \
    T dot = 0;                                                                                   \
    for (int c = 0; c < d; ++c) dot += p[c] * q[c];                                              \
    T pred = dot;                                                                                \
    if (kind == 2) pred = ((dot + global_mean) + Bi[i[k]]) + Bu[u[k]];                           \
    const T e = r[k] - pred;                                                                     \
    loss += (double)e * (double)e;                                                               \
    if (kind == 0) {                                                                             \
      const T g = lr * e;                                                                        \
      for (int c = 0; c < d; ++c) {                                                              \
        const T pn = p[c] + g * q[c];                                                            \
        q[c] = q[c] + g * pn;                                                                    \
        p[c] = pn;                                                                               \
      }                                                                                          \
    } else {                                                                                     \
      for (int c = 0; c < d; ++c) {                                                              \
        const T pn = p[c] + lr * (e * q[c] - reg_u * p[c]);                                      \
        q[c] = q[c] + lr * (e * pn - reg_i * q[c]);                                              \
        p[c] = pn;                                                                               \
      }                                                                                          \
      if (kind == 2) {                                                                           \
        Bu[u[k]] += lr * (e - reg_b * Bu[u[k]]);                                                 \
        Bi[i[k]] += lr * (e - reg_b * Bi[i[k]]);                                                 \
      }                                                                                          \
    }                                                                                            \
  }                                                                                              \
  return loss;

/* returns sum_k error_k^2; tables row-major, updated in place; Bu/Bi may be NULL unless kind == 2 */
double oracle_mf_sgd_sequential_f64(int kind, double* P, double* Q, int d, int64_t n, const int32_t* u,
                                    const int32_t* i, const double* r, double lr, double reg_u,
                                    double reg_i, double* Bu, double* Bi, double reg_b, double global_mean) {
  MF_BODY(double)
}

/* the same loop in float32 (float ratings): what a sequential fp32 engine is expected to produce */
double oracle_mf_sgd_sequential_f32(int kind, float* P, float* Q, int d, int64_t n, const int32_t* u,
                                    const int32_t* i, const float* r, float lr, float reg_u, float reg_i,
                                    float* Bu, float* Bi, float reg_b, float global_mean) {
  MF_BODY(float)
}
