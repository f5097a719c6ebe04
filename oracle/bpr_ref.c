/*
 * bpr_ref.c -- plain-C restatement (float64) of the reference's numpy BPR inner loop.
 * TEST INFRASTRUCTURE / CPU BASELINE ONLY: linked by tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs.  Never by the product path.
 *
 * Follows model/ranking/BPR.py:45-53 (BPR.optimization) statement by statement, with
 * util/qmath.py:127-128 for the sigmoid; the caller supplies the (u,i,j) stream that
 * BPR.py:31-38 would have produced.  Pinned by tests/test_oracle_golden.py against P,Q after
 * epoch 1 of the unmodified reference (tests/golden/bpr_filmtrust_seed0.npz).
 *
 * The reference loop is a serial dependency chain (every triple reads rows the previous one
 * may have written), so this baseline is single threaded by construction, like the original.
 */
#include <math.h>
#include <stdint.h>

/* returns sum_k -ln(s_k); P [U,d], Q [I,d] row-major float64, updated in place */
double oracle_bpr_sgd_sequential_f64(double* P, double* Q, int d, int64_t n, const int32_t* u,
                                     const int32_t* i, const int32_t* j, double lr, double reg_u,
                                     double reg_i) {
  double loss = 0.0;
  for (int64_t k = 0; k < n; ++k) {
    double* p = P + (int64_t)u[k] * d;
    double* qi = Q + (int64_t)i[k] * d;
    double* qj = Q + (int64_t)j[k] * d;
    double di = 0.0, dj = 0.0;
    for (int c = 0; c < d; ++c) { di += p[c] * qi[c]; dj += p[c] * qj[c]; }
    const double s = 1.0 / (1.0 + exp(-(di - dj)));
    const double g = lr * (1.0 - s);
    const double au = lr * reg_u, ai = lr * reg_i;
    for (int c = 0; c < d; ++c) {
      double pn = p[c] + g * (qi[c] - qj[c]);
      double qin = qi[c] + g * pn;
      double qjn = qj[c] - g * pn;
      p[c] = pn - au * pn;
      qi[c] = qin - ai * qin;
      qj[c] = qjn - ai * qjn;
    }
    loss += -log(s);
  }
  return loss;
}

/* the same loop in float32: what a sequential fp32 engine is expected to produce */
double oracle_bpr_sgd_sequential_f32(float* P, float* Q, int d, int64_t n, const int32_t* u,
                                     const int32_t* i, const int32_t* j, float lr, float reg_u,
                                     float reg_i) {
  double loss = 0.0;
  for (int64_t k = 0; k < n; ++k) {
    float* p = P + (int64_t)u[k] * d;
    float* qi = Q + (int64_t)i[k] * d;
    float* qj = Q + (int64_t)j[k] * d;
    float di = 0.f, dj = 0.f;
    for (int c = 0; c < d; ++c) { di += p[c] * qi[c]; dj += p[c] * qj[c]; }
    const float s = 1.0f / (1.0f + expf(-(di - dj)));
    const float g = lr * (1.0f - s);
    const float au = lr * reg_u, ai = lr * reg_i;
    for (int c = 0; c < d; ++c) {
      float pn = p[c] + g * (qi[c] - qj[c]);
      float qin = qi[c] + g * pn;
      float qjn = qj[c] - g * pn;
      p[c] = pn - au * pn;
      qi[c] = qin - ai * qin;
      qj[c] = qjn - ai * qjn;
    }
    loss += -log((double)s);
  }
  return loss;
}

/* Y = A X, CSR fp32 -- restates tf.sparse_tensor_dense_matmul (model/ranking/LightGCN.py:17)
 * for the LightGCN CPU baseline; accumulates in index order like the TF CPU kernel. */
void oracle_spmm_csr_f32(int32_t n_rows, const int64_t* rowptr, const int32_t* cols,
                         const float* vals, const float* X, float* Y, int d) {
  for (int32_t r = 0; r < n_rows; ++r) {
    float* y = Y + (int64_t)r * d;
    for (int c = 0; c < d; ++c) y[c] = 0.f;
    for (int64_t e = rowptr[r]; e < rowptr[r + 1]; ++e) {
      const float w = vals[e];
      const float* x = X + (int64_t)cols[e] * d;
      for (int c = 0; c < d; ++c) y[c] += w * x[c];
    }
  }
}
