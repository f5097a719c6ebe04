// K1: BPR.optimization(u,i,j) on the GPU (reference: model/ranking/BPR.py:45-53).
//
//   x = P[u].Q[i] - P[u].Q[j];  s = 1/(1+exp(-x));  g = lr*(1-s)
//   P[u] += g*(Q[i]-Q[j]);  Q[i] += g*P[u](new);  Q[j] -= g*P[u](new)
//   P[u] -= lr*regU*P[u];   Q[i] -= lr*regI*Q[i];  Q[j] -= lr*regI*Q[j];   loss += -ln(s)
//
// Two kernels:
//   * bpr_sgd_ordered_kernel  -- parity mode.  The reference loop is Gauss-Seidel: triple k
//     must see every earlier update of its three rows.  Instead of level-by-level launches
//     the kernel runs the epoch as a dataflow: warps claim triples in array order from a
//     ticket counter and spin until each of their three rows has reached the version
//     (= number of earlier touches) computed by qrec_bpr_order_prepare.  Because tickets are
//     handed out in order to running warps, the oldest unfinished triple always has its
//     dependencies satisfied, so the scheme cannot deadlock whatever the grid size.
//   * bpr_sgd_batch_kernel    -- throughput mode.  LPR lanes own one triple (d=64: a half
//     warp, one float4 per lane = one 128-bit LDG per row), the dot products are reduced with
//     xor-shuffles inside the lane group and the three row deltas go back with
//     REDG.E.ADD.F32x4 (red.global.add.v4.f32).  UNROLL triples per lane group are in flight
//     before the first use so each warp keeps 2*UNROLL*3 row loads outstanding.
#include <cmath>
#include <cstdlib>

#include "common.h"
#include "philox.cuh"
#include "bpr_step.cuh"

namespace {

using namespace qrec::bpr;

// ------------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add(int* p, int v) {
  asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_add_v4(float* addr, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------------------
// parity mode
// ------------------------------------------------------------------------------------------
template <typename T, int E>  // E = ceil(d/32) elements per lane, element index e*32+lane
__global__ void __launch_bounds__(256)
bpr_sgd_ordered_kernel(T* __restrict__ P, T* __restrict__ Q, int d, long long n,
                       const int* __restrict__ u, const int* __restrict__ i,
                       const int* __restrict__ j, const int* __restrict__ wu,
                       const int* __restrict__ wi, const int* __restrict__ wj, int* ver_p,
                       int* ver_q, unsigned long long* ticket, T lr, T reg_u, T reg_i,
                       double* loss) {
  const int lane = threadIdx.x & 31;
  double local_loss = 0.0;
  const T a_u = mul_rn(lr, reg_u), a_i = mul_rn(lr, reg_i);
  while (true) {
    unsigned long long k = 0;
    if (lane == 0) k = atomicAdd(ticket, 1ULL);
    k = __shfl_sync(0xffffffffu, k, 0);
    if (k >= (unsigned long long)n) break;
    const int uu = u[k], ii = i[k], jj = j[k];
    // lanes 0..2 each watch one row version
    const int* vp = lane == 0 ? ver_p + uu : (lane == 1 ? ver_q + ii : ver_q + jj);
    const int need = lane == 0 ? wu[k] : (lane == 1 ? wi[k] : wj[k]);
    unsigned backoff = 8, polls = 0;
    while (true) {
      const int have = lane < 3 ? ld_acquire_gpu(vp) : need;
      if (__all_sync(0xffffffffu, have == need)) break;
      __nanosleep(backoff);
      if (backoff < 64) backoff <<= 1;
      // a ticket waits for at most (#resident warps) predecessors, i.e. milliseconds; ~10 s of polling
      // means the wait_* arrays do not describe this triple stream -- abort the launch instead of
      // hanging the GPU (the host sees a launch failure)
      if (++polls > (1u << 27)) __trap();
    }
    T* pr = P + (size_t)uu * d;
    T* qir = Q + (size_t)ii * d;
    T* qjr = Q + (size_t)jj * d;
    T p[E], qi[E], qj[E];
    T di = 0, dj = 0;
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int c = e * 32 + lane;
      if (c < d) {
        p[e] = __ldcg(pr + c);  // L2-coherent: the row was last written by another SM
        qi[e] = __ldcg(qir + c);
        qj[e] = __ldcg(qjr + c);
        di += p[e] * qi[e];
        dj += p[e] * qj[e];
      } else {
        p[e] = qi[e] = qj[e] = 0;
      }
    }
    di = warp_sum(di);
    dj = warp_sum(dj);
    const T s = sigmoid_full(sub_rn(di, dj));
    const T g = mul_rn(lr, sub_rn((T)1, s));
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int c = e * 32 + lane;
      if (c < d) {
        T pn, qin, qjn;
        bpr_update_parity(p[e], qi[e], qj[e], g, a_u, a_i, pn, qin, qjn);
        __stcg(pr + c, pn);
        __stcg(qir + c, qin);
        __stcg(qjr + c, qjn);
      }
    }
    __threadfence();
    __syncwarp();
    if (lane < 3) red_release_gpu_add(const_cast<int*>(vp), 1);
    if (lane == 0) local_loss += neg_log(s);
  }
  if (lane == 0 && local_loss != 0.0) atomicAdd(loss, local_loss);
}

// ------------------------------------------------------------------------------------------
// throughput mode
// ------------------------------------------------------------------------------------------
// Throughput kernels: sigmoid and -ln(s) on the SFU (ex2.approx / lg2.approx / rcp.approx).  The
// relative error (~2^-21) is far below the fp32 rounding of the row update it scales; parity mode
// keeps expf/logf.
__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float fast_neg_log(float s) { return -__logf(s); }

template <int LPR>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float dot4(float4 a, float4 b) {
  return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
}

template <int LPR, int VPL, int UNROLL>
__global__ void __launch_bounds__(256)
bpr_sgd_batch_kernel(float* __restrict__ P, float* __restrict__ Q, int nvec, long long n,
                     const int* __restrict__ u, const int* __restrict__ i,
                     const int* __restrict__ j, float lr, float reg_u, float reg_i,
                     double* loss) {
  constexpr int TPW = 32 / LPR;  // triples processed side by side in one warp
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPR, l = lane % LPR;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int d = nvec * 4;
  const float a_u = lr * reg_u, a_i = lr * reg_i;
  float lsum = 0.f;

  for (long long base = warp * 32; base < n; base += nwarps * 32) {
    const long long k = base + lane;
    int mu = 0, mi = 0, mj = 0;
    if (k < n) {
      mu = __ldg(u + k);
      mi = __ldg(i + k);
      mj = __ldg(j + k);
    }
    const int cnt = (n - base) < 32 ? (int)(n - base) : 32;
    for (int s0 = 0; s0 < cnt; s0 += TPW * UNROLL) {
      float4 p[UNROLL][VPL], qi[UNROLL][VPL], qj[UNROLL][VPL];
      float* pr[UNROLL];
      float* qir[UNROLL];
      float* qjr[UNROLL];
      bool ok[UNROLL];
#pragma unroll
      for (int r = 0; r < UNROLL; ++r) {
        const int t = s0 + r * TPW + sub;
        const int uu = __shfl_sync(0xffffffffu, mu, t & 31);
        const int ii = __shfl_sync(0xffffffffu, mi, t & 31);
        const int jj = __shfl_sync(0xffffffffu, mj, t & 31);
        ok[r] = t < cnt;
        pr[r] = P + (size_t)uu * d + l * 4;
        qir[r] = Q + (size_t)ii * d + l * 4;
        qjr[r] = Q + (size_t)jj * d + l * 4;
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
          if (ok[r] && (l + v * LPR) < nvec) {
            p[r][v] = *reinterpret_cast<const float4*>(pr[r] + v * LPR * 4);
            qi[r][v] = *reinterpret_cast<const float4*>(qir[r] + v * LPR * 4);
            qj[r][v] = *reinterpret_cast<const float4*>(qjr[r] + v * LPR * 4);
          } else {
            p[r][v] = qi[r][v] = qj[r][v] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
      }
#pragma unroll
      for (int r = 0; r < UNROLL; ++r) {
        float x = 0.f;
#pragma unroll
        for (int v = 0; v < VPL; ++v) x += dot4(p[r][v], qi[r][v]) - dot4(p[r][v], qj[r][v]);
        x = group_sum<LPR>(x);
        const float s = fast_sigmoid(x);
        const float g = lr * (1.0f - s);
        if (ok[r]) {
          if (l == 0) lsum += fast_neg_log(s);
#pragma unroll
          for (int v = 0; v < VPL; ++v) {
            if ((l + v * LPR) < nvec) {
              float4 dp, dqi, dqj;
              bpr_step4(p[r][v], qi[r][v], qj[r][v], g, a_u, a_i, dp, dqi, dqj);
              red_add_v4(pr[r] + v * LPR * 4, dp);
              red_add_v4(qir[r] + v * LPR * 4, dqi);
              red_add_v4(qjr[r] + v * LPR * 4, dqj);
            }
          }
        }
      }
    }
  }
  // block reduction of the loss: one double atomic per block
  __shared__ float wsum[8];
  lsum = warp_sum(lsum);
  if (lane == 0) wsum[threadIdx.x >> 5] = lsum;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += (double)wsum[w];
    if (t != 0.0) atomicAdd(loss, t);
  }
}


// ------------------------------------------------------------------------------------------
// K1 on staged item rows (row-sharded Q): the rows of i and j were fetched from their
// owner ranks into R[pos]; the step is the same, P is updated in place (REDG.ADD.F32x4) and the item
// deltas are written next to the fetched rows (D[pos]) to be sent back and scatter-added by the
// owner.  LPR = d/4 lanes per triple like the batch kernel; one triple per lane group per step.
// ------------------------------------------------------------------------------------------
template <int LPR>
__global__ void __launch_bounds__(256)
bpr_sgd_staged_kernel(float* __restrict__ P, int nvec, long long n, const int* __restrict__ u,
                      const int* __restrict__ pos_i, const int* __restrict__ pos_j,
                      const float* __restrict__ R, float* __restrict__ D, float lr, float reg_u,
                      float reg_i, double* loss) {
  constexpr int TPW = 32 / LPR;
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPR, l = lane % LPR;
  const long long group = (((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * TPW + sub;
  const long long ngroups = (((long long)gridDim.x * blockDim.x) >> 5) * TPW;
  const int d = nvec * 4;
  const float a_u = lr * reg_u, a_i = lr * reg_i;
  float lsum = 0.f;
  const long long rounds = (n + ngroups - 1) / ngroups;
  for (long long it = 0; it < rounds; ++it) {
    const long long k = it * ngroups + group;
    const bool ok = k < n && l < nvec;
    float4 p = make_float4(0.f, 0.f, 0.f, 0.f), qi = p, qj = p;
    float* pr = nullptr;
    size_t oi = 0, oj = 0;
    if (ok) {
      pr = P + (size_t)__ldg(u + k) * d + l * 4;
      oi = (size_t)__ldg(pos_i + k) * d + l * 4;
      oj = (size_t)__ldg(pos_j + k) * d + l * 4;
      p = *reinterpret_cast<const float4*>(pr);
      qi = __ldg(reinterpret_cast<const float4*>(R + oi));
      qj = __ldg(reinterpret_cast<const float4*>(R + oj));
    }
    float x = dot4(p, qi) - dot4(p, qj);
    x = group_sum<LPR>(x);
    const float s = fast_sigmoid(x);
    const float g = lr * (1.0f - s);
    if (ok) {
      if (l == 0) lsum += fast_neg_log(s);
      float4 dp, dqi, dqj;
      bpr_step4(p, qi, qj, g, a_u, a_i, dp, dqi, dqj);
      red_add_v4(pr, dp);
      *reinterpret_cast<float4*>(D + oi) = dqi;
      *reinterpret_cast<float4*>(D + oj) = dqj;
    }
  }
  __shared__ float wsum[8];
  lsum = warp_sum(lsum);
  if (lane == 0) wsum[threadIdx.x >> 5] = lsum;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += (double)wsum[w];
    if (t != 0.0) atomicAdd(loss, t);
  }
}


// ------------------------------------------------------------------------------------------
// throughput mode, TMA scatter variant (d = 64): identical gather/compute, but the three row deltas
// of a triple are parked in shared memory and added to the tables by the bulk-copy engine
//   cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [row], [smem], 256
// (one instruction per 256-byte row, issued by one lane, executed by the TMA unit against L2)
// instead of 16 lanes x REDG.E.ADD.F32x4 through the LSU/L1TEX path that bounds the plain kernel.
// Shared memory: per warp 2 buffers x (UNROLL*TPW triples x 3 rows x 256 B); a buffer is rewritten
// only after cp.async.bulk.wait_group.read shows that the engine has read it.
// ------------------------------------------------------------------------------------------
// TMA_MASK selects which rows go through the bulk engine (bit 0: P[u], bit 1: Q[i], bit 2: Q[j]);
// the others keep the REDG path, so the two scatter paths can run side by side.
template <int UNROLL, int TMA_MASK>
__global__ void __launch_bounds__(256)
bpr_sgd_batch_tma_kernel(float* __restrict__ P, float* __restrict__ Q, long long n,
                         const int* __restrict__ u, const int* __restrict__ i,
                         const int* __restrict__ j, float lr, float reg_u, float reg_i,
                         double* loss) {
  constexpr int LPR = 16, TPW = 2, D = 64;
  constexpr int ROWS = ((TMA_MASK >> 0) & 1) + ((TMA_MASK >> 1) & 1) + ((TMA_MASK >> 2) & 1);
  constexpr int SLOT_P = 0, SLOT_I = (TMA_MASK & 1), SLOT_J = (TMA_MASK & 1) + ((TMA_MASK >> 1) & 1);
  constexpr int SLOT_FLOATS = UNROLL * TPW * ROWS * D;         // per warp, per buffer
  extern __shared__ __align__(128) float stage[];              // [8 warps][2 buffers][SLOT_FLOATS]
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int sub = lane / LPR, l = lane % LPR;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const float a_u = lr * reg_u, a_i = lr * reg_i;
  float* my_stage = stage + (size_t)wib * 2 * SLOT_FLOATS;
  float lsum = 0.f;
  int buf = 0;

  for (long long base = warp * 32; base < n; base += nwarps * 32) {
    const long long k = base + lane;
    int mu = 0, mi = 0, mj = 0;
    if (k < n) {
      mu = __ldg(u + k);
      mi = __ldg(i + k);
      mj = __ldg(j + k);
    }
    const int cnt = (n - base) < 32 ? (int)(n - base) : 32;
    for (int s0 = 0; s0 < cnt; s0 += TPW * UNROLL) {
      float4 p[UNROLL], qi[UNROLL], qj[UNROLL];
      int uu[UNROLL], ii[UNROLL], jj[UNROLL];
      bool ok[UNROLL];
#pragma unroll
      for (int r = 0; r < UNROLL; ++r) {
        const int t = s0 + r * TPW + sub;
        uu[r] = __shfl_sync(0xffffffffu, mu, t & 31);
        ii[r] = __shfl_sync(0xffffffffu, mi, t & 31);
        jj[r] = __shfl_sync(0xffffffffu, mj, t & 31);
        ok[r] = t < cnt;
        if (ok[r]) {
          p[r] = *reinterpret_cast<const float4*>(P + (size_t)uu[r] * D + l * 4);
          qi[r] = *reinterpret_cast<const float4*>(Q + (size_t)ii[r] * D + l * 4);
          qj[r] = *reinterpret_cast<const float4*>(Q + (size_t)jj[r] * D + l * 4);
        } else {
          p[r] = qi[r] = qj[r] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      // the engine must have finished READING the buffer we are about to overwrite (issued two
      // iterations ago by lanes 0..2 of each lane group; bulk groups are tracked per thread)
      if (l < 3) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      __syncwarp();
      float* slot = my_stage + (size_t)buf * SLOT_FLOATS;
#pragma unroll
      for (int r = 0; r < UNROLL; ++r) {
        float x = dot4(p[r], qi[r]) - dot4(p[r], qj[r]);
        x = group_sum<LPR>(x);
        const float s = fast_sigmoid(x);
        const float g = lr * (1.0f - s);
        if (ok[r]) {
          if (l == 0) lsum += fast_neg_log(s);
          float4 dp, dqi, dqj;
          bpr_step4(p[r], qi[r], qj[r], g, a_u, a_i, dp, dqi, dqj);
          float* trip = slot + (size_t)(r * TPW + sub) * ROWS * D;
          if (TMA_MASK & 1) *reinterpret_cast<float4*>(trip + SLOT_P * D + l * 4) = dp;
          else red_add_v4(P + (size_t)uu[r] * D + l * 4, dp);
          if (TMA_MASK & 2) *reinterpret_cast<float4*>(trip + SLOT_I * D + l * 4) = dqi;
          else red_add_v4(Q + (size_t)ii[r] * D + l * 4, dqi);
          if (TMA_MASK & 4) *reinterpret_cast<float4*>(trip + SLOT_J * D + l * 4) = dqj;
          else red_add_v4(Q + (size_t)jj[r] * D + l * 4, dqj);
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // st.shared -> visible to the bulk engine
      __syncwarp();
      if (l < 3 && ((TMA_MASK >> l) & 1)) {
        const int srow = (l == 0) ? SLOT_P : (l == 1 ? SLOT_I : SLOT_J);
#pragma unroll
        for (int r = 0; r < UNROLL; ++r) {
          if (ok[r]) {
            float* dst = (l == 0) ? (P + (size_t)uu[r] * D) : (Q + (size_t)(l == 1 ? ii[r] : jj[r]) * D);
            const float* src = slot + (size_t)(r * TPW + sub) * ROWS * D + srow * D;
            asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(dst),
                         "r"((uint32_t)__cvta_generic_to_shared(src)), "n"(D * 4)
                         : "memory");
          }
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      buf ^= 1;
    }
  }
  if (l < 3) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // all reductions performed
  __shared__ float wsum[8];
  lsum = warp_sum(lsum);
  if (lane == 0) wsum[threadIdx.x >> 5] = lsum;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += (double)wsum[w];
    if (t != 0.0) atomicAdd(loss, t);
  }
}


// ------------------------------------------------------------------------------------------
// throughput mode, user-major (P-stationary): the reference's own iteration order
// (model/ranking/BPR.py:31-33: for user: for item).  A lane group takes one user, keeps P[u] in
// registers across that user's triples -- so P[u] is updated SEQUENTIALLY inside a user exactly as
// in the reference, and costs one row load + one row RED per user instead of per triple -- while
// the two item rows of every triple are gathered (PF triples ahead) and scatter-added with
// REDG.E.ADD.F32x4 as in the batch kernel.  Per triple: 2 row loads + 2 row REDs instead of 3 + 3.
// Input: CSR over users (rowptr), i[] / j[] in that order.
// ------------------------------------------------------------------------------------------
template <int LPR>
__device__ __forceinline__ float group_sum_masked(float v, unsigned gmask) {
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(gmask, v, o);
  return v;
}

// Work is cut into chunks of CH consecutive triples of the CSR order (balanced for any degree
// distribution, like the nnz-balanced SpMM); a lane group finds the user of its first triple with an
// LPR-ary search of rowptr and walks forward, flushing the P[u] delta (one row RED) whenever the user
// changes or the chunk ends.  Inside a user P[u] is register-resident and updated sequentially; a
// user whose triples span several chunks gets the sum of the chunks' deltas.
// SAMPLE: the negatives are drawn inside the kernel (lane l draws the negative of triple base+l with
// the same Philox counter as the stand-alone sampler, so both give identical j) instead of being
// read from j[]; they are optionally written to j_out.
struct FusedSampler {
  const long long* rated_rowptr;   // rejection sets: CSR over users, sorted columns
  const int* rated_cols;
  int num_items;
  uint32_t seed_lo, seed_hi, epoch;
  int* j_out;                      // may be null
};

// SIG: the sampler pre-tests every draw against the user's 512-bit rated signature (philox.cuh).
template <int LPR, int G, int CH, bool FULL, bool SAMPLE, int MINB = 3, bool SIG = false>   // FULL: d == 4*LPR (every lane owns a slice)
__global__ void __launch_bounds__(256, MINB)
bpr_sgd_usermajor_kernel(float* __restrict__ P, float* __restrict__ Q, int nvec, int n_users,
                         long long n, const long long* __restrict__ rowptr,
                         const int* __restrict__ i, const int* __restrict__ j, float lr,
                         float reg_u, float reg_i, double* loss, FusedSampler fs, long long trip_off,
                         const uint32_t* __restrict__ rated_sig) {
  // rowptr holds GLOBAL triple offsets; i/j are indexed relative to trip_off (a chunk of users of a
  // larger epoch: the host pipeline stages one chunk at a time).  Philox counters use global indices.
  constexpr int GPW = 32 / LPR;
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPR, l = lane % LPR;
  const unsigned gmask = (LPR == 32) ? 0xffffffffu : (((1u << LPR) - 1u) << (sub * LPR));
  const long long group = (((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * GPW + sub;
  const long long ngroups = (((long long)gridDim.x * blockDim.x) >> 5) * GPW;
  const int d = FULL ? LPR * 4 : nvec * 4;
  const bool act = FULL ? true : (l < nvec);
  const float a_u = lr * reg_u, a_i = lr * reg_i;
  const float one_m_au = 1.0f - a_u, one_m_ai = 1.0f - a_i;
  const long long nchunks = (n + CH - 1) / CH;
  float lsum = 0.f;
  for (long long ch = group; ch < nchunks; ch += ngroups) {
    const long long lo = ch * CH;
    const long long hi = (lo + CH) < n ? (lo + CH) : n;
    // user of triple lo: smallest r with rowptr[r+1] > lo
    int a = 0, b = n_users - 1;
    while (a < b) {
      const int len = b - a + 1;
      const int step = (len + LPR - 1) / LPR;
      int pp = a + (l + 1) * step - 1;
      if (pp > b) pp = b;
      const bool pred = (__ldg(rowptr + pp + 1) - trip_off) > lo;
      const unsigned bal = (__ballot_sync(gmask, pred) & gmask) >> (sub * LPR);
      const int f = __ffs(bal) - 1;
      int pf = a + (f + 1) * step - 1;
      if (pf > b) pf = b;
      a = a + f * step;
      b = pf;
    }
    int uu = a;
    long long uend = __ldg(rowptr + uu + 1) - trip_off;
    float* prow = P + (size_t)uu * d + l * 4;
    float4 p = act ? *reinterpret_cast<const float4*>(prow) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 p0 = p;
    for (long long base = lo; base < hi; base += LPR) {
      const int m = (hi - base) < LPR ? (int)(hi - base) : LPR;
      int mi = 0, mj = 0;
      if (l < m) {
        mi = __ldg(i + base + l);
        if (SAMPLE) {
          // user of triple base+l: walk forward from the group's current user
          int us = uu;
          long long ue = uend;
          while (ue <= base + l) {
            ++us;
            ue = __ldg(rowptr + us + 1) - trip_off;
          }
          if (SIG)
            mj = qrec::sample_negative_sig(base + l + trip_off, fs.epoch, fs.seed_lo, fs.seed_hi, fs.num_items,
                                           fs.rated_cols, __ldg(fs.rated_rowptr + us), __ldg(fs.rated_rowptr + us + 1),
                                           rated_sig + (size_t)us * qrec::RATED_SIG_WORDS);
          else
            mj = qrec::sample_negative(base + l + trip_off, fs.epoch, fs.seed_lo, fs.seed_hi, fs.num_items, fs.rated_cols,
                                       __ldg(fs.rated_rowptr + us), __ldg(fs.rated_rowptr + us + 1));
          if (fs.j_out != nullptr) fs.j_out[base + l] = mj;
        } else {
          mj = __ldg(j + base + l);
        }
      }
      for (int t0 = 0; t0 < m; t0 += G) {
        float4 qi[G], qj[G];
        int ri[G], rj[G];
#pragma unroll
        for (int f = 0; f < G; ++f) {                    // G triples' item rows in flight
          ri[f] = __shfl_sync(gmask, mi, sub * LPR + ((t0 + f) & (LPR - 1)));
          rj[f] = __shfl_sync(gmask, mj, sub * LPR + ((t0 + f) & (LPR - 1)));
          if (t0 + f < m && act) {
            qi[f] = *reinterpret_cast<const float4*>(Q + (size_t)ri[f] * d + l * 4);
            qj[f] = *reinterpret_cast<const float4*>(Q + (size_t)rj[f] * d + l * 4);
          } else {
            qi[f] = qj[f] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
#pragma unroll
        for (int f = 0; f < G; ++f) {
          const long long t = base + t0 + f;
          if (t0 + f < m) {                              // uniform inside the lane group
            if (t >= uend) {                             // next user: flush P delta, load the new row
              if (act) red_add_v4(prow, make_float4(p.x - p0.x, p.y - p0.y, p.z - p0.z, p.w - p0.w));
              do {
                ++uu;
                uend = __ldg(rowptr + uu + 1) - trip_off;
              } while (uend <= t);
              prow = P + (size_t)uu * d + l * 4;
              p = act ? *reinterpret_cast<const float4*>(prow) : make_float4(0.f, 0.f, 0.f, 0.f);
              p0 = p;
            }
            float x = dot4(p, qi[f]) - dot4(p, qj[f]);
            x = group_sum_masked<LPR>(x, gmask);
            const float s = fast_sigmoid(x);
            const float g = lr * (1.0f - s);
            if (l == 0) lsum += fast_neg_log(s);
            if (act) {
              float4 dqi, dqj;
              bpr_step4_inplace(p, qi[f], qj[f], g, one_m_au, g * one_m_ai, a_i, dqi, dqj);   // P[u] stays in registers
              red_add_v4(Q + (size_t)ri[f] * d + l * 4, dqi);
              red_add_v4(Q + (size_t)rj[f] * d + l * 4, dqj);
            }
          }
        }
      }
    }
    if (act) red_add_v4(prow, make_float4(p.x - p0.x, p.y - p0.y, p.z - p0.z, p.w - p0.w));
  }
  __shared__ float wsum[8];
  lsum = warp_sum(lsum);
  if (lane == 0) wsum[threadIdx.x >> 5] = lsum;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += (double)wsum[w];
    if (t != 0.0) atomicAdd(loss, t);
  }
}

// one warp per user: bit (c & 511) of the user's 16-word signature for every rated column c
__global__ void __launch_bounds__(256)
rated_signature_kernel(int n_users, const long long* __restrict__ rowptr, const int* __restrict__ cols,
                       uint32_t* __restrict__ sig) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long uu = warp; uu < n_users; uu += nwarps) {
    const long long lo = __ldg(rowptr + uu), hi = __ldg(rowptr + uu + 1);
    for (long long e = lo + lane; e < hi; e += 32) {
      const int c = __ldg(cols + e);
      atomicOr(sig + (size_t)uu * qrec::RATED_SIG_WORDS + ((c >> 5) & (qrec::RATED_SIG_WORDS - 1)), 1u << (c & 31));
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
sumsq_kernel(const T* __restrict__ x, long long n, double* out) {
  double acc = 0.0;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  if constexpr (sizeof(T) == 4) {
    const long long n4 = n >> 2;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    for (long long k = tid; k < n4; k += stride) {
      const float4 v = __ldg(x4 + k);
      acc += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
    }
    for (long long k = (n4 << 2) + tid; k < n; k += stride) acc += (double)x[k] * x[k];
  } else {
    for (long long k = tid; k < n; k += stride) acc += (double)x[k] * (double)x[k];
  }
  acc = warp_sum(acc);
  __shared__ double wsum[8];
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += wsum[w];
    atomicAdd(out, t);
  }
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    cached[dev] = v;
  }
  return cached[dev];
}

template <typename T>
int launch_ordered(T* P, T* Q, int d, long long n, const int* u, const int* i, const int* j,
                   const int* wu, const int* wi, const int* wj, int* ver_p, int* ver_q,
                   unsigned long long* ticket, T lr, T reg_u, T reg_i, double* loss,
                   int n_warps, cudaStream_t st) {
  QREC_REQUIRE(P && Q && loss && ticket && ver_p && ver_q, "bpr_sgd_ordered: null pointer");
  QREC_REQUIRE(d >= 1 && d <= 256, "bpr_sgd_ordered: d=%d unsupported (1..256)", d);
  QREC_REQUIRE(n >= 0, "bpr_sgd_ordered: n < 0");
  if (n == 0) return QREC_OK;
  QREC_REQUIRE(u && i && j && wu && wi && wj, "bpr_sgd_ordered: null index pointer");
  const int e = (d + 31) / 32;
  // 2 CTAs of 8 warps per SM: enough warps to cover the dependency DAG's width at the
  // synthetic scale (~25 independent triples per level in user-major order) without
  // drowning the LSU in pollers.
  // n_warps > 0: the caller knows the width of the dependency DAG (qrec_bpr_order_depth) and asks
  // for about that many pollers -- thousands of idle warps hammering the version counters slow the
  // few that can make progress (1.4 independent triples per level on FilmTrust, ~25 at SYN scale)
  int grid = sm_count() * 2;
  if (n_warps > 0) {
    grid = (n_warps + 7) / 8;
    if (grid < 1) grid = 1;
    if (grid > sm_count() * 2) grid = sm_count() * 2;
  }
#define QREC_ORD(E)                                                                              \
  bpr_sgd_ordered_kernel<T, E><<<grid, 256, 0, st>>>(P, Q, d, n, u, i, j, wu, wi, wj, ver_p,     \
                                                     ver_q, ticket, lr, reg_u, reg_i, loss)
  if (e <= 1) QREC_ORD(1);
  else if (e <= 2) QREC_ORD(2);
  else if (e <= 4) QREC_ORD(4);
  else QREC_ORD(8);
#undef QREC_ORD
  QREC_LAUNCH_CHECK();
  return QREC_OK;
}

}  // namespace

namespace qrec {
// shared with runtime.cu (pipelined host path)
int launch_bpr_batch(float* P, float* Q, int d, long long n, const int* u, const int* i,
                     const int* j, float lr, float reg_u, float reg_i, double* loss,
                     cudaStream_t st) {
  QREC_REQUIRE(P && Q && loss, "bpr_sgd_batch: null pointer");
  QREC_REQUIRE(d >= 4 && d <= 256 && (d % 4) == 0, "bpr_sgd_batch: d=%d unsupported (multiple of 4, 4..256)", d);
  QREC_REQUIRE(n >= 0, "bpr_sgd_batch: n < 0");
  if (n == 0) return QREC_OK;
  QREC_REQUIRE(u && i && j, "bpr_sgd_batch: null index pointer");
  const int nvec = d / 4;
  const long long warps_needed = (n + 31) / 32;
  const long long blocks_needed = (warps_needed + 7) / 8;
  const long long cap = (long long)sm_count() * 8;  // 8 CTAs x 8 warps per SM, grid-stride beyond
  const int grid = (int)(blocks_needed < cap ? blocks_needed : cap);
#define QREC_BATCH(LPR, VPL, UN)                                                                \
  bpr_sgd_batch_kernel<LPR, VPL, UN><<<grid, 256, 0, st>>>(P, Q, nvec, n, u, i, j, lr, reg_u,   \
                                                           reg_i, loss)
  if (nvec <= 4) QREC_BATCH(4, 1, 2);
  else if (nvec <= 8) QREC_BATCH(8, 1, 4);
  else if (nvec <= 16) QREC_BATCH(16, 1, 4);
  else if (nvec <= 32) QREC_BATCH(32, 1, 4);
  else QREC_BATCH(32, 2, 2);
#undef QREC_BATCH
  QREC_LAUNCH_CHECK();
  return QREC_OK;
}
}  // namespace qrec

extern "C" {

int qrec_bpr_sgd_ordered_f32(float* P, float* Q, int32_t d, int64_t n, const int32_t* u,
                             const int32_t* i, const int32_t* j, const int32_t* wu,
                             const int32_t* wi, const int32_t* wj, int32_t* ver_p,
                             int32_t* ver_q, unsigned long long* ticket, float lr, float reg_u,
                             float reg_i, double* loss, int32_t n_warps, void* stream) {
  return launch_ordered<float>(P, Q, d, n, u, i, j, wu, wi, wj, ver_p, ver_q, ticket, lr, reg_u,
                               reg_i, loss, n_warps, (cudaStream_t)stream);
}

int qrec_bpr_sgd_ordered_f64(double* P, double* Q, int32_t d, int64_t n, const int32_t* u,
                             const int32_t* i, const int32_t* j, const int32_t* wu,
                             const int32_t* wi, const int32_t* wj, int32_t* ver_p,
                             int32_t* ver_q, unsigned long long* ticket, double lr,
                             double reg_u, double reg_i, double* loss, int32_t n_warps, void* stream) {
  return launch_ordered<double>(P, Q, d, n, u, i, j, wu, wi, wj, ver_p, ver_q, ticket, lr, reg_u,
                                reg_i, loss, n_warps, (cudaStream_t)stream);
}

int qrec_bpr_sgd_batch_f32(float* P, float* Q, int32_t d, int64_t n, const int32_t* u,
                           const int32_t* i, const int32_t* j, float lr, float reg_u,
                           float reg_i, double* loss, void* stream) {
  return qrec::launch_bpr_batch(P, Q, d, n, u, i, j, lr, reg_u, reg_i, loss, (cudaStream_t)stream);
}

int qrec_bpr_sgd_batch_tma_f32(float* P, float* Q, int32_t d, int64_t n, const int32_t* u,
                               const int32_t* i, const int32_t* j, float lr, float reg_u,
                               float reg_i, double* loss, void* stream) {
  QREC_REQUIRE(P && Q && loss, "qrec_bpr_sgd_batch_tma_f32: null pointer");
  QREC_REQUIRE(d == 64, "qrec_bpr_sgd_batch_tma_f32: only d=64 (got %d); use qrec_bpr_sgd_batch_f32", d);
  QREC_REQUIRE(n >= 0, "qrec_bpr_sgd_batch_tma_f32: n < 0");
  if (n == 0) return QREC_OK;
  QREC_REQUIRE(u && i && j, "qrec_bpr_sgd_batch_tma_f32: null index pointer");
  constexpr int UN = 4;
  // which rows use the bulk engine: P only by default (the measured optimum);
  // QREC_K1_TMA_MASK=7 sends all three rows through it, 6 the two item rows
  static int mask = -1;
  if (mask < 0) {
    const char* e = getenv("QREC_K1_TMA_MASK");
    mask = e ? atoi(e) : 1;
    if (mask != 1 && mask != 6 && mask != 7) mask = 1;
  }
  const long long blocks_needed = ((n + 31) / 32 + 7) / 8;
  cudaStream_t st = (cudaStream_t)stream;
#define QREC_TMA(MASK, ROWS, CTAS)                                                                \
  {                                                                                               \
    constexpr int smem = 8 * 2 * UN * 2 * ROWS * 64 * 4;                                          \
    static bool attr_set = false;                                                                 \
    if (!attr_set) {                                                                              \
      QREC_CUDA(cudaFuncSetAttribute(bpr_sgd_batch_tma_kernel<UN, MASK>,                          \
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, smem));         \
      attr_set = true;                                                                            \
    }                                                                                             \
    const long long cap = (long long)sm_count() * CTAS;                                           \
    const int grid = (int)(blocks_needed < cap ? blocks_needed : cap);                            \
    bpr_sgd_batch_tma_kernel<UN, MASK><<<grid, 256, smem, st>>>(P, Q, n, u, i, j, lr, reg_u, reg_i, loss); \
  }
  if (mask == 7) QREC_TMA(7, 3, 2)
  else if (mask == 6) QREC_TMA(6, 2, 2)
  else QREC_TMA(1, 1, 2)
#undef QREC_TMA
  QREC_LAUNCH_CHECK();
  return QREC_OK;
}

}  // extern "C" (reopened below)

namespace qrec {
int launch_usermajor(float* P, float* Q, int32_t d, int32_t n_users, int64_t n, const int64_t* rowptr,
                     const int32_t* i, const int32_t* j, float lr, float reg_u, float reg_i, double* loss,
                     bool sample, const int64_t* rated_rowptr, const int32_t* rated_cols, int32_t num_items,
                     uint64_t seed, uint32_t epoch, int32_t* j_out, long long trip_off, cudaStream_t st,
                     const uint32_t* rated_sig) {
  FusedSampler fs = {reinterpret_cast<const long long*>(rated_rowptr), rated_cols, num_items, (uint32_t)seed,
                     (uint32_t)(seed >> 32), epoch, j_out};
  const int nvec = d / 4;
  // Grid = exactly the CTAs that are resident at once (occupancy API per instantiation), so that the launch is ONE
  // sweep over the user-major stream: at any moment the lane groups work on a contiguous window of chunks
  // (chunk = group + k * ngroups).  A larger grid makes later CTAs start again at the beginning of the stream --
  // several interleaved sweeps, i.e. a much larger re-ordering against the reference loop (measured at config 2:
  // P / Q error relative to the epoch's update 14.7 % / 37.6 % with 8 CTAs per SM vs a few % with one sweep;
  // bench.py parity_check).  QREC_K1_UM_CAP=<CTAs per SM> overrides (experiment switch).
  static int cap_mult = -1;
  if (cap_mult < 0) {
    const char* e = getenv("QREC_K1_UM_CAP");
    cap_mult = e ? atoi(e) : 0;
    if (cap_mult < 0) cap_mult = 0;
  }
  constexpr int CH = 32;
  // experiment switch (d = 64 only), measured on 50 M triples: 0 = 3 CTAs/SM, 4 triples in flight
  // (default, 5.80 ms); 1 = 4 CTAs/SM at 64 registers (spills, 7.08 ms); 2 = 2 CTAs/SM, 8 triples in
  // flight (5.75 ms) -- occupancy and load depth are not the limiter any more
  static int variant = -1;
  if (variant < 0) {
    const char* e = getenv("QREC_K1_UM_VARIANT");
    variant = e ? atoi(e) : 0;
  }
#define QREC_UM_LAUNCH(KERNEL, SIGPTR)                                                           \
  {                                                                                              \
    int occ = 3;                                                                                 \
    if (cap_mult > 0) occ = cap_mult;                                                            \
    else if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, KERNEL, 256, 0) != cudaSuccess || occ < 1) occ = 3; \
    const long long cap = (long long)sm_count() * occ;                                           \
    if (blocks > cap) blocks = cap;                                                              \
    KERNEL<<<(int)blocks, 256, 0, st>>>(P, Q, nvec, n_users, n, reinterpret_cast<const long long*>(rowptr), i, j, lr, \
                                        reg_u, reg_i, loss, fs, trip_off, SIGPTR);               \
  }
#define QREC_UM2(LPR, FULLV, SAMPLEV) QREC_UM_LAUNCH((bpr_sgd_usermajor_kernel<LPR, 4, CH, FULLV, SAMPLEV>), nullptr)
#define QREC_UM(LPR)                                                                             \
  {                                                                                              \
    const long long per_block = 8 * (32 / LPR);                                                  \
    long long blocks = ((n + CH - 1) / CH + per_block - 1) / per_block;                          \
    if (nvec == LPR && sample && rated_sig != nullptr) {                                         \
      QREC_UM_LAUNCH((bpr_sgd_usermajor_kernel<LPR, 4, CH, true, true, 3, true>), rated_sig)     \
    } else if (nvec == LPR && LPR == 16 && variant == 1) {                                       \
      if (sample) QREC_UM_LAUNCH((bpr_sgd_usermajor_kernel<16, 4, CH, true, true, 4>), nullptr)  \
      else QREC_UM_LAUNCH((bpr_sgd_usermajor_kernel<16, 4, CH, true, false, 4>), nullptr)        \
    } else if (nvec == LPR && LPR == 16 && variant == 2) {                                       \
      if (sample) QREC_UM_LAUNCH((bpr_sgd_usermajor_kernel<16, 8, CH, true, true, 2>), nullptr)  \
      else QREC_UM_LAUNCH((bpr_sgd_usermajor_kernel<16, 8, CH, true, false, 2>), nullptr)        \
    } else                                                                                       \
    if (nvec == LPR) { if (sample) QREC_UM2(LPR, true, true) else QREC_UM2(LPR, true, false) }   \
    else { if (sample) QREC_UM2(LPR, false, true) else QREC_UM2(LPR, false, false) }             \
  }
  if (nvec <= 4) QREC_UM(4)
  else if (nvec <= 8) QREC_UM(8)
  else if (nvec <= 16) QREC_UM(16)
  else QREC_UM(32)
#undef QREC_UM
#undef QREC_UM2
#undef QREC_UM_LAUNCH
  QREC_LAUNCH_CHECK();
  return QREC_OK;
}
}  // namespace qrec

extern "C" {

int qrec_bpr_sgd_usermajor_f32(float* P, float* Q, int32_t d, int32_t n_users, int64_t n, const int64_t* rowptr,
                               const int32_t* i, const int32_t* j, float lr, float reg_u, float reg_i,
                               double* loss, void* stream) {
  QREC_REQUIRE(P && Q && loss, "qrec_bpr_sgd_usermajor_f32: null pointer");
  QREC_REQUIRE(d >= 4 && d <= 128 && (d % 4) == 0, "qrec_bpr_sgd_usermajor_f32: d=%d unsupported (multiple of 4, 4..128)", d);
  QREC_REQUIRE(n_users >= 0 && n >= 0, "qrec_bpr_sgd_usermajor_f32: negative size");
  if (n_users == 0 || n == 0) return QREC_OK;
  QREC_REQUIRE(rowptr && i && j, "qrec_bpr_sgd_usermajor_f32: null index pointer");
  return qrec::launch_usermajor(P, Q, d, n_users, n, rowptr, i, j, lr, reg_u, reg_i, loss, false, nullptr, nullptr, 0,
                                0, 0, nullptr, 0, (cudaStream_t)stream, nullptr);
}

int qrec_bpr_epoch_usermajor_f32(float* P, float* Q, int32_t d, int32_t n_users, int64_t n, const int64_t* rowptr,
                                 const int32_t* i, const int64_t* rated_rowptr, const int32_t* rated_cols,
                                 int32_t num_items, uint64_t seed, uint32_t epoch, int32_t* j_out, float lr,
                                 float reg_u, float reg_i, double* loss, void* stream) {
  QREC_REQUIRE(P && Q && loss, "qrec_bpr_epoch_usermajor_f32: null pointer");
  QREC_REQUIRE(d >= 4 && d <= 128 && (d % 4) == 0, "qrec_bpr_epoch_usermajor_f32: d=%d unsupported (multiple of 4, 4..128)", d);
  QREC_REQUIRE(n_users >= 0 && n >= 0 && num_items >= 1, "qrec_bpr_epoch_usermajor_f32: bad size");
  if (n_users == 0 || n == 0) return QREC_OK;
  QREC_REQUIRE(rowptr && i && rated_rowptr && rated_cols, "qrec_bpr_epoch_usermajor_f32: null index pointer");
  return qrec::launch_usermajor(P, Q, d, n_users, n, rowptr, i, nullptr, lr, reg_u, reg_i, loss, true, rated_rowptr,
                                rated_cols, num_items, seed, epoch, j_out, 0, (cudaStream_t)stream, nullptr);
}

int qrec_rated_signature_build(int32_t n_users, const int64_t* rated_rowptr, const int32_t* rated_cols,
                               uint32_t* sig, void* stream) {
  QREC_REQUIRE(n_users >= 0, "qrec_rated_signature_build: n_users < 0");
  if (n_users == 0) return QREC_OK;
  QREC_REQUIRE(rated_rowptr && rated_cols && sig, "qrec_rated_signature_build: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  QREC_CUDA(cudaMemsetAsync(sig, 0, (size_t)n_users * qrec::RATED_SIG_WORDS * sizeof(uint32_t), st));
  long long blocks = ((long long)n_users + 7) / 8;
  const long long cap = (long long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  rated_signature_kernel<<<(int)blocks, 256, 0, st>>>(n_users, reinterpret_cast<const long long*>(rated_rowptr),
                                                     rated_cols, sig);
  QREC_LAUNCH_CHECK();
  return QREC_OK;
}

int qrec_bpr_epoch_usermajor_sig_f32(float* P, float* Q, int32_t d, int32_t n_users, int64_t n, const int64_t* rowptr,
                                     const int32_t* i, const int64_t* rated_rowptr, const int32_t* rated_cols,
                                     const uint32_t* rated_sig, int32_t num_items, uint64_t seed, uint32_t epoch,
                                     int32_t* j_out, float lr, float reg_u, float reg_i, double* loss, void* stream) {
  QREC_REQUIRE(P && Q && loss, "qrec_bpr_epoch_usermajor_sig_f32: null pointer");
  QREC_REQUIRE(d == 16 || d == 32 || d == 64 || d == 128,
               "qrec_bpr_epoch_usermajor_sig_f32: d=%d unsupported (16, 32, 64 or 128); use qrec_bpr_epoch_usermajor_f32", d);
  QREC_REQUIRE(n_users >= 0 && n >= 0 && num_items >= 1, "qrec_bpr_epoch_usermajor_sig_f32: bad size");
  if (n_users == 0 || n == 0) return QREC_OK;
  QREC_REQUIRE(rowptr && i && rated_rowptr && rated_cols && rated_sig, "qrec_bpr_epoch_usermajor_sig_f32: null index pointer");
  return qrec::launch_usermajor(P, Q, d, n_users, n, rowptr, i, nullptr, lr, reg_u, reg_i, loss, true, rated_rowptr,
                                rated_cols, num_items, seed, epoch, j_out, 0, (cudaStream_t)stream, rated_sig);
}

int qrec_bpr_sgd_staged_f32(float* P, int32_t d, int64_t n, const int32_t* u, const int32_t* pos_i,
                            const int32_t* pos_j, const float* R, float* D, float lr, float reg_u,
                            float reg_i, double* loss, void* stream) {
  QREC_REQUIRE(d >= 4 && d <= 128 && (d % 4) == 0, "qrec_bpr_sgd_staged_f32: d=%d unsupported (multiple of 4, 4..128)", d);
  QREC_REQUIRE(n >= 0, "qrec_bpr_sgd_staged_f32: n < 0");
  if (n == 0) return QREC_OK;
  QREC_REQUIRE(P && u && pos_i && pos_j && R && D && loss, "qrec_bpr_sgd_staged_f32: null pointer");
  const int nvec = d / 4;
  const long long cap = (long long)sm_count() * 8;
  cudaStream_t st = (cudaStream_t)stream;
#define QREC_STAGED(LPR)                                                                        \
  {                                                                                             \
    const long long per_block = 8 * (32 / LPR);                                                 \
    long long blocks = (n + per_block - 1) / per_block;                                         \
    if (blocks > cap) blocks = cap;                                                             \
    bpr_sgd_staged_kernel<LPR><<<(int)blocks, 256, 0, st>>>(P, nvec, n, u, pos_i, pos_j, R, D,  \
                                                            lr, reg_u, reg_i, loss);            \
  }
  if (nvec <= 4) QREC_STAGED(4)
  else if (nvec <= 8) QREC_STAGED(8)
  else if (nvec <= 16) QREC_STAGED(16)
  else QREC_STAGED(32)
#undef QREC_STAGED
  QREC_LAUNCH_CHECK();
  return QREC_OK;
}

int qrec_sumsq_f32(const float* x, int64_t n, double* out, void* stream) {
  QREC_REQUIRE(out && (x || n == 0) && n >= 0, "qrec_sumsq_f32: bad argument");
  if (n == 0) return QREC_OK;
  QREC_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, "qrec_sumsq_f32: x not 16-byte aligned");
  const long long blocks = (n / 4 + 255) / 256 + 1;
  const long long cap = (long long)sm_count() * 8;
  sumsq_kernel<float><<<(int)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream>>>(x, n, out);
  QREC_LAUNCH_CHECK();
  return QREC_OK;
}

int qrec_sumsq_f64(const double* x, int64_t n, double* out, void* stream) {
  QREC_REQUIRE(out && (x || n == 0) && n >= 0, "qrec_sumsq_f64: bad argument");
  if (n == 0) return QREC_OK;
  const long long blocks = (n + 255) / 256;
  const long long cap = (long long)sm_count() * 8;
  sumsq_kernel<double><<<(int)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream>>>(x, n, out);
  QREC_LAUNCH_CHECK();
  return QREC_OK;
}

}  // extern "C"
