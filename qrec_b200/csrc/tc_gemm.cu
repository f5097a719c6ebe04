// K5 building block: the tensor-core GEMM of NeuMF's MLP (model/ranking/NeuMF.py:39-50),
//     C[M,N] = epilogue( A[M,K] * B )          fp32 in HBM, TF32 tcgen05.mma, fp32 accumulate in TMEM
// written directly against the sm_100a tensor-core path:
//   * operands are staged in shared memory in the canonical K-major SWIZZLE_128B layout
//     (rows of 32 fp32 = 128 B, 8-row 1024 B atoms, 16-byte chunk index XOR row%8) -- the MLP's
//     weight matrices are [K,N] row-major, so they are transposed on the way into shared memory;
//   * one elected thread issues tcgen05.mma.cta_group::1.kind::tf32 (M=128, N=64, K=8 per
//     instruction, 4 per 128-byte stage) with the 64-bit shared-memory matrix descriptors and the
//     32-bit instruction descriptor built below; accumulators live in 64 TMEM columns;
//   * stages are recycled through mbarriers signalled by tcgen05.commit (2-stage ring), the
//     epilogue reads the accumulator with tcgen05.ld.32x32b.x16 (warp w owns TMEM lanes 32w..32w+31),
//     applies bias / ReLU / ReLU-mask and writes fp32 rows.
//   * global loads are register-staged one k-block ahead (their latency overlaps the MMAs), the
//     epilogue is parked in shared memory and written as whole 256-byte rows.
// Tile: 128 x 64 per CTA, 128 threads.
#include "common.h"

namespace {

constexpr int BM = 128, BN = 64, BK = 32;           // BK fp32 = 128 B = one swizzle span
constexpr int STAGE_A = BM * 128, STAGE_B = BN * 128;
constexpr int SMEM_BYTES = 2 * (STAGE_A + STAGE_B) + 1024;   // + alignment slack

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

// K-major SWIZZLE_128B descriptor: start>>4 | LBO=0 | SBO=1024>>4 | version 1 | layout SWIZZLE_128B
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);           // bits 0-13  start address
  d |= (uint64_t)0 << 16;                            // bits 16-29 leading byte offset (unused, K-major swizzled)
  d |= (uint64_t)(1024 >> 4) << 32;                  // bits 32-45 stride byte offset: 8 rows * 128 B
  d |= (uint64_t)1 << 46;                            // bits 46-47 descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                            // bits 61-63 layout type: SWIZZLE_128B
  return d;
}

// instruction descriptor, kind::tf32: D=F32, A=B=TF32, both K-major, N=64, M=128
__device__ __forceinline__ uint32_t make_idesc() {
  uint32_t i = 0;
  i |= 1u << 4;                 // D format F32
  i |= 2u << 7;                 // A format TF32
  i |= 2u << 10;                // B format TF32
  i |= (uint32_t)(BN >> 3) << 17;
  i |= (uint32_t)(BM >> 4) << 24;
  return i;
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}

// byte offset of element (row, k) inside a K-major SWIZZLE_128B tile (k in [0,32) fp32)
__device__ __forceinline__ uint32_t sw_off(int row, int k) {
  const int chunk = (k >> 2) ^ (row & 7);
  return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + chunk * 16 + (k & 3) * 4);
}

// tcgen05 kind::tf32 reads the top 19 bits of each fp32 operand, i.e. truncates.  Rounding to
// nearest-away while staging removes the systematic toward-zero bias (2^-11 unbiased instead of up
// to 2^-10 one-sided per operand), which matters over the 6-GEMM forward/backward chain of the MLP.
__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ float4 to_tf32(float4 v) {
  return make_float4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
}

enum Epilogue { EPI_NONE = 0, EPI_BIAS_RELU = 1, EPI_RELU_MASK = 2, EPI_BIAS = 3 };

// B_IS_NK: B is stored [N,K] row-major (already K-major); otherwise [K,N] row-major.
template <bool B_IS_NK>
__global__ void __launch_bounds__(128)
tc_gemm_tf32_kernel(int M, int N, int K, const float* __restrict__ A, int lda,
                    const float* __restrict__ B, int ldb, float* __restrict__ C, int ldc,
                    int epi, const float* __restrict__ bias, const float* __restrict__ mask,
                    int ldmask) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t mma_done[2];
  __shared__ uint32_t tmem_base_slot;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  // stage s: A tile at smem + s*(STAGE_A+STAGE_B), B tile right behind it (computed, not looked up:
  // a pointer array indexed by the stage lands in local memory)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;

  if (tid == 0) {
    mbar_init(&mma_done[0], 1);
    mbar_init(&mma_done[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "n"(BN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_acc = tmem_base_slot;
  const uint32_t idesc = make_idesc();

  const int nkb = (K + BK - 1) / BK;
  // Register-staged global loads, one k-block ahead: the loads of block kb+1 are issued before the
  // shared-memory stores / MMAs of block kb, so their latency overlaps the tensor-core work.
  float4 ra[8];
  float4 rb4[4];
  float rb1[16];
  auto load_block = [&](int kb) {
    const int k0 = kb * BK;
#pragma unroll
    for (int p = 0; p < 8; ++p) {
      const int row = (tid >> 3) + 16 * p, c = tid & 7;
      const int gm = m0 + row, gk = k0 + c * 4;
      ra[p] = (gm < M && gk < K) ? __ldg(reinterpret_cast<const float4*>(A + (size_t)gm * lda + gk))
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (B_IS_NK) {
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const int row = (tid >> 3) + 16 * p, c = tid & 7;
        const int gn = n0 + row, gk = k0 + c * 4;
        rb4[p] = (gn < N && gk < K) ? __ldg(reinterpret_cast<const float4*>(B + (size_t)gn * ldb + gk))
                                    : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    } else {
#pragma unroll
      for (int p = 0; p < 16; ++p) {
        const int n = tid & 63, k = (tid >> 6) + 2 * p;
        const int gn = n0 + n, gk = k0 + k;
        rb1[p] = (gn < N && gk < K) ? __ldg(B + (size_t)gk * ldb + gn) : 0.f;     // coalesced along n
      }
    }
  };
  load_block(0);
  for (int kb = 0; kb < nkb; ++kb) {
    const int s = kb & 1;
    uint8_t* const sA_s = smem + s * (STAGE_A + STAGE_B);
    uint8_t* const sB_s = sA_s + STAGE_A;
    if (kb >= 2) mbar_wait(&mma_done[s], (uint32_t)(((kb >> 1) - 1) & 1));   // MMAs that read stage s are done
#pragma unroll
    for (int p = 0; p < 8; ++p) {
      const int row = (tid >> 3) + 16 * p, c = tid & 7;
      *reinterpret_cast<float4*>(sA_s + sw_off(row, c * 4)) = to_tf32(ra[p]);
    }
    if (B_IS_NK) {
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const int row = (tid >> 3) + 16 * p, c = tid & 7;
        *reinterpret_cast<float4*>(sB_s + sw_off(row, c * 4)) = to_tf32(rb4[p]);
      }
    } else {
#pragma unroll
      for (int p = 0; p < 16; ++p) {
        const int n = tid & 63, k = (tid >> 6) + 2 * p;
        *reinterpret_cast<float*>(sB_s + sw_off(n, k)) = to_tf32(rb1[p]);        // transposed into K-major
      }
    }
    if (kb + 1 < nkb) load_block(kb + 1);                                        // in flight during the MMAs
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> async proxy (UMMA)
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint64_t da = make_desc(smem_u32(sA_s)), db = make_desc(smem_u32(sB_s));
#pragma unroll
      for (int k4 = 0; k4 < BK / 8; ++k4) {
        const uint32_t acc = (kb > 0 || k4 > 0) ? 1u : 0u;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_acc),
            "l"(da + (uint64_t)(k4 * 2)), "l"(db + (uint64_t)(k4 * 2)), "r"(idesc), "r"(acc)
            : "memory");                                              // +2 = 32 bytes (8 tf32) along K
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mma_done[s]))
                   : "memory");
    }
  }
  // the last commit covers every earlier MMA (they retire in issue order)
  {
    const int last = nkb - 1;
    mbar_wait(&mma_done[last & 1], (uint32_t)((last >> 1) & 1));
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // ---- epilogue: TMEM -> registers -> shared memory (the operand stages are free now) -> coalesced
  // global rows.  Warp w reads TMEM lanes [32w, 32w+32) (one accumulator row per thread) and parks
  // them in a [128][64+4] fp32 tile (row pitch 272 B: conflict-free 16-byte stores); then every warp
  // writes whole 256-byte rows.
  float* tile = reinterpret_cast<float*>(smem);
  constexpr int PITCH = BN + 4;
  {
    const int r_in_tile = warp * 32 + lane;
#pragma unroll
    for (int c0 = 0; c0 < BN; c0 += 16) {
      uint32_t r[16];
      const uint32_t taddr = tmem_acc + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
            "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int q = 0; q < 16; q += 4)
        *reinterpret_cast<uint4*>(tile + r_in_tile * PITCH + c0 + q) = make_uint4(r[q], r[q + 1], r[q + 2], r[q + 3]);
    }
  }
  __syncthreads();
  {
    const int c4 = (tid & 15) * 4;                 // 16 threads cover one 64-float row
    const int col = n0 + c4;
    float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
    const bool vec_ok = (col + 3 < N) && ((ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
    if ((epi == EPI_BIAS_RELU || epi == EPI_BIAS) && col < N) {
      bv.x = bias[col];
      if (col + 1 < N) bv.y = bias[col + 1];
      if (col + 2 < N) bv.z = bias[col + 2];
      if (col + 3 < N) bv.w = bias[col + 3];
    }
    for (int rr = tid >> 4; rr < BM; rr += 8) {
      const int row = m0 + rr;
      if (row >= M || col >= N) continue;
      float4 v = *reinterpret_cast<const float4*>(tile + rr * PITCH + c4);
      if (epi == EPI_BIAS_RELU) {
        v.x = fmaxf(v.x + bv.x, 0.f); v.y = fmaxf(v.y + bv.y, 0.f); v.z = fmaxf(v.z + bv.z, 0.f); v.w = fmaxf(v.w + bv.w, 0.f);
      } else if (epi == EPI_BIAS) {
        v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
      } else if (epi == EPI_RELU_MASK) {
        const float* mk = mask + (size_t)row * ldmask + col;
        v.x = mk[0] > 0.f ? v.x : 0.f;
        if (col + 1 < N) v.y = mk[1] > 0.f ? v.y : 0.f;
        if (col + 2 < N) v.z = mk[2] > 0.f ? v.z : 0.f;
        if (col + 3 < N) v.w = mk[3] > 0.f ? v.w : 0.f;
      }
      float* dst = C + (size_t)row * ldc + col;
      if (vec_ok) {
        *reinterpret_cast<float4*>(dst) = v;
      } else {
        dst[0] = v.x;
        if (col + 1 < N) dst[1] = v.y;
        if (col + 2 < N) dst[2] = v.z;
        if (col + 3 < N) dst[3] = v.w;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "n"(BN));
  }
}

}  // namespace

extern "C" int qrec_tc_gemm_tf32(int32_t b_is_nk, int32_t M, int32_t N, int32_t K, const float* A,
                                 int32_t lda, const float* B, int32_t ldb, float* C, int32_t ldc,
                                 int32_t epilogue, const float* bias, const float* mask,
                                 int32_t ldmask, void* stream) {
  QREC_REQUIRE(M >= 0 && N >= 0 && K >= 1, "qrec_tc_gemm_tf32: bad dimensions");
  if (M == 0 || N == 0) return QREC_OK;
  QREC_REQUIRE(A && B && C, "qrec_tc_gemm_tf32: null pointer");
  QREC_REQUIRE(K % 4 == 0 && lda % 4 == 0 && (reinterpret_cast<uintptr_t>(A) & 15) == 0,
               "qrec_tc_gemm_tf32: A must be 16-byte aligned with K and lda multiples of 4");
  QREC_REQUIRE(!b_is_nk || (ldb % 4 == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0),
               "qrec_tc_gemm_tf32: [N,K] B must be 16-byte aligned with ldb a multiple of 4");
  QREC_REQUIRE((M + BM - 1) / BM <= 65535, "qrec_tc_gemm_tf32: M=%d exceeds the 65535 x 128 rows of one launch; split the batch", M);
  QREC_REQUIRE(epilogue >= 0 && epilogue <= 3, "qrec_tc_gemm_tf32: unknown epilogue %d", epilogue);
  QREC_REQUIRE((epilogue != EPI_BIAS_RELU && epilogue != EPI_BIAS) || bias, "qrec_tc_gemm_tf32: bias epilogue without bias");
  QREC_REQUIRE(epilogue != EPI_RELU_MASK || mask, "qrec_tc_gemm_tf32: mask epilogue without mask");
  static bool attr_set[2] = {false, false};
  if (!attr_set[b_is_nk ? 1 : 0]) {
    if (b_is_nk) QREC_CUDA(cudaFuncSetAttribute(tc_gemm_tf32_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    else QREC_CUDA(cudaFuncSetAttribute(tc_gemm_tf32_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_set[b_is_nk ? 1 : 0] = true;
  }
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM);
  cudaStream_t st = (cudaStream_t)stream;
  if (b_is_nk) tc_gemm_tf32_kernel<true><<<grid, 128, SMEM_BYTES, st>>>(M, N, K, A, lda, B, ldb, C, ldc, epilogue, bias, mask, ldmask);
  else tc_gemm_tf32_kernel<false><<<grid, 128, SMEM_BYTES, st>>>(M, N, K, A, lda, B, ldb, C, ldc, epilogue, bias, mask, ldmask);
  QREC_LAUNCH_CHECK();
  return QREC_OK;
}
