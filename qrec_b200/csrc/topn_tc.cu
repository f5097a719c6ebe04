// K8 on the tensor cores ("tiled P_tile . Q^T on tensor cores -> rated positions := 0 -> per-row
// top-N").  Same contract and same selection rule as csrc/topn_kernels.cu (the reference flow of
// base/recommender.py:143-152 + util/qmath.py:134-146); what changes is where the scores come from:
//
//   * a CTA owns 128 users.  Their rows of U are split once into two TF32 operands, hi = rna_tf32(x) and
//     lo = rna_tf32(x - hi), stored K-major / SWIZZLE_128B in shared memory (the layout of csrc/tc_gemm.cu);
//   * the item table is split the same way ONCE per call by a small pre-pass (split_items_kernel) that writes each
//     128-item tile as one contiguous block already in the shared-memory layout; the main kernel streams those blocks
//     through a 2-stage ring with cp.async.bulk (one 64 KB bulk copy per tile, completion on an mbarrier) -- no
//     thread touches the item operands;
//   * one thread issues tcgen05.mma.cta_group::1.kind::tf32 (M=128, N=128, K=8) three times per k-step --
//     hi.hi + lo.hi + hi.lo, the classical 3xTF32 error-compensated product: the dropped lo.lo term is 2^-22
//     relative, i.e. fp32-level scores, which is what keeps the index lists equal to an fp32 GEMV's wherever
//     scores are distinct (plain TF32's 10-bit mantissa reorders close scores) -- into one of two 128-column TMEM
//     accumulators;
//   * while the tensor cores work on tile t, 256 threads read tile t-1's accumulator with tcgen05.ld -- two threads
//     per user row (warps w and w+4 own TMEM lanes 32 (w%4)..; one takes columns 0-63 of the tile, the other 64-127),
//     each with its OWN candidate list and cut-off (the N best overall are among the N best of the two halves; the
//     lists are merged at the end) -- and run the selection of topn_kernels.cu with the count and cut-off in
//     REGISTERS: 16 scores are compared without a branch; the few that beat the cut-off look up the row's 512-bit
//     rated-set signature (built in shared memory when the kernel starts), run the exact rated test (bisection) only
//     on a signature hit, and are appended to the row's 512-key list (L2-resident scratch).  Whenever a list could
//     overflow, its warp finds the list's N-th largest key by a bit-wise search (16 keys per lane in registers, one
//     warp reduction per bit -- no sort) and keeps the N keys at or above it; rows are sorted once, at the end.
// Nothing of the [users x items] matrix is written.  d <= 64, a multiple of 4 (one or two 128-byte k-blocks, zero-padded).
#include "common.h"

namespace {

constexpr int CAP = 320;   // candidate slots per half-row list (>= N_max + TRIG_EXTRA + the 64 items a tile can add)
constexpr int TRIG_EXTRA = 96;   // a list is cut back to its N best once it holds more than N + TRIG_EXTRA keys
constexpr int SORTN = 256; // keys of the final per-row sort (two lists of at most N_max keys)
constexpr int NT = 256;    // selecting threads per CTA: 8 warps, two per TMEM lane quarter (a 9th warp drives TMA and the MMAs)
constexpr int NBUF = 4;    // TMEM accumulators of 128 columns: the tensor cores may run up to 3 tiles ahead of the slowest warp
constexpr int SIGW = 16;   // 32-bit words of a row's rated-set signature (512 bits) kept in shared memory
constexpr int NMAX = 100;  // base/recommender.py:131-134 clamps N to <= 100
constexpr int TM = 128, TN = 128;
constexpr int KBLK = TM * 128;                       // bytes of one k-block (32 fp32 = 128 B per row) of a 128-row operand

__device__ __forceinline__ uint32_t ord_of(float s) {          // monotone float -> uint
  const uint32_t u = __float_as_uint(s);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float score_of(uint32_t o) {
  return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}
__device__ __forceinline__ bool is_rated(const int* __restrict__ cols, long long lo, long long hi, int item) {
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    const int c = __ldg(cols + mid);
    if (c == item) return true;
    if (c < item) lo = mid + 1; else hi = mid;
  }
  return false;
}
// one warp sorts SZ keys, descending (bitonic network in shared memory)
template <int SZ>
__device__ __forceinline__ void warp_sort_desc(unsigned long long* k, int lane) {
#pragma unroll 1
  for (int size = 2; size <= SZ; size <<= 1) {
#pragma unroll 1
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncwarp();
      for (int t = lane; t < SZ / 2; t += 32) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool desc = (lo & size) == 0;
        const unsigned long long a = k[lo], b = k[hi];
        if ((a < b) == desc) { k[lo] = b; k[hi] = a; }
      }
    }
  }
  __syncwarp();
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major SWIZZLE_128B shared-memory matrix descriptor (csrc/tc_gemm.cu): start>>4 | SBO = 1024 B | version 1 | swizzle 128B
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// instruction descriptor, kind::tf32: D = F32, A = B = TF32, both K-major, N = 128, M = 128
__device__ __forceinline__ uint32_t make_idesc() {
  uint32_t i = 0;
  i |= 1u << 4;
  i |= 2u << 7;
  i |= 2u << 10;
  i |= (uint32_t)(TN >> 3) << 17;
  i |= (uint32_t)(TM >> 4) << 24;
  return i;
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
// Bounded wait: a protocol error traps (the launch fails with an error) instead of hanging the device.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  for (uint32_t spins = 0;; ++spins) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (ok) return;
    if (spins > (1u << 22)) __trap();
  }
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// bulk copy global -> shared (TMA engine), completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// byte offset of element (row, k) inside one K-major SWIZZLE_128B k-block (k in [0,32) fp32)
__device__ __forceinline__ uint32_t sw_off(int row, int k) {
  const int chunk = (k >> 2) ^ (row & 7);
  return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + chunk * 16 + (k & 3) * 4);
}
__device__ __forceinline__ float rna_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
// x = hi + lo with hi, lo representable in TF32 (up to 2^-22 |x|)
__device__ __forceinline__ void split_tf32(const float4 v, float4& hi, float4& lo) {
  hi = make_float4(rna_tf32(v.x), rna_tf32(v.y), rna_tf32(v.z), rna_tf32(v.w));
  lo = make_float4(rna_tf32(v.x - hi.x), rna_tf32(v.y - hi.y), rna_tf32(v.z - hi.z), rna_tf32(v.w - hi.w));
}

// Pre-pass: tile t of the item table (items [128 t, 128 t + 128), zero rows past the end, zero columns past d) as one
// block of 2 * KB * 16 KB in the workspace: hi operand, then lo operand, each K-major SWIZZLE_128B -- byte for byte what
// the main kernel wants in shared memory.
template <int KB>
__global__ void __launch_bounds__(128)
split_items_kernel(const float* __restrict__ V, int d, int n_items, uint8_t* __restrict__ blocks) {
  constexpr int D = KB * 32;
  constexpr int OPER = KB * KBLK;
  constexpr int VPT = TN * D / 4 / 128;
  const int c0 = blockIdx.x * TN;
  uint8_t* const out = blocks + (size_t)blockIdx.x * 2 * OPER;
#pragma unroll
  for (int p = 0; p < VPT; ++p) {
    const int q = threadIdx.x + 128 * p;
    const int row = q / (D / 4), c4 = (q % (D / 4)) * 4;
    const float4 v = (c0 + row < n_items && c4 < d) ? __ldg(reinterpret_cast<const float4*>(V + (size_t)(c0 + row) * d + c4))
                                                    : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 hi, lo;
    split_tf32(v, hi, lo);
    const uint32_t off = (uint32_t)(c4 >> 5) * KBLK + sw_off(row, c4 & 31);
    *reinterpret_cast<float4*>(out + off) = hi;
    *reinterpret_cast<float4*>(out + OPER + off) = lo;
  }
}

// KB = d / 32 k-blocks.  Shared memory: A hi | A lo (KB x 16 KB each), then two B stages (hi | lo, KB x 16 KB each),
// then one 2 KB sort buffer per warp (8 warps) and the 128 rows' 512-bit rated-set signatures (8 KB).
template <int KB>
__global__ void __launch_bounds__(NT + 32, 1)
score_topn_tc_kernel(const float* __restrict__ U, const uint8_t* __restrict__ item_blocks, int d, int n_items,
                     const int* __restrict__ user_ids, int n_rows, const long long* __restrict__ rated_rowptr,
                     const int* __restrict__ rated_cols, float rated_value, int N, int* __restrict__ out_ids,
                     float* __restrict__ out_scores, unsigned long long* __restrict__ workspace) {
  constexpr int D = KB * 32;
  constexpr int OPER = KB * KBLK;                     // bytes of one 128-row operand (hi or lo)
  constexpr int APT = TM * D / 4 / NT;                // float4 per thread of the users' 128-row operand (4 or 8)
  extern __shared__ uint8_t smem_raw[];
  __shared__ int cnt_sh[2][TM];
  __shared__ uint64_t mma_done[NBUF];                 // accumulator b holds a finished tile (tcgen05.commit)
  __shared__ uint64_t acc_free[NBUF];                 // all 8 selecting warps are done with accumulator b
  __shared__ uint64_t full[2];                        // operand stage s holds a whole tile (bulk-copy bytes counted)
  __shared__ uint32_t tmem_base_slot;
  // 1024-byte alignment by an offset from the array itself (not an integer round trip): the compiler keeps the
  // shared address space, so the sort / rated buffers are read with LDS / written with STS
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* const sA_hi = smem;
  uint8_t* const sA_lo = smem + OPER;
  uint8_t* const sB = smem + 2 * OPER;                // stage s: hi at sB + s * 2 * OPER, lo right behind it
  unsigned long long* const sort_buf = reinterpret_cast<unsigned long long*>(smem + 6 * OPER);
  uint32_t* const sig = reinterpret_cast<uint32_t*>(smem + 6 * OPER + 8 * SORTN * sizeof(unsigned long long));   // [SIGW][TM], word-major
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wq = warp & 3, half = warp >> 2;          // TMEM lane quarter; which 64 columns of a tile this thread selects from
  const int rowl = wq * 32 + lane;                    // the thread's user row inside the CTA (= its TMEM lane)
  const int row0 = blockIdx.x * TM;
  const int my_row = row0 + rowl;
  const int u = (my_row < n_rows) ? __ldg(user_ids + my_row) : -1;
  unsigned long long* const cand = workspace + (size_t)blockIdx.x * TM * 2 * CAP;   // this CTA's 2 x 128 lists
  unsigned long long* const my_cand = cand + ((size_t)rowl * 2 + half) * CAP;
  unsigned long long* const my_sort = sort_buf + (size_t)warp * SORTN;
  long long rlo = 0, rhi = 0;
  if (u >= 0) { rlo = __ldg(rated_rowptr + u); rhi = __ldg(rated_rowptr + u + 1); }
  const unsigned long long rated_key_hi = (unsigned long long)ord_of(rated_value) << 32;
  // this row's rated-set signature: bit hash(item) of 512 (the owning thread is the only writer of its column)
  if (half == 0) {
#pragma unroll
    for (int w = 0; w < SIGW; ++w) sig[w * TM + rowl] = 0u;
    for (long long k = rlo; k < rhi; ++k) {
      const uint32_t h = ((uint32_t)__ldg(rated_cols + k) * 0x9E3779B1u) >> 23;
      sig[(h >> 5) * TM + rowl] |= 1u << (h & 31);
    }
  }

  if (tid == 0) {
    for (int b = 0; b < NBUF; ++b) {
      mbar_init(&mma_done[b], 1);
      mbar_init(&acc_free[b], NT / 32);
    }
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "n"(NBUF * TN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // ---- the users' rows, split and stored once: float4 number q of the tile is (row q / (D/4), columns 4 * (q % (D/4)))
#pragma unroll
  for (int p = 0; p < APT; ++p) {
    if (tid >= NT) break;                             // the 9th warp stages nothing
    const int q = tid + NT * p;
    const int row = q / (D / 4), c4 = (q % (D / 4)) * 4;
    const int ur = (row0 + row < n_rows) ? __ldg(user_ids + row0 + row) : -1;
    const float4 v = (ur >= 0 && c4 < d) ? __ldg(reinterpret_cast<const float4*>(U + (size_t)ur * d + c4)) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 hi, lo;
    split_tf32(v, hi, lo);
    const uint32_t off = (uint32_t)(c4 >> 5) * KBLK + sw_off(row, c4 & 31);
    *reinterpret_cast<float4*>(sA_hi + off) = hi;
    *reinterpret_cast<float4*>(sA_lo + off) = lo;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");        // the A operands: generic-proxy writes -> async proxy (UMMA)
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_acc = tmem_base_slot;
  const uint32_t idesc = make_idesc();

  const int n_tiles = (n_items + TN - 1) / TN;
  int cnt = 0;                                        // this thread's row: candidates in its list, current cut-off
  unsigned long long thr = 0ULL;

  // Row `src` of this warp goes back to its N best: the N-th largest of its c keys by a bit-wise search on the keys
  // held in registers (keys are distinct: the item id is part of the key), then the keys at or above it move to the
  // front of the list.  The owner's count and cut-off are updated.
  auto compact_row = [&](int src) {
    const int c = __shfl_sync(0xffffffffu, cnt, src);
    unsigned long long* list = cand + ((size_t)(wq * 32 + src) * 2 + half) * CAP;
    unsigned long long k[CAP / 32];
#pragma unroll
    for (int i = 0; i < CAP / 32; ++i) k[i] = (lane + 32 * i < c) ? __ldcg(list + lane + 32 * i) : 0ULL;
    // high words first (the score): the largest T with at least N keys whose high word is >= T
    uint32_t th = 0;
#pragma unroll 1
    for (int b = 31; b >= 0; --b) {
      const uint32_t t1 = th | (1u << b);
      int n = 0;
#pragma unroll
      for (int i = 0; i < CAP / 32; ++i) n += ((uint32_t)(k[i] >> 32) >= t1) ? 1 : 0;
      if (__reduce_add_sync(0xffffffffu, n) >= N) th = t1;
    }
    int above = 0, equal = 0;
#pragma unroll
    for (int i = 0; i < CAP / 32; ++i) {
      above += ((uint32_t)(k[i] >> 32) > th) ? 1 : 0;
      equal += ((uint32_t)(k[i] >> 32) == th) ? 1 : 0;
    }
    above = __reduce_add_sync(0xffffffffu, above);
    equal = __reduce_add_sync(0xffffffffu, equal);
    uint32_t tl = 0;                                   // low word (inverted item id) of the N-th key among the ties
    if (equal > N - above) {                           // more keys tie on the score than fit: the smallest item ids win
#pragma unroll 1
      for (int b = 31; b >= 0; --b) {
        const uint32_t t1 = tl | (1u << b);
        int n = 0;
#pragma unroll
        for (int i = 0; i < CAP / 32; ++i) n += ((uint32_t)(k[i] >> 32) == th && (uint32_t)k[i] >= t1) ? 1 : 0;
        if (__reduce_add_sync(0xffffffffu, n) >= N - above) tl = t1;
      }
    }
    const unsigned long long cut = ((unsigned long long)th << 32) | tl;     // keys >= cut: exactly N of them
    unsigned long long low = ~0ULL;
    int base = 0;
    __syncwarp();
#pragma unroll
    for (int i = 0; i < CAP / 32; ++i) {
      const bool keep = k[i] >= cut && k[i] != 0ULL;
      const unsigned m = __ballot_sync(0xffffffffu, keep);
      if (keep) {
        __stcg(list + base + __popc(m & ((1u << lane) - 1u)), k[i]);
        low = k[i] < low ? k[i] : low;
      }
      base += __popc(m);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, low, o);
      low = other < low ? other : low;
    }
    if (lane == src) { cnt = base; thr = low; }        // base == N; low = the N-th best key
    __syncwarp();
  };
  // selection over one finished accumulator (tile t, TMEM buffer t & 1)
  auto select_tile = [&](int t) {
    // a row that could overflow during this tile goes back to its N best first (its warp works on it together)
    unsigned need = __ballot_sync(0xffffffffu, cnt > N + TRIG_EXTRA);
    while (need) {
      const int src = __ffs(need) - 1;
      need &= need - 1;
      compact_row(src);
    }
    // pre-filter for the 128 scores of this tile (the cut-off only moves in the compaction above): a score below the
    // cut-off's score cannot pass, unless the list is not full yet or a rated item's fixed value could pass
    const bool open_row = thr == 0ULL || (uint32_t)(thr >> 32) <= (uint32_t)(rated_key_hi >> 32);
    const float thr_f = open_row ? -INFINITY : score_of((uint32_t)(thr >> 32));
    const int buf = t % NBUF;
    mbar_wait(&mma_done[buf], (uint32_t)((t / NBUF) & 1));
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int c0 = t * TN + half * 64;                                // first item of this thread's 64 columns
    const int valid = n_items - c0;                                    // columns of them that are items (may be <= 0 or > 64)
#pragma unroll 1
    for (int cc = 0; cc < 64; cc += 16) {
      uint32_t r[16];
      __syncwarp();                                   // the rare path below diverges; tcgen05.ld is warp-collective
      const uint32_t taddr = tmem_acc + ((uint32_t)(wq * 32) << 16) + (uint32_t)(buf * TN + half * 64 + cc);
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
            "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      // 16 compares without a branch: bit q of `pass` <=> score q is at or above the cut-off's score
      uint32_t pass = 0;
#pragma unroll
      for (int q = 0; q < 16; ++q) pass |= (__uint_as_float(r[q]) >= thr_f ? 1u : 0u) << q;
      if (valid - cc < 16) pass &= (valid - cc) <= 0 ? 0u : ((1u << (valid - cc)) - 1u);
      if (u < 0) pass = 0;
      if (pass) {                                     // the few that pass: exact 64-bit test, rated test on a signature hit
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          if (pass & (1u << q)) {
            const int c = c0 + cc + q;
            const unsigned long long low = (unsigned long long)(0xffffffffu - (uint32_t)c);
            unsigned long long key = ((unsigned long long)ord_of(__uint_as_float(r[q])) << 32) | low;
            // a rated item scores `rated_value` whatever its dot product: it can pass even when the raw score does not
            if (key > thr || (rated_key_hi | low) > thr) {
              const uint32_t h = ((uint32_t)c * 0x9E3779B1u) >> 23;
              if (((sig[(h >> 5) * TM + rowl] >> (h & 31)) & 1u) && is_rated(rated_cols, rlo, rhi, c)) key = rated_key_hi | low;
              if (key > thr) {
                __stcg(my_cand + cnt, key);           // cnt < CAP by the compaction rule
                ++cnt;
              }
            }
          }
        }
      }
    }
    // the accumulator may be overwritten once all 8 warps are past their loads: each warp says so on acc_free
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncwarp();
    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&acc_free[buf])) : "memory");
  };

  // ---- main loop.  The 9th warp's lane 0 drives the two engines -- bulk copies (tile j into the operand stage that
  // MMA(j-2) has released) and the 24 MMAs of tile j into accumulator j % 4 once the 8 selecting warps have released
  // it -- and never selects; the selecting warps follow at their own pace (no CTA-wide barrier per tile: a warp that
  // compacts a list only holds back the accumulator it has not released yet).
  constexpr uint32_t TILE_BYTES = 2 * OPER;
  if (warp == NT / 32) {
    if (lane == 0) {
      for (int j = 0; j < n_tiles; ++j) {
        const int s = j & 1, b = j % NBUF;
        if (j >= 2) mbar_wait(&mma_done[(j - 2) % NBUF], (uint32_t)(((j - 2) / NBUF) & 1));   // MMA(j-2) done: stage s is free
        mbar_expect_tx(&full[s], TILE_BYTES);
        bulk_g2s(sB + s * TILE_BYTES, item_blocks + (size_t)j * TILE_BYTES, TILE_BYTES, &full[s]);
        if (j >= NBUF) mbar_wait(&acc_free[b], (uint32_t)((j / NBUF - 1) & 1));              // select(j-4) done everywhere
        mbar_wait(&full[s], (uint32_t)((j >> 1) & 1));                                        // tile j has landed
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint8_t* const sB_hi = sB + s * TILE_BYTES;
        uint8_t* const sB_lo = sB_hi + OPER;
        const uint32_t acc_addr = tmem_acc + (uint32_t)(b * TN);
        bool first = true;
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) {
          const uint64_t a_hi = make_desc(smem_u32(sA_hi + kb * KBLK)), a_lo = make_desc(smem_u32(sA_lo + kb * KBLK));
          const uint64_t b_hi = make_desc(smem_u32(sB_hi + kb * KBLK)), b_lo = make_desc(smem_u32(sB_lo + kb * KBLK));
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            const uint64_t step = (uint64_t)(k4 * 2);   // +2 = 32 bytes (8 tf32) along K inside the 128-byte span
#pragma unroll
            for (int term = 0; term < 3; ++term) {      // small terms first: lo.hi, hi.lo, then hi.hi
              const uint64_t da = (term == 0 ? a_lo : a_hi) + step;
              const uint64_t db = (term == 1 ? b_lo : b_hi) + step;
              const uint32_t accf = first ? 0u : 1u;
              first = false;
              asm volatile(
                  "{\n\t.reg .pred p;\n\t"
                  "setp.ne.b32 p, %4, 0;\n\t"
                  "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(acc_addr), "l"(da), "l"(db), "r"(idesc),
                  "r"(accf)
                  : "memory");
            }
          }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mma_done[b]))
                     : "memory");
      }
    }
  } else {
    for (int t = 0; t < n_tiles; ++t) select_tile(t);
  }

  // ---- final: every list down to at most N keys, then the two lists of a row merged, sorted and written
  __syncwarp();
  if (warp < NT / 32) {
    unsigned need = __ballot_sync(0xffffffffu, cnt > N);
    while (need) {
      const int src = __ffs(need) - 1;
      need &= need - 1;
      compact_row(src);
    }
  }
  if (warp < NT / 32) cnt_sh[half][rowl] = cnt;
  __syncthreads();                                    // both halves' lists (global, st.cg) and counts are visible
  for (int src = half; src < 32 && warp < NT / 32; src += 2) {   // the two warps of a lane quarter share its 32 rows
    const int r = wq * 32 + src;
    const int ur = __shfl_sync(0xffffffffu, u, src);
    if (ur < 0) continue;
    const int ca = cnt_sh[0][r], cb = cnt_sh[1][r];   // <= N each
    const unsigned long long* la = cand + (size_t)r * 2 * CAP;
    const unsigned long long* lb = la + CAP;
    __syncwarp();
    for (int k = lane; k < SORTN; k += 32) my_sort[k] = k < ca ? __ldcg(la + k) : (k - ca < cb ? __ldcg(lb + (k - ca)) : 0ULL);
    warp_sort_desc<SORTN>(my_sort, lane);
    const size_t orow = (size_t)(row0 + r) * N;
    for (int k = lane; k < N; k += 32) {
      const unsigned long long key = my_sort[k];
      out_ids[orow + k] = (int)(0xffffffffu - (uint32_t)(key & 0xffffffffULL));
      out_scores[orow + k] = score_of((uint32_t)(key >> 32));
    }
    __syncwarp();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "n"(NBUF * TN));
  }
}

template <int KB>
int launch_tc(const float* U, const float* V, int d, int n_items, const int* user_ids, int n_rows, const long long* rowptr,
              const int* cols, float rated_value, int N, int* out_ids, float* out_scores, cudaStream_t st) {
  constexpr int SMEM = 6 * KB * KBLK + 8 * SORTN * (int)sizeof(unsigned long long) + SIGW * TM * (int)sizeof(uint32_t) + 1024;   // operands + sort buffers + signatures + alignment
  static bool attr_set = false;
  if (!attr_set) {
    QREC_CUDA(cudaFuncSetAttribute(score_topn_tc_kernel<KB>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    attr_set = true;
  }
  const int grid = (n_rows + TM - 1) / TM;
  const int n_tiles = (n_items + TN - 1) / TN;
  const size_t list_bytes = (size_t)grid * TM * 2 * CAP * sizeof(unsigned long long);   // candidate lists: 2 x 2.5 KB per user
  const size_t block_bytes = (size_t)n_tiles * 2 * KB * KBLK;                         // the split item table, tile by tile
  uint8_t* ws = nullptr;                                                              // stream-ordered scratch
  QREC_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&ws), list_bytes + block_bytes, st));
  uint8_t* const blocks = ws + list_bytes;                                            // (list_bytes is a multiple of 640 KB)
  split_items_kernel<KB><<<n_tiles, 128, 0, st>>>(V, d, n_items, blocks);
  score_topn_tc_kernel<KB><<<grid, NT + 32, SMEM, st>>>(U, blocks, d, n_items, user_ids, n_rows, rowptr, cols, rated_value, N, out_ids,
                                                    out_scores, reinterpret_cast<unsigned long long*>(ws));
  const cudaError_t launch_err = cudaGetLastError();
  QREC_CUDA(cudaFreeAsync(ws, st));
  if (launch_err != cudaSuccess) return qrec::cuda_fail(launch_err, "kernel launch", __FILE__, __LINE__);
  qrec::count_launch(2);
  return QREC_OK;
}

}  // namespace

extern "C" int qrec_score_topn_tc_f32(const float* dev_U, const float* dev_V, int32_t d, int32_t n_items,
                                      const int32_t* dev_user_ids, int32_t n_rows, const int64_t* dev_rated_rowptr,
                                      const int32_t* dev_rated_cols, float rated_value, int32_t N, int32_t* dev_out_ids,
                                      float* dev_out_scores, void* stream) {
  QREC_REQUIRE(n_rows >= 0 && n_items >= 1, "qrec_score_topn_tc_f32: bad size");
  QREC_REQUIRE(d >= 4 && d <= 64 && d % 4 == 0, "qrec_score_topn_tc_f32: d=%d unsupported (multiple of 4, <= 64; use qrec_score_topn_f32)", d);
  QREC_REQUIRE(N >= 1 && N <= NMAX && N <= n_items, "qrec_score_topn_tc_f32: N=%d must be in 1..min(%d, n_items)", N, NMAX);
  if (n_rows == 0) return QREC_OK;
  QREC_REQUIRE(dev_U && dev_V && dev_user_ids && dev_rated_rowptr && dev_rated_cols && dev_out_ids && dev_out_scores,
               "qrec_score_topn_tc_f32: null pointer");
  QREC_REQUIRE((reinterpret_cast<uintptr_t>(dev_U) & 15) == 0 && (reinterpret_cast<uintptr_t>(dev_V) & 15) == 0,
               "qrec_score_topn_tc_f32: tables must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const long long* rp = reinterpret_cast<const long long*>(dev_rated_rowptr);
  if (d <= 32)                                            // columns beyond d are zero-filled up to the 32-wide k-block
    return launch_tc<1>(dev_U, dev_V, d, n_items, dev_user_ids, n_rows, rp, dev_rated_cols, rated_value, N, dev_out_ids, dev_out_scores, st);
  return launch_tc<2>(dev_U, dev_V, d, n_items, dev_user_ids, n_rows, rp, dev_rated_cols, rated_value, N, dev_out_ids, dev_out_scores, st);
}
