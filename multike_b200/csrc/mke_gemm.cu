// Tensor-core GEMM for the literal auto-encoder (code/literal_encoder.py:19-144: six affine layers
// 1500 -> 1024 -> 512 -> dim -> 512 -> 1024 -> 1500, forward and backward): hand-written tcgen05 (UMMA) kernel,
// TMA-fed, TMEM accumulator, at fp32-EQUIVALENT precision by the 3xTF32 split.
//
//   C[M, N] = A[M, K] . B[N, K]^T (+ bias[N])          A, B row-major with K contiguous ("K-major"), fp32
//
// Precision: the tensor cores take TF32 operands (10-bit mantissa) and accumulate in fp32.  Every operand x is
// split on the way in into x_hi = x with the low 13 mantissa bits cleared and x_lo = x - x_hi (exact in fp32; the
// hardware reads its top 10 mantissa bits), and the product is accumulated as a_hi b_hi + a_hi b_lo + a_lo b_hi
// the dropped terms are below 2^-21 relative -- the rounding level of an fp32 FMA chain.  The tensor core adds
// into its fp32 accumulator with truncation, once per instruction, so the error of ONE accumulator taking all three
// products grows linearly with K (measured at K = 1500: 7x the error of an fp32 cuBLAS GEMM); the large terms
// a_hi b_hi therefore have an accumulator of their own (a third of the truncations) and the cross terms, 2^-11 smaller,
// a second one whose truncations do not matter; and K is summed inside the tensor core only in chunks of 512: the
// epilogue warps drain the (double-buffered) accumulators chunk by chunk into fp32 registers with round-to-nearest
// adds (tests/test_gpu_gemm.py: against fp64, the error of torch's fp32 cuBLAS GEMM on the same operands).
//
// Structure (one 128 x 128 tile of C per CTA, K in blocks of 32 floats = one 128-byte swizzle row):
//   warp 0, one lane : TMA producer -- four boxes per stage (A_hi, A_lo, B_hi, B_lo; 128 rows x 128 B, SWIZZLE_128B,
//                      out-of-range rows / columns arrive as zeros), completion on the stage's `full` mbarrier
//   warp 1, one lane : MMA issuer -- per stage 4 k-steps x 3 tcgen05.mma.kind::tf32 (M = N = 128, K = 8) from
//                      shared-memory descriptors into 2 x 128 TMEM columns (two such buffers, alternating per chunk of
//                      16 k-blocks); tcgen05.commit frees the stage (`empty`) and hands a finished chunk to the epilogue
//   warps 2-5        : epilogue -- per chunk tcgen05.ld (32 lanes x 32 columns per instruction) into a register row of
//                      128 sums; at the end + bias, bounds-checked stores; warp w reads TMEM lanes 32 (w mod 4) ..
// Every mbarrier wait is bounded (trap, do not hang).
#include <cstdlib>
#include "mke_umma.cuh"

namespace mke {

constexpr int kGemmBM = 128, kGemmBN = 128, kGemmBK = 32, kGemmStages = 3;
constexpr int kGemmTileBytes = 128 * kGemmBK * 4;           // one operand box
constexpr int kGemmStageBytes = 4 * kGemmTileBytes;         // A_hi, A_lo, B_hi, B_lo
constexpr int kGemmThreads = 192;
constexpr int kGemmTmemCols = 512;  // 2 buffers x (a_hi b_hi | cross terms) x 128 fp32 columns
constexpr int kGemmChunk = 16;      // k-blocks (of 32) summed inside the tensor core before the accumulator is drained

struct GemmParams {
  int M, N, K;
  const float* bias;  // [N] or NULL
  float* C;
  long long ldc;
  int kb_per_split;   // k-blocks (of 32) per CTA along grid z (split-K); >= all of them: no split
  int accumulate;     // 1: C += result with atomic adds (split-K partial sums, or a gradient accumulator)
};

__global__ void __launch_bounds__(kGemmThreads, 1)
    gemm_tf32x3_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                       const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                       const GemmParams p) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);  // SWIZZLE_128B: 1024-byte tiles
  uint64_t* bars = (uint64_t*)(smem + kGemmStages * kGemmStageBytes);  // full[S], empty[S], accum_full[2], accum_empty[2]
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * kGemmStages + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * kGemmBM, n0 = blockIdx.y * kGemmBN;
  const int all_kb = (p.K + kGemmBK - 1) / kGemmBK;
  const int kb0 = blockIdx.z * p.kb_per_split;                                   // this CTA's K range (split-K)
  const int num_kb = all_kb - kb0 < p.kb_per_split ? all_kb - kb0 : p.kb_per_split;
  auto full = [&](int s) { return smem_u32(bars + s); };
  auto empty = [&](int s) { return smem_u32(bars + kGemmStages + s); };
  auto accum_full = [&](int b) { return smem_u32(bars + 2 * kGemmStages + b); };
  auto accum_empty = [&](int b) { return smem_u32(bars + 2 * kGemmStages + 2 + b); };
  const int num_chunks = (num_kb + kGemmChunk - 1) / kGemmChunk;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < kGemmStages; ++s) {
      mbar_init(full(s), 1);
      mbar_init(empty(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(accum_full(b), 1);
      mbar_init(accum_empty(b), 4);  // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {  // one warp allocates the accumulator columns (and frees them at the end)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(kGemmTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {  // ---- TMA producer ----
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % kGemmStages;
        const uint32_t ph = (uint32_t)(kb / kGemmStages) & 1u;
        mbar_wait(empty(s), ph ^ 1u);
        mbar_expect_tx(full(s), kGemmStageBytes);
        const uint32_t dst = smem_u32(smem + s * kGemmStageBytes);
        tma_load_2d(&map_a_hi, full(s), dst + 0 * kGemmTileBytes, (kb0 + kb) * kGemmBK, m0);
        tma_load_2d(&map_a_lo, full(s), dst + 1 * kGemmTileBytes, (kb0 + kb) * kGemmBK, m0);
        tma_load_2d(&map_b_hi, full(s), dst + 2 * kGemmTileBytes, (kb0 + kb) * kGemmBK, n0);
        tma_load_2d(&map_b_lo, full(s), dst + 3 * kGemmTileBytes, (kb0 + kb) * kGemmBK, n0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ---- MMA issuer ----
      for (int c = 0; c < num_chunks; ++c) {
        const int buf = c & 1;
        mbar_wait(accum_empty(buf), ((uint32_t)(c >> 1) & 1u) ^ 1u);  // the epilogue has drained this buffer
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d_main = tmem_base + (uint32_t)(buf * 2 * kGemmBN), d_cross = d_main + kGemmBN;
        const int kb_end = min(num_kb, (c + 1) * kGemmChunk);
        for (int kb = c * kGemmChunk; kb < kb_end; ++kb) {
          const int s = kb % kGemmStages;
          const uint32_t ph = (uint32_t)(kb / kGemmStages) & 1u;
          mbar_wait(full(s), ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t base = smem_u32(smem + s * kGemmStageBytes);
          const uint64_t a_hi = umma_desc_k_sw128(base + 0 * kGemmTileBytes), a_lo = umma_desc_k_sw128(base + 1 * kGemmTileBytes);
          const uint64_t b_hi = umma_desc_k_sw128(base + 2 * kGemmTileBytes), b_lo = umma_desc_k_sw128(base + 3 * kGemmTileBytes);
#pragma unroll
          for (int k = 0; k < kGemmBK / 8; ++k) {
            const uint64_t adv = (uint64_t)((k * 8 * 4) >> 4);  // 8 TF32 elements = 32 bytes along K inside the swizzle row
            const uint32_t acc = (kb > c * kGemmChunk || k > 0) ? 1u : 0u;  // a chunk starts from zero
            umma_tf32(d_main, a_hi + adv, b_hi + adv, acc);
            umma_tf32(d_cross, a_hi + adv, b_lo + adv, acc);
            umma_tf32(d_cross, a_lo + adv, b_hi + adv, 1u);
          }
          umma_commit(empty(s));  // the stage is free once these MMAs have read it
        }
        umma_commit(accum_full(buf));  // this chunk's partial sums are complete
      }
    }
  } else {  // ---- epilogue: warps 2..5, TMEM lanes 32 (warp mod 4) .. ----
    const int lane_base = (warp & 3) * 32;
    const int row = m0 + lane_base + lane;
    float acc[kGemmBN];  // this thread's row of the tile, summed over the chunks with round-to-nearest adds
#pragma unroll
    for (int j = 0; j < kGemmBN; ++j) acc[j] = 0.f;
#pragma unroll 1
    for (int c = 0; c < num_chunks; ++c) {
      const int buf = c & 1;
      mbar_wait(accum_full(buf), (uint32_t)(c >> 1) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem_base + ((uint32_t)lane_base << 16) + (uint32_t)(buf * 2 * kGemmBN);
#pragma unroll
      for (int q = 0; q < kGemmBN / 32; ++q) {
        uint32_t v[32], u[32];
        tmem_ld32(taddr + q * 32, v);            // sum of a_hi b_hi
        tmem_ld32(taddr + kGemmBN + q * 32, u);  // sum of a_hi b_lo + a_lo b_hi
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[q * 32 + j] += __uint_as_float(v[j]) + __uint_as_float(u[j]);
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(accum_empty(buf)) : "memory");
    }
    if (row < p.M) {
      float* out = p.C + (size_t)row * p.ldc + n0;
#pragma unroll
      for (int j = 0; j < kGemmBN; ++j) {
        const int col = n0 + j;
        if (col < p.N) {
          const float v = acc[j] + ((p.bias != nullptr && blockIdx.z == 0) ? __ldg(p.bias + col) : 0.f);
          if (p.accumulate)
            atomicAdd(out + j, v);
          else
            out[j] = v;
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kGemmTmemCols) : "memory");
  }
}

// x -> (hi, lo): hi = x with the 13 low mantissa bits cleared (a TF32 value), lo = x - hi (exact)
__global__ void split_tf32_kernel(const float* __restrict__ x, float* __restrict__ hi, float* __restrict__ lo, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    split_tf32(x[i], hi[i], lo[i]);
  }
}

}  // namespace mke

using namespace mke;

extern "C" int mke_split_tf32(const float* x, float* hi, float* lo, int64_t n, mke_stream_t stream) {
  MKE_CHECK_ARG(n >= 0 && (n == 0 || (x && hi && lo)), "bad split arguments");
  if (n == 0) return 0;
  long long blocks = (n + 255) / 256;
  const long long full = (long long)sm_count() * 16;
  if (blocks > full) blocks = full;
  split_tf32_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, hi, lo, n);
  MKE_CHECK_LAUNCH("split_tf32_kernel");
  return 0;
}

namespace mke {
// C (+)= A . B^T (+ bias), optionally split along K over grid z (then C must accumulate)
int gemm_tf32x3_launch(const float* a_hi, const float* a_lo, int64_t lda, const float* b_hi, const float* b_lo, int64_t ldb,
                       int M, int N, int K, const float* bias_or_null, float* C, int64_t ldc, int k_splits, int accumulate,
                       cudaStream_t stream) {
  MKE_CHECK_ARG(M > 0 && N > 0 && K > 0, "bad GEMM shape %d x %d x %d", M, N, K);
  MKE_CHECK_ARG(a_hi && a_lo && b_hi && b_lo && C, "null operand");
  MKE_CHECK_ARG(lda >= K && ldb >= K && ldc >= N, "leading dimensions");
  MKE_CHECK_ARG(lda % 4 == 0 && ldb % 4 == 0, "lda / ldb must be multiples of 4 floats (TMA: 16-byte row pitch)");
  MKE_CHECK_ARG(((uintptr_t)a_hi | (uintptr_t)a_lo | (uintptr_t)b_hi | (uintptr_t)b_lo) % 16 == 0, "operands must be 16-byte aligned");
  MKE_CHECK_ARG(k_splits >= 1 && (k_splits == 1 || accumulate), "split-K needs an accumulating epilogue");
  CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo;
  if (int rc = make_map(&ma_hi, a_hi, M, K, lda)) return rc;
  if (int rc = make_map(&ma_lo, a_lo, M, K, lda)) return rc;
  if (int rc = make_map(&mb_hi, b_hi, N, K, ldb)) return rc;
  if (int rc = make_map(&mb_lo, b_lo, N, K, ldb)) return rc;
  constexpr int smem = kGemmStages * kGemmStageBytes + 1024 + 256;
  static bool configured = false;
  if (!configured) {
    if (cudaError_t e = cudaFuncSetAttribute(gemm_tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem))
      return cuda_fail(e, "cudaFuncSetAttribute(gemm_tf32x3_kernel)");
    configured = true;
  }
  const int all_kb = (K + kGemmBK - 1) / kGemmBK;
  if (k_splits > all_kb) k_splits = all_kb;
  const int per = (all_kb + k_splits - 1) / k_splits;
  k_splits = (all_kb + per - 1) / per;  // no empty split
  const GemmParams p{M, N, K, bias_or_null, C, (long long)ldc, per, accumulate};
  const dim3 grid((M + kGemmBM - 1) / kGemmBM, (N + kGemmBN - 1) / kGemmBN, k_splits);
  gemm_tf32x3_kernel<<<grid, kGemmThreads, smem, stream>>>(ma_hi, ma_lo, mb_hi, mb_lo, p);
  MKE_CHECK_LAUNCH("gemm_tf32x3_kernel");
  return 0;
}
}  // namespace mke

extern "C" int mke_gemm_tf32x3(const float* a_hi, const float* a_lo, int64_t lda, const float* b_hi, const float* b_lo,
                               int64_t ldb, int32_t M, int32_t N, int32_t K, const float* bias_or_null, float* C,
                               int64_t ldc, mke_stream_t stream) {
  return gemm_tf32x3_launch(a_hi, a_lo, lda, b_hi, b_lo, ldb, M, N, K, bias_or_null, C, ldc, 1, 0, (cudaStream_t)stream);
}
