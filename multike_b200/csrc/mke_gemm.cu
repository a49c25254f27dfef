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
#include <cuda.h>
#include <cstdlib>
#include "mke_common.cuh"

namespace mke {

constexpr int kGemmBM = 128, kGemmBN = 128, kGemmBK = 32, kGemmStages = 3;
constexpr int kGemmTileBytes = 128 * kGemmBK * 4;           // one operand box
constexpr int kGemmStageBytes = 4 * kGemmTileBytes;         // A_hi, A_lo, B_hi, B_lo
constexpr int kGemmThreads = 192;
constexpr int kGemmTmemCols = 512;  // 2 buffers x (a_hi b_hi | cross terms) x 128 fp32 columns
constexpr int kGemmChunk = 16;      // k-blocks (of 32) summed inside the tensor core before the accumulator is drained
constexpr unsigned long long kGemmWaitNs = 4000000000ull;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned long long gemm_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)::"memory");
  return t;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  const unsigned long long t0 = gemm_now();
  while (true) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
    if (gemm_now() - t0 > kGemmWaitNs) __trap();
  }
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// shared-memory matrix descriptor, K-major operand, SWIZZLE_128B (cute::UMMA::SmemDescriptor): start address >> 4,
// leading byte offset (unused for swizzled K-major) = 1, stride byte offset = 8 rows x 128 B = 1024 >> 4, version 1
// (Blackwell), layout type 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 (bits 4-5 = 1), A = B = TF32 (bits 7-9, 10-12 = 2),
// both K-major (bits 15, 16 = 0), N >> 3 at bit 17, M >> 4 at bit 24
constexpr uint32_t kIdescTf32 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kGemmBN >> 3) << 17) |
                                ((uint32_t)(kGemmBM >> 4) << 24);
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(kIdescTf32), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// 32 lanes x 32 consecutive columns of an fp32 accumulator: thread = lane (row of the tile), v[j] = column j
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}

struct GemmParams {
  int M, N, K;
  const float* bias;  // [N] or NULL
  float* C;
  long long ldc;
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
  const int num_kb = (p.K + kGemmBK - 1) / kGemmBK;
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
        tma_load_2d(&map_a_hi, full(s), dst + 0 * kGemmTileBytes, kb * kGemmBK, m0);
        tma_load_2d(&map_a_lo, full(s), dst + 1 * kGemmTileBytes, kb * kGemmBK, m0);
        tma_load_2d(&map_b_hi, full(s), dst + 2 * kGemmTileBytes, kb * kGemmBK, n0);
        tma_load_2d(&map_b_lo, full(s), dst + 3 * kGemmTileBytes, kb * kGemmBK, n0);
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
        if (col < p.N) out[j] = acc[j] + (p.bias != nullptr ? __ldg(p.bias + col) : 0.f);
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
    const float v = x[i];
    const float h = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    hi[i] = h;
    lo[i] = v - h;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}
// row-major [rows, K] fp32 with leading dimension ld: box = 32 floats of K x 128 rows, 128-byte swizzle
static int make_map(CUtensorMap* map, const float* base, int rows, int K, long long ld) {
  EncodeTiledFn enc = encode_tiled();
  if (enc == nullptr) {
    set_error("cuTensorMapEncodeTiled is not available from this driver");
    return MKE_EINVAL;
  }
  const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  const cuuint32_t box[2] = {(cuuint32_t)kGemmBK, 128u};
  const cuuint32_t estr[2] = {1u, 1u};
  const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) for a [%d, %d] operand with ld %lld", (int)r, rows, K, ld);
    return MKE_EINVAL;
  }
  return 0;
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

extern "C" int mke_gemm_tf32x3(const float* a_hi, const float* a_lo, int64_t lda, const float* b_hi, const float* b_lo,
                               int64_t ldb, int32_t M, int32_t N, int32_t K, const float* bias_or_null, float* C,
                               int64_t ldc, mke_stream_t stream) {
  MKE_CHECK_ARG(M > 0 && N > 0 && K > 0, "bad GEMM shape %d x %d x %d", M, N, K);
  MKE_CHECK_ARG(a_hi && a_lo && b_hi && b_lo && C, "null operand");
  MKE_CHECK_ARG(lda >= K && ldb >= K && ldc >= N, "leading dimensions");
  MKE_CHECK_ARG(lda % 4 == 0 && ldb % 4 == 0, "lda / ldb must be multiples of 4 floats (TMA: 16-byte row pitch)");
  MKE_CHECK_ARG(((uintptr_t)a_hi | (uintptr_t)a_lo | (uintptr_t)b_hi | (uintptr_t)b_lo) % 16 == 0, "operands must be 16-byte aligned");
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
  const GemmParams p{M, N, K, bias_or_null, C, (long long)ldc};
  const dim3 grid((M + kGemmBM - 1) / kGemmBM, (N + kGemmBN - 1) / kGemmBN);
  gemm_tf32x3_kernel<<<grid, kGemmThreads, smem, (cudaStream_t)stream>>>(ma_hi, ma_lo, mb_hi, mb_lo, p);
  MKE_CHECK_LAUNCH("gemm_tf32x3_kernel");
  return 0;
}
