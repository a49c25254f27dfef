// Microbenchmark of the row gather / scatter-add primitives the fused relation kernel can be
// built from, at the shape of BASELINE config 2 (200 000 rows x 75 fp32, stride 80; 20 000
// positives x 13 rows per launch).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
// -o gpurun_out/mb profiles/microbench/scatter_gather.cu ; run on the GPU box.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int ROWS_PER_POS = 13;
constexpr int NCHUNK = 19;  // float4 pieces that carry data (75 floats)

__device__ __forceinline__ void red_v4(float* p, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void red_v2(float* p, float x, float y) {
  asm volatile("red.global.add.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(x), "f"(y) : "memory");
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// mode 0: gather LDG.128 (19 lanes / row), sum into sink
// mode 1: scatter RED.v4     mode 2: scatter RED.f32 (3 rounds of 32 lanes)   mode 3: plain ST.v4
// mode 4: RED.v2 (38 lanes -> 2 rounds)
template <int MODE>
__global__ void __launch_bounds__(256) k_reg(const float* __restrict__ tab, float* __restrict__ grad,
                                             const int32_t* __restrict__ idx, int npos, int stride, float* sink) {
  const int lane = threadIdx.x & 31;
  const int gw = blockIdx.x * 8 + (threadIdx.x >> 5), nw = gridDim.x * 8;
  float acc = 0.f;
  for (int i = gw; i < npos; i += nw) {
    int32_t my = (lane < ROWS_PER_POS) ? __ldg(idx + (size_t)i * ROWS_PER_POS + lane) : 0;
    if (MODE == 0) {
      float4 v[ROWS_PER_POS];
#pragma unroll
      for (int j = 0; j < ROWS_PER_POS; ++j) {
        const int32_t r = __shfl_sync(0xffffffffu, my, j);
        v[j] = (lane < NCHUNK) ? __ldg(reinterpret_cast<const float4*>(tab + (size_t)r * stride) + lane) : make_float4(0, 0, 0, 0);
      }
#pragma unroll
      for (int j = 0; j < ROWS_PER_POS; ++j) acc += v[j].x + v[j].y + v[j].z + v[j].w;
    } else {
#pragma unroll
      for (int j = 0; j < ROWS_PER_POS; ++j) {
        const int32_t r = __shfl_sync(0xffffffffu, my, j);
        float* g = grad + (size_t)r * stride;
        if (MODE == 1) { if (lane < NCHUNK) red_v4(g + 4 * lane, make_float4(1.f, 1.f, 1.f, 1.f)); }
        if (MODE == 2) { for (int c = lane; c < 75; c += 32) atomicAdd(g + c, 1.f); }
        if (MODE == 3) { if (lane < NCHUNK) *reinterpret_cast<float4*>(g + 4 * lane) = make_float4(1.f, 1.f, 1.f, 1.f); }
        if (MODE == 4) { for (int c = lane; c < 38; c += 32) red_v2(g + 2 * c, 1.f, 1.f); }
      }
    }
  }
  if (MODE == 0 && acc == 123.456f) *sink = acc;
}

// mode 7: gather a row with LDG.128 and RED.v4 it into the gradient table (both directions in one
//         kernel, the access pattern of the fused relation kernel without its arithmetic)
// mode 8: the same in the quarter-warp layout (8 lanes x (v4,v4,v2) per row, 4 rows per warp op)
template <int MODE>
__global__ void __launch_bounds__(256) k_both(const float* __restrict__ tab, float* __restrict__ grad,
                                              const int32_t* __restrict__ idx, int npos, int stride) {
  const int lane = threadIdx.x & 31;
  if (MODE == 7) {
    const int gw = blockIdx.x * 8 + (threadIdx.x >> 5), nw = gridDim.x * 8;
    for (int i = gw; i < npos; i += nw) {
      int32_t my = (lane < ROWS_PER_POS) ? __ldg(idx + (size_t)i * ROWS_PER_POS + lane) : 0;
      float4 v[ROWS_PER_POS];
#pragma unroll
      for (int j = 0; j < ROWS_PER_POS; ++j) {
        const int32_t r = __shfl_sync(0xffffffffu, my, j);
        v[j] = (lane < NCHUNK) ? __ldg(reinterpret_cast<const float4*>(tab + (size_t)r * stride) + lane) : make_float4(0, 0, 0, 0);
      }
#pragma unroll
      for (int j = 0; j < ROWS_PER_POS; ++j) {
        const int32_t r = __shfl_sync(0xffffffffu, my, j);
        if (lane < NCHUNK) red_v4(grad + (size_t)r * stride + 4 * lane, v[j]);
      }
    }
  } else {
    const int sub = lane & 7, q = lane >> 3;
    const int gq = (blockIdx.x * 8 + (threadIdx.x >> 5)) * 4 + q, nq = gridDim.x * 32;
    for (int i = gq; i < npos; i += nq) {
      for (int j = 0; j < ROWS_PER_POS; ++j) {
        const int32_t r = __ldg(idx + (size_t)i * ROWS_PER_POS + j);
        const float* row = tab + (size_t)r * stride;
        float* g = grad + (size_t)r * stride;
        const float4 a = __ldg(reinterpret_cast<const float4*>(row) + sub);
        const float4 b = __ldg(reinterpret_cast<const float4*>(row) + 8 + sub);
        const float2 c = __ldg(reinterpret_cast<const float2*>(row + 64) + sub);
        red_v4(g + 4 * sub, a);
        red_v4(g + 32 + 4 * sub, b);
        red_v2(g + 64 + 2 * sub, c.x, c.y);
      }
    }
  }
}

// mode 5: bulk gather (cp.async.bulk g2s, 304 B / row, one mbarrier per warp, 13 rows in flight)
// mode 6: bulk reduce-add s2g (cp.reduce.async.bulk .add.f32, 304 B / row)
template <int MODE, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k_bulk(const float* __restrict__ tab, float* __restrict__ grad,
                                                     const int32_t* __restrict__ idx, int npos, int stride, float* sink) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t bars[WARPS];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int gw = blockIdx.x * WARPS + wib, nw = gridDim.x * WARPS;
  float* buf = reinterpret_cast<float*>(smem) + (size_t)wib * ROWS_PER_POS * stride;
  const uint32_t bar = smem_u32(&bars[wib]);
  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int c = lane; c < ROWS_PER_POS * stride; c += 32) buf[c] = 1.f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  uint32_t phase = 0;
  float acc = 0.f;
  for (int i = gw; i < npos; i += nw) {
    int32_t my = (lane < ROWS_PER_POS) ? __ldg(idx + (size_t)i * ROWS_PER_POS + lane) : 0;
    if (MODE == 5) {
      if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(ROWS_PER_POS * 304u) : "memory");
      __syncwarp();
      if (lane < ROWS_PER_POS)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(buf + lane * stride)), "l"(tab + (size_t)my * stride), "r"(304u), "r"(bar) : "memory");
      uint32_t done;
      do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p; }" : "=r"(done) : "r"(bar), "r"(phase) : "memory");
      } while (!done);
      phase ^= 1;
      acc += buf[lane];
    } else {
      if (lane < ROWS_PER_POS)
        asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(grad + (size_t)my * stride), "r"(smem_u32(buf + lane * stride)), "r"(304u) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
    }
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  if (acc == 123.456f) *sink = acc;
}

int main(int argc, char** argv) {
  const int rows = 200000, stride = 80, npos = 20000;
  const int n_rel = 550;
  const bool with_rel = argc > 1;  // "rel": row 1 of each positive comes from a 550-row hot table region
  float *tab, *grad, *sink;
  int32_t* idx;
  CK(cudaMalloc(&tab, (size_t)rows * stride * 4));
  CK(cudaMalloc(&grad, (size_t)rows * stride * 4));
  CK(cudaMalloc(&sink, 4));
  CK(cudaMemset(tab, 0, (size_t)rows * stride * 4));
  CK(cudaMemset(grad, 0, (size_t)rows * stride * 4));
  const int NSETS = 16;
  std::vector<int32_t> h((size_t)NSETS * npos * ROWS_PER_POS);
  uint64_t s = 88172645463325252ull;
  for (size_t k = 0; k < h.size(); ++k) {
    s ^= s << 13; s ^= s >> 7; s ^= s << 17;
    int32_t r = (int32_t)(s % rows);
    if (with_rel && (k % ROWS_PER_POS) == 1) {  // skewed relation ids: min of two uniforms squared-ish
      double u = (double)((s >> 20) % 100000) / 100000.0;
      r = (int32_t)(u * u * u * n_rel);
    }
    h[k] = r;
  }
  CK(cudaMalloc(&idx, h.size() * 4));
  CK(cudaMemcpy(idx, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  const double bytes = (double)npos * ROWS_PER_POS * 300.0;
  auto timeit = [&](const char* name, auto launch) {
    for (int w = 0; w < 3; ++w) launch(w % NSETS);
    CK(cudaDeviceSynchronize());
    const int iters = 48;
    CK(cudaEventRecord(e0));
    for (int it = 0; it < iters; ++it) launch(it % NSETS);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    CK(cudaGetLastError());
    printf("%-34s %8.2f us/launch  %8.1f GB/s (300 B/row, one direction)\n", name, 1e3 * ms / iters, bytes / (ms / iters * 1e-3) / 1e9);
  };
  for (int bps : {4, 8}) {
    const int grid = 148 * bps;
    printf("-- register kernels, grid = 148 x %d blocks of 256 thr%s\n", bps, with_rel ? " (row 1 = hot relation rows)" : "");
    timeit("gather LDG.128", [&](int q) { k_reg<0><<<grid, 256>>>(tab, grad, idx + (size_t)q * npos * ROWS_PER_POS, npos, stride, sink); });
    timeit("scatter RED.v4.f32", [&](int q) { k_reg<1><<<grid, 256>>>(tab, grad, idx + (size_t)q * npos * ROWS_PER_POS, npos, stride, sink); });
    timeit("scatter RED.v2.f32", [&](int q) { k_reg<4><<<grid, 256>>>(tab, grad, idx + (size_t)q * npos * ROWS_PER_POS, npos, stride, sink); });
    timeit("scatter RED.f32 (scalar)", [&](int q) { k_reg<2><<<grid, 256>>>(tab, grad, idx + (size_t)q * npos * ROWS_PER_POS, npos, stride, sink); });
    timeit("scatter ST.v4 (no add)", [&](int q) { k_reg<3><<<grid, 256>>>(tab, grad, idx + (size_t)q * npos * ROWS_PER_POS, npos, stride, sink); });
  }
  for (int bps : {4, 8}) {
    const int grid = 148 * bps;
    printf("-- gather + scatter in one kernel, grid = 148 x %d blocks of 256 thr\n", bps);
    timeit("LDG.128 + RED.v4, warp/positive", [&](int q) { k_both<7><<<grid, 256>>>(tab, grad, idx + (size_t)q * npos * ROWS_PER_POS, npos, stride); });
    timeit("LDG + RED, quarter/positive", [&](int q) { k_both<8><<<grid, 256>>>(tab, grad, idx + (size_t)q * npos * ROWS_PER_POS, npos, stride); });
  }
  {
    constexpr int W = 8;
    const size_t smem = (size_t)W * ROWS_PER_POS * stride * 4;  // 33 KB
    CK(cudaFuncSetAttribute(k_bulk<5, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(k_bulk<6, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int bps : {2, 4, 6}) {
      const int grid = 148 * bps;
      printf("-- bulk (TMA engine) kernels, grid = 148 x %d blocks of %d warps, %zu B smem\n", bps, W, smem);
      timeit("gather cp.async.bulk 304 B", [&](int q) { k_bulk<5, W><<<grid, W * 32, smem>>>(tab, grad, idx + (size_t)q * npos * ROWS_PER_POS, npos, stride, sink); });
      timeit("scatter cp.reduce.async.bulk add", [&](int q) { k_bulk<6, W><<<grid, W * 32, smem>>>(tab, grad, idx + (size_t)q * npos * ROWS_PER_POS, npos, stride, sink); });
    }
  }
  return 0;
}
