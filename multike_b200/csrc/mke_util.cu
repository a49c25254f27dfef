// Error plumbing, device queries and small table utilities.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include "mke_common.cuh"

namespace mke {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int cuda_fail(cudaError_t e, const char* what) {
  set_error("%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
  return -(int)e;
}
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

// out[i, 0:dim] = l2_normalize(var[idx[i]]) -- one warp per row
struct ExportShards {
  const float* var[MKE_MAX_SHARDS];
  ShardMap map;
  int sharded;  // 0: var[0] is the whole table
};
__global__ void table_export_kernel(ExportShards sh, int stride, int dim,
                                    int normalised, const int32_t* __restrict__ idx, int n,
                                    float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int i = blockIdx.x * wpb + (threadIdx.x >> 5); i < n; i += gridDim.x * wpb) {
    const int row = idx ? __ldg(idx + i) : i;
    int shard = 0;
    int32_t local = row;
    if (sh.sharded) sh.map.locate(row, shard, local);
    const float* pv = sh.var[shard] + (size_t)local * stride;
    float ss = 0.f;
    for (int c = lane; c < dim; c += 32) {
      const float x = pv[c];
      ss += x * x;
    }
    ss = warp_sum(ss);
    const float inv = normalised ? rsqrtf(fmaxf(ss, kNormEps)) : 1.f;
    for (int c = lane; c < dim; c += 32) out[(size_t)i * dim + c] = pv[c] * inv;
  }
}

__global__ void fill_rows_kernel(float* __restrict__ buf, size_t total, int stride, int dim,
                                 float value) {
  for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < total;
       k += (size_t)gridDim.x * blockDim.x)
    buf[k] = ((int)(k % stride) < dim) ? value : 0.f;
}

}  // namespace mke

using namespace mke;

extern "C" int mke_abi_version(void) { return MKE_ABI_VERSION; }
extern "C" const char* mke_last_error(void) { return g_err; }
extern "C" uint64_t mke_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" int mke_table_export(const mke_table_t* table, const int32_t* idx_or_null, int32_t n,
                                float* out, mke_stream_t stream) {
  MKE_CHECK_ARG(table && table->var && out, "null pointer");
  MKE_CHECK_ARG(n >= 0, "negative n");
  if (n == 0) return 0;
  int blocks = (n + 7) / 8;
  const int full = sm_count() * 8;
  if (blocks > full) blocks = full;
  ExportShards sh{};
  sh.var[0] = table->var;
  if (table->n_shards > 1) {
    sh.sharded = 1;
    sh.map = shard_map(table);
    for (int k = 0; k < table->n_shards; ++k) {
      MKE_CHECK_ARG(table->peer_var[k], "peer pointer %d is null", k);
      sh.var[k] = table->peer_var[k];
    }
  }
  table_export_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(
      sh, table->stride, table->dim, table->normalised, idx_or_null, n, out);
  MKE_CHECK_LAUNCH("table_export_kernel");
  return 0;
}

extern "C" int mke_fill_rows(float* buf, int32_t rows, int32_t stride, int32_t dim, float value,
                             mke_stream_t stream) {
  MKE_CHECK_ARG(buf && rows >= 0 && stride > 0 && dim >= 0 && dim <= stride, "bad fill arguments");
  const size_t total = (size_t)rows * stride;
  if (total == 0) return 0;
  size_t blocks = (total + 255) / 256;
  const size_t full = (size_t)sm_count() * 16;
  if (blocks > full) blocks = full;
  fill_rows_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(buf, total, stride, dim, value);
  MKE_CHECK_LAUNCH("fill_rows_kernel");
  return 0;
}

// ---- peer memory (row-sharded tables): plain cudaMalloc so that the block can be exported ------
extern "C" int mke_peer_alloc(uint64_t bytes, void** ptr) {
  MKE_CHECK_ARG(ptr && bytes > 0, "bad peer_alloc arguments");
  if (cudaError_t e = cudaMalloc(ptr, bytes)) return cuda_fail(e, "cudaMalloc");
  if (cudaError_t e = cudaMemset(*ptr, 0, bytes)) return cuda_fail(e, "cudaMemset");
  return 0;
}
extern "C" int mke_peer_free(void* ptr) {
  if (ptr == nullptr) return 0;
  if (cudaError_t e = cudaFree(ptr)) return cuda_fail(e, "cudaFree");
  return 0;
}
extern "C" int mke_ipc_export(const void* ptr, unsigned char handle[64]) {
  MKE_CHECK_ARG(ptr && handle, "null pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
  cudaIpcMemHandle_t h;
  if (cudaError_t e = cudaIpcGetMemHandle(&h, const_cast<void*>(ptr))) return cuda_fail(e, "cudaIpcGetMemHandle");
  memcpy(handle, &h, 64);
  return 0;
}
extern "C" int mke_ipc_open(const unsigned char handle[64], void** ptr) {
  MKE_CHECK_ARG(ptr && handle, "null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  if (cudaError_t e = cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess))
    return cuda_fail(e, "cudaIpcOpenMemHandle");
  return 0;
}
extern "C" int mke_ipc_close(void* ptr) {
  if (ptr == nullptr) return 0;
  if (cudaError_t e = cudaIpcCloseMemHandle(ptr)) return cuda_fail(e, "cudaIpcCloseMemHandle");
  return 0;
}
