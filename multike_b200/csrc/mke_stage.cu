// Row staging for row-sharded tables (BASELINE configs[3]: full multi-view on the sharded
// entity tables).  The small-batch graphs of the other views (attribute CNN MultiKE_model.py:134-151, cross-KG
// inference :158-221, ITC :225-239, SSL mapping :241-261) run on a STAGED copy of the rows their batch touches:
//   mke_table_stage_rows     raw rows of GLOBAL ids -> rows 0 .. n-1 of a plain local table (peer reads through the
//                            owner's mapping), its gradient rows zeroed: the unchanged single-GPU kernels then run on
//                            the staged table with indices 0 .. n-1;
//   mke_table_commit_grads   gradient rows of the staged table -> added to the gradient rows of the ids THIS rank owns
//                            (+ touched flags), after which the owner's phase 2 (mke_rows_apply_adagrad) updates them;
//   mke_peer_barrier         the flag barrier of mke_sharded.cu as a launch of its own (nobody updates rows others still
//                            stage; nobody stages rows others still update).
// Every rank stages the whole batch and computes the same gradients (these graphs couple the batch through global
// l2-norms and take ~100 us at B = 5 000), so nothing but row reads crosses NVLink and the dense parameters
// (CNN weights, attr_embeds, mappings) stay replicated without a collective.
#include "mke_common.cuh"

namespace mke {

struct StageShards {
  const float* var[MKE_MAX_SHARDS];
  ShardMap map;
  int sharded;
};

// one warp per staged row
__global__ void stage_rows_kernel(StageShards sh, int stride, const int32_t* __restrict__ ids, int n,
                                  float* __restrict__ dst_var, float* __restrict__ dst_grad, int dst_stride) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int i = blockIdx.x * wpb + (threadIdx.x >> 5); i < n; i += gridDim.x * wpb) {
    const int row = __ldg(ids + i);
    int shard = 0;
    int32_t local = row;
    if (sh.sharded) sh.map.locate(row, shard, local);
    const float* pv = sh.var[shard] + (size_t)local * stride;
    for (int c = lane; c < dst_stride; c += 32) {
      dst_var[(size_t)i * dst_stride + c] = c < stride ? pv[c] : 0.f;
      if (dst_grad != nullptr) dst_grad[(size_t)i * dst_stride + c] = 0.f;
    }
  }
}

__global__ void commit_grads_kernel(ShardMap map, int sharded, int my_shard, int stride, const int32_t* __restrict__ ids,
                                    int n, const float* __restrict__ src_grad, int src_stride, float* __restrict__ grad,
                                    uint8_t* __restrict__ touched) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int i = blockIdx.x * wpb + (threadIdx.x >> 5); i < n; i += gridDim.x * wpb) {
    const int row = __ldg(ids + i);
    int shard = 0;
    int32_t local = row;
    if (sharded) map.locate(row, shard, local);
    if (sharded && shard != my_shard) continue;
    float* g = grad + (size_t)local * stride;
    const float* s = src_grad + (size_t)i * src_stride;
    for (int c = lane; c < stride && c < src_stride; c += 32) {
      const float v = s[c];
      if (v != 0.f) atomicAdd(g + c, v);  // the same id may be staged more than once
    }
    if (touched != nullptr && lane == 0) touched[local] = 1;
  }
}

struct BarrierPeers {
  uint32_t* flags[MKE_MAX_SHARDS];
  int world, rank;
};
constexpr unsigned long long kStageWaitNs = 20000000000ull;

__global__ void stage_barrier_kernel(const BarrierPeers ps, uint32_t seq) {
  if ((int)threadIdx.x < ps.world) {
    const int k = threadIdx.x;
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(ps.flags[k] + ps.rank), "r"(seq) : "memory");
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0)::"memory");
    while (true) {
      uint32_t v;
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(ps.flags[ps.rank] + k) : "memory");
      if ((int32_t)(v - seq) >= 0) break;
      __nanosleep(200);
      unsigned long long t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1)::"memory");
      if (t1 - t0 > kStageWaitNs) __trap();  // a rank that never arrives is a bug: fail, do not hang
    }
  }
}

}  // namespace mke

using namespace mke;

extern "C" int mke_table_stage_rows(const mke_table_t* table, const int32_t* ids, int32_t n, const mke_table_t* staged,
                                    mke_stream_t stream) {
  MKE_CHECK_ARG(table && table->var && staged && staged->var && ids, "null pointer");
  MKE_CHECK_ARG(n >= 0 && n <= staged->rows, "n=%d rows do not fit the staged table (%d rows)", n, staged->rows);
  MKE_CHECK_ARG(staged->n_shards <= 1 && staged->stride >= table->stride && staged->dim == table->dim,
                "the staged table is a plain local table of the same dim and at least the same stride");
  if (n == 0) return 0;
  StageShards sh{};
  sh.var[0] = table->var;
  if (table->n_shards > 1) {
    sh.sharded = 1;
    sh.map = shard_map(table);
    for (int k = 0; k < table->n_shards; ++k) {
      MKE_CHECK_ARG(table->peer_var[k], "peer pointer %d is null", k);
      sh.var[k] = table->peer_var[k];
    }
  }
  int blocks = (n + 7) / 8;
  const int full = sm_count() * 8;
  if (blocks > full) blocks = full;
  stage_rows_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(sh, table->stride, ids, n, staged->var, staged->grad,
                                                              staged->stride);
  MKE_CHECK_LAUNCH("stage_rows_kernel");
  return 0;
}

extern "C" int mke_table_commit_grads(const mke_table_t* table, const int32_t* ids, int32_t n, const mke_table_t* staged,
                                      mke_stream_t stream) {
  MKE_CHECK_ARG(table && table->grad && staged && staged->grad && ids, "null pointer");
  MKE_CHECK_ARG(n >= 0 && n <= staged->rows, "n=%d rows exceed the staged table (%d rows)", n, staged->rows);
  MKE_CHECK_ARG(table->grad_replicas <= 1 && staged->grad_replicas <= 1, "tables with one gradient copy");
  if (n == 0) return 0;
  int blocks = (n + 7) / 8;
  const int full = sm_count() * 8;
  if (blocks > full) blocks = full;
  const int sharded = table->n_shards > 1 ? 1 : 0;
  commit_grads_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(shard_map(table), sharded, table->shard_rank, table->stride,
                                                                ids, n, staged->grad, staged->stride, table->grad,
                                                                table->touched);
  MKE_CHECK_LAUNCH("commit_grads_kernel");
  return 0;
}

extern "C" int mke_peer_barrier(void* const* flags, int32_t world, int32_t rank, uint32_t seq, mke_stream_t stream) {
  MKE_CHECK_ARG(flags && world >= 1 && world <= MKE_MAX_SHARDS && rank >= 0 && rank < world, "bad barrier arguments");
  BarrierPeers ps{};
  ps.world = world;
  ps.rank = rank;
  for (int k = 0; k < world; ++k) {
    MKE_CHECK_ARG(flags[k], "flag array %d is null", k);
    ps.flags[k] = (uint32_t*)flags[k];
  }
  stage_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(ps, seq);
  MKE_CHECK_LAUNCH("stage_barrier_kernel");
  return 0;
}
