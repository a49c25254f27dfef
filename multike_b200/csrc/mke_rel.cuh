// Parameters and small device helpers shared by the fused relation-view kernels.
#pragma once
#include "mke_common.cuh"

namespace mke {

// Row-sharded entity table as phase 1 sees it (mke_table_t.peer_*): row id lives in shard
// id & (G-1) at local row id >> log2(G).
struct EntShards {
  const float* var[MKE_MAX_SHARDS];
  float* grad[MKE_MAX_SHARDS];
  uint8_t* touched[MKE_MAX_SHARDS];
};

struct RelStepParams {
  const float* ent_var;
  float* ent_grad;
  uint8_t* ent_touched;
  const float* rel_var;
  float* rel_grad;
  int rel_rep;              // gradient replicas of the relation table (>= 1)
  size_t rel_rep_floats;    // floats per replica = rows * stride
  uint8_t* rel_touched;
  int stride;    // floats per row (both tables)
  int nchunk;    // float4 pieces per row that carry data = ceil(dim/4)
  int ent_norm;  // read l2_normalize(ent_var,1)
  int rel_norm;
  const int32_t* pos1;
  int len1;
  const int32_t* pos2;
  int len2;
  mke_kg_sampler_t kg1, kg2;
  int K;
  int sampled;  // 1: draw negatives on device, 0: read neg_ent / neg_side
  uint64_t skey;
  const int32_t* neg_ent;
  const uint32_t* neg_side;
  // "negatives where they live" (row-sharded tables, sharded.py): bit j of neg_valid[i] clear =>
  // negative j of positive i belongs to another rank (its slot holds a local dummy row and
  // contributes nothing); the positive term of positive i is this launch's only if
  // pos_own_lo <= i < pos_own_hi.  NULL / [0, INT_MAX) => everything is this launch's.
  const uint32_t* neg_valid;
  int neg_compact;  // neg_valid words are low_ones(count): this launch's negatives come first (mke_neg_keep_owned2)
  int pos_own_lo, pos_own_hi;
  const float* w;
  float pos_scale;
  double* loss;
  int32_t* neg_out;
  EntShards sh;             // used when sharded != 0
  ShardMap smap;            // placement of an entity id (mke_table_t.n_shards / shard_split)
  int sharded;              // 0 = one local table
  int index_base;           // position of this launch's first positive inside its global batch
  unsigned long long* trace;  // debug: per-warp milestone clocks (MKE_TRACE), else NULL
  int dbg;  // timing experiments only (MKE_DEBUG_SKIP): bit0 no rel RED, bit1 no touched, bit2 no loss atomic, bit3 no ent RED, bit4 no hash probe
};

// log(1+exp(x)) and sigmoid(x) as the reference writes them (losses.py:9-10: naive
// tf.log(1 + tf.exp(x)); no softplus stabilisation -- x = +-||.||^2, |x| <= 9 for unit rows).
// MUFU.EX2 / MUFU.LG2 / MUFU.RCP forms: relative error of exp <= 2^-21 on |x| <= 16, absolute
// error of log <= 2^-21 -- far inside the 1e-5 loss / 1e-4 gradient tolerances.
__device__ __forceinline__ void softplus_sigmoid(float x, float& sp, float& sg) {
  const float ex = __expf(x);
  const float one_p = 1.0f + ex;
  sp = __logf(one_p);
  sg = __fdividef(ex, one_p);
}

// entity rows: local table, or the owner's shard through its peer mapping
__device__ __forceinline__ const float* ent_var_row(const RelStepParams& p, int32_t id, int stride) {
  if (!p.sharded) return p.ent_var + (size_t)id * stride;
  int s;
  int32_t l;
  p.smap.locate(id, s, l);
  return p.sh.var[s] + (size_t)l * stride;
}
__device__ __forceinline__ float* ent_grad_row(const RelStepParams& p, int32_t id, int stride) {
  if (!p.sharded) return p.ent_grad + (size_t)id * stride;
  int s;
  int32_t l;
  p.smap.locate(id, s, l);
  return p.sh.grad[s] + (size_t)l * stride;
}
__device__ __forceinline__ void ent_mark(const RelStepParams& p, int32_t id) {
  if (!p.sharded) {
    mark_touched(p.ent_touched, id);
  } else {
    int s;
    int32_t l;
    p.smap.locate(id, s, l);
    mark_touched(p.sh.touched[s], l);
  }
}

// this thread block's copy of the relation gradient table (mke_table_t.grad_replicas)
__device__ __forceinline__ float* rel_grad_replica(const RelStepParams& p) {
  return p.rel_grad + (size_t)(blockIdx.x % (unsigned)p.rel_rep) * p.rel_rep_floats;
}

// quarter-warp kernel (mke_rel_q8.cu); returns 1 when the stride has no instantiation
int launch_rel_q8(const RelStepParams& p, cudaStream_t stream);
// persistent row-stream schedule of the same kernel (mke_rel_q8p.cu); returns 1 when the launch shape is not covered
int launch_rel_q8p(const RelStepParams& p, int cfg, cudaStream_t stream);

}  // namespace mke
