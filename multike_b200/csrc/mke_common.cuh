// Shared device/host helpers for the MultiKE B200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/multike_b200.h"

#ifndef __CUDA_ARCH__
#define MKE_HOST_ONLY 1
#endif

namespace mke {

constexpr int kWarp = 32;
constexpr float kNormEps = 1e-12f;  // tf.nn.l2_normalize epsilon [TF semantics]
constexpr uint64_t kGamma = 0x9E3779B97F4A7C15ull;
constexpr uint64_t kEmptySlot = 0xFFFFFFFFFFFFFFFFull;
constexpr uint32_t kSideDraw = 0xFFFFu;  // draw index reserved for the head/tail coin flip

// ---- error plumbing (host) -----------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
void count_launch();
int sm_count();

#define MKE_CHECK_ARG(cond, ...)                                                              \
  do {                                                                                        \
    if (!(cond)) {                                                                            \
      ::mke::set_error(__VA_ARGS__);                                                          \
      return MKE_EINVAL;                                                                      \
    }                                                                                         \
  } while (0)

#define MKE_CHECK_LAUNCH(what)                                                                \
  do {                                                                                        \
    ::mke::count_launch();                                                                    \
    cudaError_t e__ = cudaGetLastError();                                                     \
    if (e__ != cudaSuccess) return ::mke::cuda_fail(e__, what);                               \
  } while (0)

// rows of a (possibly row-sharded) table that live on this rank: ids r with r % G == rank
inline int table_local_rows(const mke_table_t* t) {
  if (t->n_shards <= 1) return t->rows;
  auto part = [](int rows, int r, int g) { return rows > r ? (rows - r + g - 1) / g : 0; };
  if (t->shard_split <= 0) return part(t->rows, t->shard_rank, t->n_shards);
  const int half = t->n_shards / 2;  // KG-block placement
  return t->shard_rank < half ? part(t->shard_split, t->shard_rank, half)
                              : part(t->rows - t->shard_split, t->shard_rank - half, half);
}
// Placement of a row id (device + host): shard index and local row.  log2g = log2 of the number
// of ranks an id is spread over (G, or G/2 per KG with split > 0).
struct ShardMap {
  int log2g;
  int split;  // 0 => plain id % G
  __host__ __device__ __forceinline__ void locate(int32_t id, int& shard, int32_t& local) const {
    const int m = (1 << log2g) - 1;
    if (split <= 0) {
      shard = id & m;
      local = id >> log2g;
    } else {
      const bool second = id >= split;
      const int32_t x = second ? id - split : id;
      shard = (second ? (1 << log2g) : 0) | (x & m);
      local = x >> log2g;
    }
  }
};
inline ShardMap shard_map(const mke_table_t* t) {
  ShardMap m;
  m.split = t->shard_split > 0 ? t->shard_split : 0;
  const int spread = m.split > 0 ? t->n_shards / 2 : t->n_shards;
  m.log2g = spread >= 8 ? 3 : spread >= 4 ? 2 : spread >= 2 ? 1 : 0;
  return m;
}
inline int shard_log2(int n_shards) {
  return n_shards >= 8 ? 3 : n_shards >= 4 ? 2 : n_shards >= 2 ? 1 : 0;
}

// ---- counter-based RNG (restated bit-exactly in oracle/sampler.py) --------------------------
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x ^= x >> 30;
  x *= 0xBF58476D1CE4E5B9ull;
  x ^= x >> 27;
  x *= 0x94D049BB133111EBull;
  x ^= x >> 31;
  return x;
}
__host__ __device__ __forceinline__ uint64_t stream_key(uint64_t seed, uint64_t step) {
  return mix64(seed + kGamma * (step + 1));
}
// one 64-bit draw at coordinates (positive index in batch, try, draw index)
__host__ __device__ __forceinline__ uint64_t draw64(uint64_t skey, uint32_t i, uint32_t tr,
                                                    uint32_t c) {
  uint64_t coord = ((uint64_t)i << 20) | ((uint64_t)tr << 16) | (uint64_t)c;
  return mix64(skey + (coord + 1) * kGamma);
}
// unbiased-enough index in [0, n): high 32 bits scaled by n (multiply-high)
__host__ __device__ __forceinline__ uint32_t draw_index(uint64_t r, uint32_t n) {
  return (uint32_t)(((r >> 32) * (uint64_t)n) >> 32);
}
__host__ __device__ __forceinline__ uint64_t triple_key(int32_t h, int32_t r, int32_t t) {
  return ((uint64_t)(uint32_t)h << 40) | ((uint64_t)(uint32_t)r << 24) | (uint64_t)(uint32_t)t;
}

#ifdef __CUDACC__
// ---- device helpers ------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ void warp_sum2(float& a, float& b) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
}
__device__ __forceinline__ void warp_sum3(float& a, float& b, float& c) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
    c += __shfl_xor_sync(0xffffffffu, c, o);
  }
}
__device__ __forceinline__ float dot4(const float4& a, const float4& b) {
  return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
}
__device__ __forceinline__ float4 f4_add(const float4& a, const float4& b) {
  return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}
__device__ __forceinline__ float4 f4_sub(const float4& a, const float4& b) {
  return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w);
}
__device__ __forceinline__ float4 f4_scale(const float4& a, float s) {
  return make_float4(a.x * s, a.y * s, a.z * s, a.w * s);
}
__device__ __forceinline__ float4 f4_fma(const float4& a, float s, const float4& b) {  // a*s+b
  return make_float4(fmaf(a.x, s, b.x), fmaf(a.y, s, b.y), fmaf(a.z, s, b.z), fmaf(a.w, s, b.w));
}
__device__ __forceinline__ float4 f4_zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }

// 16-byte read-only gather of one quarter-sector-aligned piece of a table row
__device__ __forceinline__ float4 ldg_f4(const float* p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}
// vectorised fire-and-forget reduction: SASS REDG.E.ADD.F32x4 (sm_90+)
__device__ __forceinline__ void red_add_f4(float* p, const float4& v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}

// Row flag for phase 2: a plain fire-and-forget byte store.  (Test-before-set was measured and is
// worse: the flag load joins the warp's dependent chain; hot small tables carry no flags at all.)
__device__ __forceinline__ void mark_touched(uint8_t* __restrict__ flags, int32_t row) {
  if (flags != nullptr) flags[row] = 1;
}

// Triple set: open addressing over BUCKETS of four 64-bit keys (one 32-byte sector); a bucket
// fills front to back and never empties, so a lookup reads one sector and stops at the first
// bucket whose last slot is still empty.  A miss costs one memory round trip at load <= 0.5.
constexpr int kBucketSlots = 4;
// continue a lookup at bucket b (buckets before b were full and did not hold the key)
__device__ __forceinline__ bool tripleset_probe_from(const mke_tripleset_t& s, uint64_t key, uint64_t b) {
  const uint64_t mask = (s.capacity / kBucketSlots) - 1;
  while (true) {
    b &= mask;
    const ulonglong2* p = reinterpret_cast<const ulonglong2*>(s.slots + b * kBucketSlots);
    const ulonglong2 lo = __ldg(p), hi = __ldg(p + 1);
    if (lo.x == key || lo.y == key || hi.x == key || hi.y == key) return true;
    if (hi.y == kEmptySlot) return false;
    ++b;
  }
}
__device__ __forceinline__ bool tripleset_contains(const mke_tripleset_t& s, uint64_t key) {
  if (s.slots == nullptr) return false;
  return tripleset_probe_from(s, key, mix64(key));
}
// two independent lookups with both sectors in flight at once
__device__ __forceinline__ void tripleset_contains2(const mke_tripleset_t& s, uint64_t k0, uint64_t k1,
                                                    bool& in0, bool& in1) {
  in0 = in1 = false;
  if (s.slots == nullptr) return;
  const uint64_t mask = (s.capacity / kBucketSlots) - 1;
  const uint64_t b0 = mix64(k0) & mask, b1 = mix64(k1) & mask;
  const ulonglong2* p0 = reinterpret_cast<const ulonglong2*>(s.slots + b0 * kBucketSlots);
  const ulonglong2* p1 = reinterpret_cast<const ulonglong2*>(s.slots + b1 * kBucketSlots);
  const ulonglong2 a0 = __ldg(p0), a1 = __ldg(p0 + 1), c0 = __ldg(p1), c1 = __ldg(p1 + 1);
  in0 = (a0.x == k0 || a0.y == k0 || a1.x == k0 || a1.y == k0);
  in1 = (c0.x == k1 || c0.y == k1 || c1.x == k1 || c1.y == k1);
  if (!in0 && a1.y != kEmptySlot) in0 = tripleset_probe_from(s, k0, b0 + 1);  // full bucket: rare
  if (!in1 && c1.y != kEmptySlot) in1 = tripleset_probe_from(s, k1, b1 + 1);
}

// candidate pool for replacing `anchor` (base/batch.py:93-94 neighbor.get(e, entities_list))
struct CandPool {
  const int32_t* list;  // nullptr => base + idx
  int32_t base;
  uint32_t n;
  __device__ __forceinline__ int32_t at(uint32_t idx) const {
    return list ? __ldg(list + idx) : base + (int32_t)idx;
  }
};
__device__ __forceinline__ CandPool cand_pool(const mke_kg_sampler_t& kg, int32_t anchor) {
  CandPool c;
  if (kg.neighbours != nullptr) {
    const int32_t* row = kg.neighbours + (size_t)anchor * (size_t)kg.n_neighbours;
    if (__ldg(row) >= 0) {
      c.list = row;
      c.base = 0;
      c.n = (uint32_t)kg.n_neighbours;
      return c;
    }
  }
  c.list = kg.entity_list;
  c.base = kg.entity_base;
  c.n = (uint32_t)kg.n_entities;
  return c;
}

// Sequential statement of generate_neg_triples_fast (base/batch.py:86-116) for ONE positive,
// run redundantly by all 32 lanes (uniform control flow); s_pick is a 32-int scratch in shared
// memory private to the warp.  Returns the corrupted entity for lane j (< K) and the side mask
// (bit j set => negative j replaces the head).
static __device__ __noinline__ void sample_negs_sequential(const mke_kg_sampler_t& kg, int32_t h,
                                                    int32_t r, int32_t t, int K, uint64_t skey,
                                                    uint32_t i, int lane, volatile int32_t* s_pick,
                                                    int32_t& e_out, uint32_t& side_out) {
  int n_acc = 0;
  uint32_t side_mask = 0;
  int remaining = K;
  for (uint32_t tr = 0; tr < MKE_MAX_TRY; ++tr) {
    const bool head_side = (draw64(skey, i, tr, kSideDraw) >> 63) != 0;
    const CandPool pool = cand_pool(kg, head_side ? h : t);
    int np = 0;
    uint32_t c = 0;
    while (np < remaining) {
      const int32_t e = pool.at(draw_index(draw64(skey, i, tr, c), pool.n));
      ++c;
      bool dup = false;
      for (int q = 0; q < np; ++q) dup |= (s_pick[n_acc + q] == e);
      if (dup && c < kSideDraw) continue;  // random.sample draws without replacement
      __syncwarp();
      if (lane == 0) s_pick[n_acc + np] = e;
      __syncwarp();
      ++np;
    }
    int kept = np;
    if (tr != MKE_MAX_TRY - 1) {  // the last try is accepted unfiltered (batch.py:103-105)
      kept = 0;
      for (int q = 0; q < np; ++q) {
        const int32_t e = s_pick[n_acc + q];
        const uint64_t key = head_side ? triple_key(e, r, t) : triple_key(h, r, e);
        if (!tripleset_contains(kg.set, key)) {
          __syncwarp();
          if (lane == 0) s_pick[n_acc + kept] = e;
          __syncwarp();
          ++kept;
        }
      }
    }
    if (head_side && kept > 0) {
      const uint32_t ones = (kept >= 32) ? 0xffffffffu : ((1u << kept) - 1u);
      side_mask |= ones << n_acc;
    }
    n_acc += kept;
    if (n_acc >= K) break;
    remaining = K - n_acc;
  }
  __syncwarp();
  e_out = (lane < K) ? s_pick[lane] : -1;
  side_out = side_mask;
  __syncwarp();
}

// Warp-parallel front end: lane j draws negative j of try 0; falls back to the sequential
// statement when a duplicate or a filtered candidate shows up (about 0.3 % of positives).
__device__ __forceinline__ void sample_negs_warp(const mke_kg_sampler_t& kg, int32_t h, int32_t r,
                                                 int32_t t, int K, uint64_t skey, uint32_t i,
                                                 int lane, volatile int32_t* s_pick,
                                                 int32_t& e_out, uint32_t& side_out) {
  const bool head_side = (draw64(skey, i, 0, kSideDraw) >> 63) != 0;
  const CandPool pool = cand_pool(kg, head_side ? h : t);
  int32_t e = -1 - lane;  // distinct dummies for idle lanes
  bool bad = false;
  if (lane < K) {
    e = pool.at(draw_index(draw64(skey, i, 0, (uint32_t)lane), pool.n));
    const uint64_t key = head_side ? triple_key(e, r, t) : triple_key(h, r, e);
    bad = tripleset_contains(kg.set, key);
  }
  const uint32_t peers = __match_any_sync(0xffffffffu, e);
  bad |= (peers & (peers - 1)) != 0;  // another lane drew the same entity
  if (__any_sync(0xffffffffu, bad)) {
    sample_negs_sequential(kg, h, r, t, K, skey, i, lane, s_pick, e_out, side_out);
    return;
  }
  e_out = (lane < K) ? e : -1;
  side_out = head_side ? ((K >= 32) ? 0xffffffffu : ((1u << K) - 1u)) : 0u;
}
#endif  // __CUDACC__

}  // namespace mke
