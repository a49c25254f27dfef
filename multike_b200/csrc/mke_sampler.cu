// Device-resident triple set + stand-alone negative sampler.
// Replaces the Python set membership test and generate_neg_triples_fast of
// base/batch.py:86-116 (all_triples_set = kg.local_relation_triples_set, which aliases
// relation_triples_set and therefore also holds the swapped sup triples: base/kg.py:59,134).
#include "mke_sampler.cuh"

namespace mke {

__global__ void tripleset_build_kernel(mke_tripleset_t set, const int32_t* __restrict__ triples,
                                       int n) {
  const uint64_t mask = (set.capacity / kBucketSlots) - 1;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint64_t key = triple_key(triples[3 * i], triples[3 * i + 1], triples[3 * i + 2]);
    uint64_t b = mix64(key) & mask;
    bool done = false;
    while (!done) {
      unsigned long long* slot = reinterpret_cast<unsigned long long*>(set.slots + b * kBucketSlots);
      for (int q = 0; q < kBucketSlots && !done; ++q) {  // front to back: buckets fill in order
        const unsigned long long prev =
            atomicCAS(slot + q, (unsigned long long)kEmptySlot, (unsigned long long)key);
        done = (prev == kEmptySlot || prev == key);
      }
      b = (b + 1) & mask;
    }
  }
}

__global__ void tripleset_contains_kernel(mke_tripleset_t set, const int32_t* __restrict__ triples,
                                          int n, uint8_t* __restrict__ out) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint64_t key = triple_key(triples[3 * i], triples[3 * i + 1], triples[3 * i + 2]);
    out[i] = tripleset_contains(set, key) ? 1 : 0;
  }
}

constexpr int kSampThreads = 128;
constexpr int kSampWarps = kSampThreads / 32;

// One positive per quarter warp, the same sampler (and therefore the same draws) as the fused
// relation kernel.  Output either as (h,r,t) rows, positive-major (what base/batch.py:116
// returns), or in the structured form mke_rel_step_structured consumes.
__global__ void __launch_bounds__(kSampThreads)
    sample_kernel(const int32_t* __restrict__ pos1, int len1, mke_kg_sampler_t kg1,
                  const int32_t* __restrict__ pos2, int len2, mke_kg_sampler_t kg2, int K,
                  uint64_t skey, int index_base, int32_t* __restrict__ neg_out, int32_t* __restrict__ neg_ent,
                  uint32_t* __restrict__ neg_side) {
  __shared__ int32_t s_pick_all[kSampWarps][kQPerWarp][kPickStride];
  const int lane = threadIdx.x & 31;
  const int sub = lane & 7;
  const int q = lane >> 3;
  const int wib = threadIdx.x >> 5;
  volatile int32_t* pick = s_pick_all[wib][q];
  const int n = len1 + len2;
  const int per_pass = gridDim.x * kSampWarps * kQPerWarp;
  for (int i0 = (blockIdx.x * kSampWarps + wib) * kQPerWarp; i0 < n; i0 += per_pass) {
    const int i = i0 + q;
    if (i < n) {  // quarters are independent: the sampler only synchronises inside a quarter
      const bool first = i < len1;
      const int32_t* row = first ? pos1 + 3 * (size_t)i : pos2 + 3 * (size_t)(i - len1);
      const int32_t h = __ldg(row), r = __ldg(row + 1), t = __ldg(row + 2);
      const KgView kg = kg_view(kg1, kg2, first);
      const uint32_t side = sample_negs_quarter(kg, h, r, t, K, skey, (uint32_t)(index_base + i), lane, pick);
      for (int c = sub; c < K; c += 8) {
        const int32_t e = pick[c];
        if (neg_ent != nullptr) neg_ent[(size_t)i * K + c] = e;
        if (neg_out != nullptr) {
          const bool hs = (side >> c) & 1u;
          int32_t* o = neg_out + ((size_t)i * K + c) * 3;
          o[0] = hs ? e : h;
          o[1] = r;
          o[2] = hs ? t : e;
        }
      }
      if (neg_side != nullptr && sub == 0) neg_side[i] = side;
      __syncwarp(0xffu << (lane & 24));  // pick[] is rewritten by this quarter's next positive
    }
  }
}

static int launch_sample(const int32_t* pos1, int32_t len1, const mke_kg_sampler_t* kg1,
                         const int32_t* pos2, int32_t len2, const mke_kg_sampler_t* kg2, int32_t K,
                         uint64_t seed, uint64_t step, int32_t index_base, int32_t* neg_out, int32_t* neg_ent,
                         uint32_t* neg_side, cudaStream_t stream) {
  MKE_CHECK_ARG(index_base >= 0 && (long long)index_base + len1 + len2 < (1ll << 31), "bad index_base");
  MKE_CHECK_ARG(K >= 1 && K <= MKE_MAX_NEG, "K=%d outside [1,%d]", K, MKE_MAX_NEG);
  MKE_CHECK_ARG(len1 >= 0 && len2 >= 0, "negative batch length");
  MKE_CHECK_ARG(len1 == 0 || (pos1 && kg1), "kg1 slice needs positives and a sampler");
  MKE_CHECK_ARG(len2 == 0 || (pos2 && kg2), "kg2 slice needs positives and a sampler");
  const int n = len1 + len2;
  if (n == 0) return 0;
  mke_kg_sampler_t a{}, b{};
  if (kg1) a = *kg1;
  if (kg2) b = *kg2;
  for (const mke_kg_sampler_t* kg : {len1 ? kg1 : nullptr, len2 ? kg2 : nullptr}) {
    if (!kg) continue;
    MKE_CHECK_ARG(kg->n_entities >= K, "KG has fewer entities (%d) than K=%d", kg->n_entities, K);
    MKE_CHECK_ARG(!kg->neighbours || kg->n_neighbours >= K, "n_neighbours < K");
    MKE_CHECK_ARG(!kg->set.slots || (kg->set.capacity >= 8 && (kg->set.capacity & (kg->set.capacity - 1)) == 0),
                  "triple-set capacity must be a power of two >= 8");
  }
  constexpr int per_block = kSampWarps * kQPerWarp;
  int blocks = (n + per_block - 1) / per_block;
  const int full = sm_count() * 16;
  if (blocks > full) blocks = full;
  sample_kernel<<<blocks, kSampThreads, 0, stream>>>(pos1, len1, a, pos2, len2, b, K, stream_key(seed, step),
                                                     index_base, neg_out, neg_ent, neg_side);
  MKE_CHECK_LAUNCH("sample_kernel");
  return 0;
}

// attr_batch.py:13-25 generate_neg_attribute_triples: head-only corruption, K independent draws per
// positive (WITH replacement across the K), each redrawn until (h', a, v) is not a known attribute
// triple.  The reference retries without bound; here try kAttrMaxTry - 1 is accepted unfiltered.
constexpr uint32_t kAttrMaxTry = 64;
__global__ void attr_sample_kernel(const int32_t* __restrict__ pos1, int len1, mke_kg_sampler_t kg1,
                                   const int32_t* __restrict__ pos2, int len2, mke_kg_sampler_t kg2, int K,
                                   uint64_t skey, int index_base, int32_t* __restrict__ neg_head) {
  const long long total = (long long)(len1 + len2) * K;
  for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < total;
       g += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(g / K), j = (int)(g % K);
    const bool first = i < len1;
    const int32_t* row = first ? pos1 + 3 * (size_t)i : pos2 + 3 * (size_t)(i - len1);
    const int32_t h = __ldg(row), a = __ldg(row + 1), v = __ldg(row + 2);
    const KgView kg = kg_view(kg1, kg2, first);
    const CandPool pool = kg.pool(h);
    int32_t e = h;
    for (uint32_t tr = 0; tr < kAttrMaxTry; ++tr) {
      // coordinates: (positive, try, draw) -- try fits the 4-bit field only up to 15, so the try
      // index is folded into the draw field instead: draw = j + 32 * try  (j < 32, try < 64)
      e = pool.at(draw_index(draw64(skey, (uint32_t)(index_base + i), 0u, (uint32_t)j + 32u * tr), pool.n));
      if (tr == kAttrMaxTry - 1 || !tripleset_contains(kg.set, triple_key(e, a, v))) break;
    }
    neg_head[g] = e;
  }
}

static int check_set(const mke_tripleset_t* set) {
  MKE_CHECK_ARG(set && set->slots, "null triple set");
  MKE_CHECK_ARG(set->capacity >= 8 && (set->capacity & (set->capacity - 1)) == 0,
                "triple-set capacity must be a power of two >= 8");
  return 0;
}

}  // namespace mke

using namespace mke;

extern "C" int mke_tripleset_build(const mke_tripleset_t* set, const int32_t* triples, int32_t n,
                                   mke_stream_t stream) {
  if (int rc = check_set(set)) return rc;
  MKE_CHECK_ARG(n >= 0 && (uint64_t)n * 2 <= set->capacity, "capacity %llu < 2*n (n=%d)",
                (unsigned long long)set->capacity, n);
  if (n == 0) return 0;
  MKE_CHECK_ARG(triples, "null triples");
  const int blocks = (n + 255) / 256;
  tripleset_build_kernel<<<blocks < 4096 ? blocks : 4096, 256, 0, (cudaStream_t)stream>>>(*set, triples, n);
  MKE_CHECK_LAUNCH("tripleset_build_kernel");
  return 0;
}

extern "C" int mke_tripleset_contains(const mke_tripleset_t* set, const int32_t* triples, int32_t n,
                                      uint8_t* out, mke_stream_t stream) {
  if (int rc = check_set(set)) return rc;
  MKE_CHECK_ARG(n >= 0, "negative n");
  if (n == 0) return 0;
  MKE_CHECK_ARG(triples && out, "null pointer");
  const int blocks = (n + 255) / 256;
  tripleset_contains_kernel<<<blocks < 4096 ? blocks : 4096, 256, 0, (cudaStream_t)stream>>>(*set, triples, n, out);
  MKE_CHECK_LAUNCH("tripleset_contains_kernel");
  return 0;
}

extern "C" int mke_sample_uniform(const int32_t* pos1, int32_t len1, const mke_kg_sampler_t* kg1,
                                  const int32_t* pos2, int32_t len2, const mke_kg_sampler_t* kg2,
                                  int32_t K, uint64_t seed, uint64_t step, int32_t* neg_out,
                                  mke_stream_t stream) {
  MKE_CHECK_ARG(neg_out || len1 + len2 == 0, "neg_out is null");
  return launch_sample(pos1, len1, kg1, pos2, len2, kg2, K, seed, step, 0, neg_out, nullptr, nullptr,
                       (cudaStream_t)stream);
}

extern "C" int mke_sample_structured(const int32_t* pos1, int32_t len1, const mke_kg_sampler_t* kg1,
                                     const int32_t* pos2, int32_t len2, const mke_kg_sampler_t* kg2,
                                     int32_t K, uint64_t seed, uint64_t step, int32_t* neg_ent,
                                     uint32_t* neg_side, mke_stream_t stream) {
  return mke_sample_structured_at(pos1, len1, kg1, pos2, len2, kg2, K, seed, step, 0, neg_ent, neg_side, stream);
}

extern "C" int mke_sample_structured_at(const int32_t* pos1, int32_t len1, const mke_kg_sampler_t* kg1,
                                        const int32_t* pos2, int32_t len2, const mke_kg_sampler_t* kg2,
                                        int32_t K, uint64_t seed, uint64_t step, int32_t index_base,
                                        int32_t* neg_ent, uint32_t* neg_side, mke_stream_t stream) {
  MKE_CHECK_ARG((neg_ent && neg_side) || len1 + len2 == 0, "neg_ent/neg_side are null");
  return launch_sample(pos1, len1, kg1, pos2, len2, kg2, K, seed, step, index_base, nullptr, neg_ent, neg_side,
                       (cudaStream_t)stream);
}

extern "C" int mke_sample_attribute_heads(const int32_t* pos1, int32_t len1, const mke_kg_sampler_t* kg1,
                                          const int32_t* pos2, int32_t len2, const mke_kg_sampler_t* kg2,
                                          int32_t K, uint64_t seed, uint64_t step, int32_t index_base,
                                          int32_t* neg_head, mke_stream_t stream) {
  MKE_CHECK_ARG(K >= 1 && K <= MKE_MAX_NEG, "K=%d outside [1,%d]", K, MKE_MAX_NEG);
  MKE_CHECK_ARG(len1 >= 0 && len2 >= 0 && index_base >= 0, "bad batch length / index_base");
  MKE_CHECK_ARG(len1 == 0 || (pos1 && kg1), "kg1 slice needs positives and a sampler");
  MKE_CHECK_ARG(len2 == 0 || (pos2 && kg2), "kg2 slice needs positives and a sampler");
  const long long total = (long long)(len1 + len2) * K;
  if (total == 0) return 0;
  MKE_CHECK_ARG(neg_head, "neg_head is null");
  mke_kg_sampler_t a{}, b{};
  if (kg1) a = *kg1;
  if (kg2) b = *kg2;
  for (const mke_kg_sampler_t* kg : {len1 ? kg1 : nullptr, len2 ? kg2 : nullptr}) {
    if (!kg) continue;
    MKE_CHECK_ARG(kg->n_entities >= 1, "empty candidate pool");
    MKE_CHECK_ARG(!kg->set.slots || (kg->set.capacity >= 8 && (kg->set.capacity & (kg->set.capacity - 1)) == 0),
                  "triple-set capacity must be a power of two >= 8");
  }
  long long blocks = (total + 255) / 256;
  const long long full = (long long)sm_count() * 16;
  if (blocks > full) blocks = full;
  attr_sample_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(pos1, len1, a, pos2, len2, b, K,
                                                                         stream_key(seed, step), index_base,
                                                                         neg_head);
  MKE_CHECK_LAUNCH("attr_sample_kernel");
  return 0;
}

// ---- "negatives where they live" (sharded.py) ---------------------------------------------------
// One thread per positive: negatives whose row lives on another shard are replaced by `dummy_id`
// (a row of this shard: its gather stays local and is served by L1) and their bit in
// neg_valid[i] is cleared, so the phase-1 kernels skip their contribution.
namespace mke {
__global__ void neg_keep_owned_kernel(int32_t* __restrict__ neg_ent, int n, int K, ShardMap smap, int my_shard,
                                      int32_t dummy_id, uint32_t* __restrict__ neg_valid) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t mine = 0u;
  int32_t* row = neg_ent + (size_t)i * K;
  for (int j = 0; j < K; ++j) {
    int s;
    int32_t l;
    smap.locate(row[j], s, l);
    if (s == my_shard)
      mine |= 1u << j;
    else
      row[j] = dummy_id;
  }
  neg_valid[i] = mine;
}
}  // namespace mke

extern "C" int mke_neg_keep_owned(int32_t* neg_ent, int32_t n, int32_t K, int32_t n_shards, int32_t shard_split,
                                  int32_t my_shard, int32_t dummy_id, uint32_t* neg_valid, mke_stream_t stream) {
  MKE_CHECK_ARG(n >= 0 && K >= 1 && K <= MKE_MAX_NEG, "bad n / K");
  MKE_CHECK_ARG(n_shards == 2 || n_shards == 4 || n_shards == 8, "n_shards=%d (2, 4 or 8)", n_shards);
  MKE_CHECK_ARG(my_shard >= 0 && my_shard < n_shards && shard_split >= 0, "bad shard / split");
  if (n == 0) return 0;
  MKE_CHECK_ARG(neg_ent && neg_valid, "null pointer");
  mke_table_t t{};
  t.n_shards = n_shards;
  t.shard_split = shard_split;
  const mke::ShardMap smap = mke::shard_map(&t);
  {
    int s;
    int32_t l;
    smap.locate(dummy_id, s, l);
    MKE_CHECK_ARG(s == my_shard, "dummy row %d does not live on shard %d", dummy_id, my_shard);
  }
  mke::neg_keep_owned_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(neg_ent, n, K, smap, my_shard,
                                                                               dummy_id, neg_valid);
  MKE_CHECK_LAUNCH("neg_keep_owned_kernel");
  return 0;
}

// ---- the same, COMPACTED: this rank's negatives first (in their original order), then dummies ------------------
// neg_valid[i] = low_ones(count) and the side word is permuted along, so that bit j still belongs to slot j.  The
// one-wave phase-1 kernel then stops its K-loop at the longest list of a warp's four positives
// (mke_rel_step_structured4 with compact = 1): at 8 ranks a positive keeps 2.5 of its 10 negatives on average.
namespace mke {
__global__ void neg_keep_owned_compact_kernel(int32_t* __restrict__ neg_ent, uint32_t* __restrict__ neg_side, int n, int K,
                                              ShardMap smap, int my_shard, int32_t dummy_id,
                                              uint32_t* __restrict__ neg_valid) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int32_t* row = neg_ent + (size_t)i * K;
  const uint32_t side = neg_side[i];
  uint32_t new_side = 0u;
  int cnt = 0;
  for (int j = 0; j < K; ++j) {
    const int32_t e = row[j];
    int s;
    int32_t l;
    smap.locate(e, s, l);
    if (s == my_shard) {
      row[cnt] = e;  // cnt <= j: never overwrites an entry not yet read
      new_side |= ((side >> j) & 1u) << cnt;
      ++cnt;
    }
  }
  for (int j = cnt; j < K; ++j) row[j] = dummy_id;
  neg_side[i] = new_side;
  neg_valid[i] = cnt >= 32 ? 0xffffffffu : ((1u << cnt) - 1u);
}
}  // namespace mke

extern "C" int mke_neg_keep_owned2(int32_t* neg_ent, uint32_t* neg_side, int32_t n, int32_t K, int32_t n_shards,
                                   int32_t shard_split, int32_t my_shard, int32_t dummy_id, uint32_t* neg_valid,
                                   mke_stream_t stream) {
  MKE_CHECK_ARG(n >= 0 && K >= 1 && K <= MKE_MAX_NEG, "bad n / K");
  MKE_CHECK_ARG(n_shards == 2 || n_shards == 4 || n_shards == 8, "n_shards=%d (2, 4 or 8)", n_shards);
  MKE_CHECK_ARG(my_shard >= 0 && my_shard < n_shards && shard_split >= 0, "bad shard / split");
  if (n == 0) return 0;
  MKE_CHECK_ARG(neg_ent && neg_side && neg_valid, "null pointer");
  mke_table_t t{};
  t.n_shards = n_shards;
  t.shard_split = shard_split;
  const mke::ShardMap smap = mke::shard_map(&t);
  {
    int s;
    int32_t l;
    smap.locate(dummy_id, s, l);
    MKE_CHECK_ARG(s == my_shard, "dummy row %d does not live on shard %d", dummy_id, my_shard);
  }
  mke::neg_keep_owned_compact_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(neg_ent, neg_side, n, K, smap,
                                                                                       my_shard, dummy_id, neg_valid);
  MKE_CHECK_LAUNCH("neg_keep_owned_compact_kernel");
  return 0;
}

// ---- random.sample(range(n), count) without a sort ------------------------------------------------------------------
// The cross-KG / entity batches of the reference are random.sample(list, B) every step (MultiKE_model.py:355-358, :377,
// :399, :422, :443, :462).  A sample without replacement is the first `count` images of a random permutation of [0, n):
// here a keyed Feistel network over the smallest even-width power-of-two domain >= n, cycle-walked back into [0, n)
// (a bijection for every key; 3.5 us for 5 000 picks where torch.randperm(n)[:B] sorts n keys: 135 us at n = 554 173).
namespace mke {
__device__ __forceinline__ uint32_t feistel_permute(uint32_t x, int half_bits, uint64_t key) {
  const uint32_t mask = (1u << half_bits) - 1u;
  uint32_t l = x >> half_bits, r = x & mask;
#pragma unroll
  for (int round = 0; round < 6; ++round) {
    const uint32_t f = (uint32_t)(mix64(key + (uint64_t)(round + 1) * kGamma + r) >> 20) & mask;
    const uint32_t nl = r;
    r = l ^ f;
    l = nl;
  }
  return (l << half_bits) | r;
}
__global__ void sample_distinct_kernel(uint32_t n, int count, int half_bits, uint64_t key, int32_t* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  uint32_t x = (uint32_t)i;
  do {
    x = feistel_permute(x, half_bits, key);
  } while (x >= n);  // cycle walking: the domain is < 4 n, so two tries on average at most
  out[i] = (int32_t)x;
}
}  // namespace mke

extern "C" int mke_sample_distinct(int32_t n, int32_t count, uint64_t seed, uint64_t draw, int32_t* out, mke_stream_t stream) {
  MKE_CHECK_ARG(n > 0 && count >= 0 && count <= n, "count=%d outside [0, n=%d]", count, n);
  if (count == 0) return 0;
  MKE_CHECK_ARG(out, "null output");
  int half_bits = 1;
  while ((1ull << (2 * half_bits)) < (unsigned long long)n) ++half_bits;
  const uint64_t key = mke::stream_key(seed ^ 0x5DEECE66Dull, draw);
  mke::sample_distinct_kernel<<<(count + 255) / 256, 256, 0, (cudaStream_t)stream>>>((uint32_t)n, count, half_bits, key, out);
  MKE_CHECK_LAUNCH("sample_distinct_kernel");
  return 0;
}
