// Device-resident triple set + stand-alone negative sampler.
// Replaces the Python set membership test and generate_neg_triples_fast of
// base/batch.py:86-116 (all_triples_set = kg.local_relation_triples_set, which aliases
// relation_triples_set and therefore also holds the swapped sup triples: base/kg.py:59,134).
#include "mke_common.cuh"

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

constexpr int kSampThreads = 256;
constexpr int kSampWarps = kSampThreads / 32;

__global__ void __launch_bounds__(kSampThreads)
    sample_kernel(const int32_t* __restrict__ pos1, int len1, mke_kg_sampler_t kg1,
                  const int32_t* __restrict__ pos2, int len2, mke_kg_sampler_t kg2, int K,
                  uint64_t skey, int32_t* __restrict__ neg_out) {
  __shared__ int32_t s_pick_all[kSampWarps][32];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int n = len1 + len2;
  for (int i = blockIdx.x * kSampWarps + wib; i < n; i += gridDim.x * kSampWarps) {
    const bool first = i < len1;
    const int32_t* row = first ? pos1 + 3 * (size_t)i : pos2 + 3 * (size_t)(i - len1);
    const int32_t h = __ldg(row), r = __ldg(row + 1), t = __ldg(row + 2);
    int32_t e;
    uint32_t side;
    sample_negs_warp(first ? kg1 : kg2, h, r, t, K, skey, (uint32_t)i, lane, s_pick_all[wib], e,
                     side);
    if (lane < K) {
      const bool hs = (side >> lane) & 1u;
      int32_t* o = neg_out + ((size_t)i * K + lane) * 3;
      o[0] = hs ? e : h;
      o[1] = r;
      o[2] = hs ? t : e;
    }
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
  MKE_CHECK_ARG(K >= 1 && K <= MKE_MAX_NEG, "K=%d outside [1,%d]", K, MKE_MAX_NEG);
  MKE_CHECK_ARG(len1 >= 0 && len2 >= 0, "negative batch length");
  MKE_CHECK_ARG(len1 == 0 || (pos1 && kg1), "kg1 slice needs positives and a sampler");
  MKE_CHECK_ARG(len2 == 0 || (pos2 && kg2), "kg2 slice needs positives and a sampler");
  const int n = len1 + len2;
  if (n == 0) return 0;
  MKE_CHECK_ARG(neg_out, "neg_out is null");
  mke_kg_sampler_t a{}, b{};
  if (kg1) a = *kg1;
  if (kg2) b = *kg2;
  for (const mke_kg_sampler_t* kg : {len1 ? kg1 : nullptr, len2 ? kg2 : nullptr}) {
    if (!kg) continue;
    MKE_CHECK_ARG(kg->n_entities >= K, "KG has fewer entities (%d) than K=%d", kg->n_entities, K);
    MKE_CHECK_ARG(!kg->neighbours || kg->n_neighbours >= K, "n_neighbours < K");
  }
  int blocks = (n + kSampWarps - 1) / kSampWarps;
  const int full = sm_count() * 8;
  if (blocks > full) blocks = full;
  sample_kernel<<<blocks, kSampThreads, 0, (cudaStream_t)stream>>>(pos1, len1, a, pos2, len2, b, K,
                                                                  stream_key(seed, step), neg_out);
  MKE_CHECK_LAUNCH("sample_kernel");
  return 0;
}
