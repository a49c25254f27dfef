// Quarter-warp negative sampler shared by the fused relation kernel (mke_rel_q8.cu) and the
// stand-alone sampling kernels (mke_sampler.cu).
#pragma once
#include "mke_rel.cuh"

namespace mke {

constexpr int kQPerWarp = 4;     // positives per warp
constexpr int kPickStride = 33;  // MKE_MAX_NEG + 1: the four quarters of a warp hit distinct banks
constexpr uint32_t kFull = 0xffffffffu;

// ---- sampler, quarter layout --------------------------------------------------------------
struct KgView {  // mke_kg_sampler_t selected per quarter (kg1 / kg2), held in registers
  const int32_t* list;
  const int32_t* neighbours;
  mke_tripleset_t set;
  int32_t base, n, n_nb;
  __device__ __forceinline__ CandPool pool(int32_t anchor) const {
    CandPool c;
    if (neighbours != nullptr) {
      const int32_t* row = neighbours + (size_t)anchor * (size_t)n_nb;
      if (__ldg(row) >= 0) {
        c.list = row;
        c.base = 0;
        c.n = (uint32_t)n_nb;
        return c;
      }
    }
    c.list = list;
    c.base = base;
    c.n = (uint32_t)n;
    return c;
  }
};
__device__ __forceinline__ KgView kg_view(const mke_kg_sampler_t& kg1, const mke_kg_sampler_t& kg2,
                                          bool first) {
  KgView k;
  k.list = first ? kg1.entity_list : kg2.entity_list;
  k.neighbours = first ? kg1.neighbours : kg2.neighbours;
  k.set.slots = first ? kg1.set.slots : kg2.set.slots;
  k.set.capacity = first ? kg1.set.capacity : kg2.set.capacity;
  k.base = first ? kg1.entity_base : kg2.entity_base;
  k.n = first ? kg1.n_entities : kg2.n_entities;
  k.n_nb = first ? kg1.n_neighbours : kg2.n_neighbours;
  return k;
}
__device__ __forceinline__ KgView kg_view(const RelStepParams& p, bool first) {
  return kg_view(p.kg1, p.kg2, first);
}
__device__ __forceinline__ uint32_t low_ones(int k) { return (k >= 32) ? kFull : ((1u << k) - 1u); }

__device__ __forceinline__ unsigned long long gtimer_raw() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)::"memory");
  return t;
}

// generate_neg_triples_fast (base/batch.py:86-116) for ONE positive, executed by the 8 lanes of a
// quarter with exactly the sequential semantics that oracle/device_sampler.py restates:
//   per round (<= MKE_MAX_TRY): one head/tail coin; candidates c = 0, 1, 2, ... are drawn from the
//   pool of the replaced entity and accepted unless they repeat an entity already accepted in this
//   round (random.sample = without replacement) until `remaining` are accepted; accepted
//   candidates that are known triples are dropped (not in the last round); stop at K.
// Draws are counter based, so the 8 lanes evaluate candidates c_base + sub of a chunk at once
// and ranks inside the chunk reproduce the sequential order.  All synchronisation is scoped to
// the quarter (qmask): the four quarters of a warp may be in different rounds.
// Writes pick[0..K) and returns the side mask (bit j: negative j replaces the head).
__device__ __forceinline__ uint32_t sample_negs_quarter(const KgView& kg, int32_t h, int32_t r,
                                                        int32_t t, int K, uint64_t skey, uint32_t i,
                                                        int lane, volatile int32_t* pick,
                                                        unsigned long long* tr_slot = nullptr) {
  const int sub = lane & 7;
  const int qshift = lane & 24;
  const uint32_t qmask = 0xffu << qshift;
  const uint32_t below = (1u << sub) - 1u;  // lower lanes of my quarter, after shifting to bit 0
  int n_acc = 0, remaining = K;
  uint32_t side_mask = 0;
  for (uint32_t tr = 0; tr < MKE_MAX_TRY; ++tr) {
    const bool head_side = (draw64(skey, i, tr, kSideDraw) >> 63) != 0;
    const CandPool pool = kg.pool(head_side ? h : t);
    if (tr_slot && tr == 0 && lane == 0) tr_slot[16] = gtimer_raw();
    // ---- draw: accept the first `remaining` candidates that do not repeat an accepted one ----
    int np = 0;
    // one chunk = candidates c .. c+7 (one per lane), accepted in lane order
    auto accept = [&](int32_t e, uint32_t c) {
      bool dup = false;
#pragma unroll
      for (int d = 1; d < 8; ++d) {  // same entity drawn by a lower lane of this chunk?
        const int32_t v = __shfl_up_sync(qmask, e, d, 8);
        dup |= (sub >= d) && (v == e);
      }
      for (int k = 0; k < np; ++k) dup |= (pick[n_acc + k] == e);
      dup = dup && (c + 1u < kSideDraw);
      // the ballot is also the barrier between the reads of pick[] above and the writes below
      const uint32_t fresh = (__ballot_sync(qmask, !dup) >> qshift) & 0xffu;
      const int rank = __popc(fresh & below);
      if (!dup && np + rank < remaining) pick[n_acc + np + rank] = e;
      __syncwarp(qmask);
      np = min(remaining, np + __popc(fresh));
    };
    for (uint32_t c_base = 0; np < remaining; c_base += 16) {
      const uint32_t c0 = c_base + (uint32_t)sub, c1 = c0 + 8u;
      const int32_t e0 = pool.at(draw_index(draw64(skey, i, tr, c0), pool.n));
      const int32_t e1 = pool.at(draw_index(draw64(skey, i, tr, c1), pool.n));
      accept(e0, c0);
      if (np < remaining) accept(e1, c1);
    }
    if (tr_slot && tr == 0 && lane == 0) tr_slot[17] = gtimer_raw();
    // ---- filter: drop known triples (the last round is accepted as is, batch.py:103-105) ------
    int kept = np;
    if (tr != MKE_MAX_TRY - 1) {
      kept = 0;
      for (int k0 = 0; k0 < np; k0 += 16) {  // two picks per lane, both probes in flight
        const int ka = k0 + sub, kb = ka + 8;
        const int32_t ea = pick[n_acc + (ka < np ? ka : 0)], eb = pick[n_acc + (kb < np ? kb : 0)];
        const uint64_t keya = head_side ? triple_key(ea, r, t) : triple_key(h, r, ea);
        const uint64_t keyb = head_side ? triple_key(eb, r, t) : triple_key(h, r, eb);
        bool ina, inb;
        tripleset_contains2(kg.set, keya, keyb, ina, inb);
        const bool keepa = (ka < np) && !ina, keepb = (kb < np) && !inb;
        // the ballots are the barrier between reading pick[] above and compacting it below
        const uint32_t ba = (__ballot_sync(qmask, keepa) >> qshift) & 0xffu;
        const uint32_t bb = (__ballot_sync(qmask, keepb) >> qshift) & 0xffu;
        __syncwarp(qmask);  // (racecheck does not count a ballot as ordering the shared-memory reads above)
        if (keepa) pick[n_acc + kept + __popc(ba & below)] = ea;
        if (keepb) pick[n_acc + kept + __popc(ba) + __popc(bb & below)] = eb;
        __syncwarp(qmask);
        kept += __popc(ba) + __popc(bb);
      }
    }
    if (tr_slot && tr == 0 && lane == 0) tr_slot[18] = gtimer_raw();
    if (head_side && kept > 0) side_mask |= low_ones(kept) << n_acc;
    n_acc += kept;
    if (n_acc >= K) break;
    remaining = K - n_acc;
  }
  return side_mask;
}

}  // namespace mke
