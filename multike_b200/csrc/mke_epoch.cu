// Step / epoch driver: the inner loop of MultiKE.train_relation_view_1epo
// (MultiKE_model.py:302-313) as a sequence of kernel launches on two streams, issued from C so
// that a step costs no interpreter time.  Batching follows base/batch.py:33-54.
#include <cstdlib>
#include <utility>
#include <vector>
#include "mke_common.cuh"

namespace mke {

struct EventPool {
  cudaEvent_t ev[8];
  bool ready = false;
  int next = 0;
  cudaEvent_t get() {
    if (!ready) {
      for (auto& e : ev) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
      ready = true;
    }
    cudaEvent_t e = ev[next];
    next = (next + 1) & 7;
    return e;
  }
};
static thread_local EventPool g_events[16];  // per device

// Optional CUDA-event timing of every phase-1 launch issued by the driver (bench.py's roofline
// figure: the kernel's average launch duration measured live inside the timed region).
struct PhaseTimer {
  std::vector<cudaEvent_t> ev;  // pairs
  int used = 0;
  int every = 1;  // time one launch in `every` (a timed event pair costs the step a few microseconds)
  long long seen = 0;
};
static PhaseTimer g_timer;

struct Slice {
  int a1, len1, a2, len2;
};
static Slice step_slice(const mke_rel_view_t* v, int step) {
  // base/batch.py:36-37: kg1's share is floored; :45-54: slices clipped at the list end
  const int b1 = (int)((double)v->n1 / ((double)v->n1 + (double)v->n2) * (double)v->batch_size);
  const int b2 = v->batch_size - b1;
  auto clip = [](long long start, int bs, int n) {
    const int s = (int)(start < n ? start : n);
    const int e = (int)(start + bs < n ? start + bs : n);
    return std::pair<int, int>(s, e - s);
  };
  const auto s1 = clip((long long)step * b1, b1, v->n1);
  const auto s2 = clip((long long)step * b2, b2, v->n2);
  return Slice{s1.first, s1.second, s2.first, s2.second};
}

}  // namespace mke

using namespace mke;

extern "C" int mke_rel_train_steps(const mke_rel_view_t* v, int32_t first_step, int32_t n_steps,
                                   uint64_t first_global_step, int64_t* positives_out,
                                   mke_stream_t main_, mke_stream_t side_) {
  MKE_CHECK_ARG(v && v->ent && v->rel && v->ent_acc && v->rel_acc, "view needs tables and Adagrad slots");
  MKE_CHECK_ARG(v->n1 >= 0 && v->n2 >= 0 && (long long)v->n1 + v->n2 > 0, "empty triple lists");
  MKE_CHECK_ARG(v->batch_size > 0 && v->K >= 0 && v->K <= MKE_MAX_NEG, "bad batch_size / K");
  MKE_CHECK_ARG(n_steps >= 0 && first_step >= 0, "bad step range");
  MKE_CHECK_ARG(v->step_loss, "step_loss is null");
  const bool host_fed = v->host_triples1 != nullptr || v->host_triples2 != nullptr;
  MKE_CHECK_ARG(host_fed || ((v->n1 == 0 || v->triples1) && (v->n2 == 0 || v->triples2)),
                "device triple lists are null");
  MKE_CHECK_ARG(!host_fed || ((v->n1 == 0 || (v->host_triples1 && v->stage1[0] && v->stage1[1])) &&
                              (v->n2 == 0 || (v->host_triples2 && v->stage2[0] && v->stage2[1]))),
                "host-fed batches need pinned triples and two staging buffers per KG");
  const bool ahead = v->K > 0 && v->neg_ent[0] && v->neg_ent[1] && v->neg_side[0] && v->neg_side[1] &&
                     side_ != nullptr && side_ != main_;
  cudaStream_t main = (cudaStream_t)main_, side = ahead ? (cudaStream_t)side_ : (cudaStream_t)main_;
  // Experiment knob (MKE_L2_PERSIST=<MB>): pin the entity gradient table in L2 -- phase 1 reduces
  // into it and phase 2 reads and re-zeroes it, so it need not travel to HBM in between.
  static const int l2_mb = getenv("MKE_L2_PERSIST") ? atoi(getenv("MKE_L2_PERSIST")) : 0;
  if (l2_mb > 0) {
    static bool limit_set = false;
    if (!limit_set) {
      cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)l2_mb << 20);
      limit_set = true;
    }
    cudaStreamAttrValue attr{};
    attr.accessPolicyWindow.base_ptr = v->ent->grad;
    attr.accessPolicyWindow.num_bytes = (size_t)table_local_rows(v->ent) * v->ent->stride * sizeof(float);
    attr.accessPolicyWindow.hitRatio = 1.0f;
    attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    cudaStreamSetAttribute((cudaStream_t)main_, cudaStreamAttributeAccessPolicyWindow, &attr);
    cudaGetLastError();
  }
  const int steps_per_epoch = (int)(((long long)v->n1 + v->n2 + v->batch_size - 1) / v->batch_size);
  int dev = 0;
  cudaGetDevice(&dev);
  EventPool& pool = g_events[dev & 15];
  long long positives = 0;

  // positives of step `s` (device pointers), copying them in first when they live on the host
  auto stage = [&](int s, int step_in_epoch, cudaStream_t st, const int32_t*& p1, const int32_t*& p2, Slice& sl) {
    sl = step_slice(v, step_in_epoch);
    if (!host_fed) {
      p1 = v->triples1 ? v->triples1 + 3 * (size_t)sl.a1 : nullptr;
      p2 = v->triples2 ? v->triples2 + 3 * (size_t)sl.a2 : nullptr;
      return cudaSuccess;
    }
    cudaError_t e = cudaSuccess;
    if (sl.len1 > 0)
      e = cudaMemcpyAsync(v->stage1[s & 1], v->host_triples1 + 3 * (size_t)sl.a1, (size_t)sl.len1 * 12,
                          cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && sl.len2 > 0)
      e = cudaMemcpyAsync(v->stage2[s & 1], v->host_triples2 + 3 * (size_t)sl.a2, (size_t)sl.len2 * 12,
                          cudaMemcpyHostToDevice, st);
    p1 = v->stage1[s & 1];
    p2 = v->stage2[s & 1];
    return e;
  };

  const int32_t *c1 = nullptr, *c2 = nullptr, *n1p = nullptr, *n2p = nullptr;
  Slice cur{}, nxt{};
  if (n_steps > 0) {
    // step 0: stage + (if drawn ahead) sample in line on the main stream
    if (cudaError_t e = stage(0, first_step % steps_per_epoch, main, c1, c2, cur)) return cuda_fail(e, "H2D batch");
    if (ahead && cur.len1 + cur.len2 > 0)
      if (int rc = mke_sample_structured(c1, cur.len1, v->kg1, c2, cur.len2, v->kg2, v->K, v->seed,
                                         first_global_step, v->neg_ent[0], v->neg_side[0], main))
        return rc;
  }
  for (int s = 0; s < n_steps; ++s) {
    const int n = cur.len1 + cur.len2;
    const bool have_next = s + 1 < n_steps;
    cudaEvent_t ready = nullptr;
    if (n > 0) {
      // ---- phase 1 ------------------------------------------------------------------------
      int rc;
      const bool timed = g_timer.used + 2 <= (int)g_timer.ev.size() && (g_timer.seen++ % g_timer.every) == 0;
      if (timed) cudaEventRecord(g_timer.ev[g_timer.used], main);
      if (ahead)
        rc = mke_rel_step_structured2(v->ent, v->rel, c1, cur.len1, c2, cur.len2, v->K, v->neg_ent[s & 1],
                                      v->neg_side[s & 1], nullptr, 1.0f, v->step_loss + s, v->variant, main);
      else
        rc = mke_rel_step_sampled(v->ent, v->rel, c1, cur.len1, v->kg1, c2, cur.len2, v->kg2, v->K, v->seed,
                                  first_global_step + (uint64_t)s, nullptr, 1.0f, v->step_loss + s, nullptr,
                                  v->variant, main);
      if (timed) {
        cudaEventRecord(g_timer.ev[g_timer.used + 1], main);
        g_timer.used += 2;
      }
      if (rc) return rc;
    }
    // ---- next step's batch (+ negatives) on the side stream, under this step's phase 2 -------
    if (have_next) {
      if (ahead) {
        cudaEvent_t fence = pool.get();
        cudaEventRecord(fence, main);  // after phase 1: the other buffers are free again
        cudaStreamWaitEvent(side, fence, 0);
      }
      if (cudaError_t e = stage(s + 1, (first_step + s + 1) % steps_per_epoch, side, n1p, n2p, nxt))
        return cuda_fail(e, "H2D batch");
      if (ahead) {
        if (nxt.len1 + nxt.len2 > 0)
          if (int rc = mke_sample_structured(n1p, nxt.len1, v->kg1, n2p, nxt.len2, v->kg2, v->K, v->seed,
                                             first_global_step + (uint64_t)(s + 1), v->neg_ent[(s + 1) & 1],
                                             v->neg_side[(s + 1) & 1], side))
            return rc;
        ready = pool.get();
        cudaEventRecord(ready, side);
      }
    }
    if (n > 0) {
      // ---- phase 2 ------------------------------------------------------------------------
      if (int rc = mke_rows_apply_adagrad_pair(v->ent, v->ent_acc, v->lr, v->rel, v->rel_acc, v->lr, main))
        return rc;
      if (v->host_step_loss != nullptr)
        if (cudaError_t e = cudaMemcpyAsync(v->host_step_loss + s, v->step_loss + s, sizeof(double),
                                            cudaMemcpyDeviceToHost, main))
          return cuda_fail(e, "D2H loss");
      positives += n;
    }
    if (ready != nullptr) cudaStreamWaitEvent(main, ready, 0);  // joins the side stream
    c1 = n1p;
    c2 = n2p;
    cur = nxt;
  }
  if (positives_out != nullptr) *positives_out = positives;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "mke_rel_train_steps");
  return 0;
}

extern "C" int mke_timing_enable(int32_t max_launches) {
  MKE_CHECK_ARG(max_launches >= 0 && max_launches <= (1 << 20), "bad max_launches");
  for (cudaEvent_t e : g_timer.ev) cudaEventDestroy(e);
  g_timer.ev.assign((size_t)max_launches * 2, nullptr);
  g_timer.used = 0;
  for (auto& e : g_timer.ev)
    if (cudaError_t err = cudaEventCreate(&e)) return cuda_fail(err, "cudaEventCreate");
  return 0;
}

extern "C" int mke_timing_stride(int32_t every) {
  MKE_CHECK_ARG(every >= 1, "every=%d", every);
  g_timer.every = every;
  g_timer.seen = 0;
  return 0;
}

extern "C" int mke_timing_read(double* total_ms, int32_t* launches) {
  MKE_CHECK_ARG(total_ms && launches, "null output");
  double sum = 0.0;
  for (int k = 0; k + 1 < g_timer.used; k += 2) {
    float ms = 0.f;
    if (cudaError_t err = cudaEventSynchronize(g_timer.ev[k + 1])) return cuda_fail(err, "cudaEventSynchronize");
    if (cudaError_t err = cudaEventElapsedTime(&ms, g_timer.ev[k], g_timer.ev[k + 1]))
      return cuda_fail(err, "cudaEventElapsedTime");
    sum += ms;
  }
  *total_ms = sum;
  *launches = g_timer.used / 2;
  g_timer.used = 0;
  return 0;
}
