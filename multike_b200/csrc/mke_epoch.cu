// Step / epoch driver: the inner loop of MultiKE.train_relation_view_1epo
// (MultiKE_model.py:302-313) as a sequence of kernel launches on two streams, issued from C so
// that a step costs no interpreter time.  Batching follows base/batch.py:33-54.
#include <cstdio>
#include <cstdlib>
#include <utility>
#include <vector>
#include "mke_rel_persist.cuh"

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
// used by mke_sharded.cu: an event pair around a launch of the timed kernel, when timing is on
bool timer_begin(cudaStream_t st) {
  const bool timed = g_timer.used + 2 <= (int)g_timer.ev.size() && (g_timer.seen++ % g_timer.every) == 0;
  if (timed) cudaEventRecord(g_timer.ev[g_timer.used], st);
  return timed;
}
void timer_end(cudaStream_t st) {
  cudaEventRecord(g_timer.ev[g_timer.used + 1], st);
  g_timer.used += 2;
}

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


// ---- persistent step kernel (variant 4): workspace layout and driver ----------------------------
struct PersistLayout {
  size_t sync, trace, flags[2], stage1[2], stage2[2], total;
};
static PersistLayout persist_layout(int chunk, int b1, int b2) {
  auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
  PersistLayout L{};
  size_t off = 0;
  L.sync = off;
  off = up(off + kSyncWords * sizeof(uint32_t));
  L.trace = off;
  off = up(off + (2 * (size_t)chunk + 2) * sizeof(unsigned long long));
  for (int k = 0; k < 2; ++k) {
    L.flags[k] = off;
    off = up(off + (size_t)chunk * sizeof(uint32_t));
  }
  for (int k = 0; k < 2; ++k) {
    L.stage1[k] = off;
    off = up(off + (size_t)chunk * (b1 > 0 ? b1 : 1) * 3 * sizeof(int32_t));
    L.stage2[k] = off;
    off = up(off + (size_t)chunk * (b2 > 0 ? b2 : 1) * 3 * sizeof(int32_t));
  }
  L.total = off;
  return L;
}
static void split_b(const mke_rel_view_t* v, int& b1, int& b2) {
  b1 = (int)((double)v->n1 / ((double)v->n1 + (double)v->n2) * (double)v->batch_size);
  b2 = v->batch_size - b1;
}

static unsigned g_persist_seq = 0;  // launch counter: the value a launch's batch flags are set to (mod 256)

// returns 1 when the persistent kernel does not cover this view (caller uses one launch per phase)
static int train_steps_persistent(const mke_rel_view_t* v, int first_step, int n_steps, uint64_t first_global_step,
                                  bool host_fed, int64_t* positives_out, cudaStream_t main, cudaStream_t side) {
  const int chunk = v->persist_chunk;
  if (chunk < 1 || chunk > kPersistMaxChunk || v->persist_ws == nullptr || v->K < 1 || !v->neg_ent[0] || !v->neg_ent[1] || !v->neg_side[0] ||
      !v->neg_side[1] || v->ent->n_shards > 1)
    return 1;
  if (host_fed && (v->persist_flag_src == nullptr || side == nullptr || side == main)) return 1;
  int b1, b2;
  split_b(v, b1, b2);
  const PersistLayout L = persist_layout(chunk, b1, b2);
  MKE_CHECK_ARG((size_t)v->persist_ws_bytes >= L.total, "persist_ws_bytes %lld < %zu (mke_rel_persist_workspace_bytes)",
                (long long)v->persist_ws_bytes, L.total);
  unsigned char* ws = (unsigned char*)v->persist_ws;
  const int steps_per_epoch = (int)(((long long)v->n1 + v->n2 + v->batch_size - 1) / v->batch_size);

  RelStepParams p{};
  p.pos_own_lo = 0;
  p.pos_own_hi = 0x7fffffff;
  fill_tables(p, v->ent, v->rel);
  if (v->kg1) p.kg1 = *v->kg1;
  if (v->kg2) p.kg2 = *v->kg2;
  p.K = v->K;
  p.pos_scale = 1.0f;
  PersistParams q{};
  q.n1 = v->n1;
  q.n2 = v->n2;
  q.b1 = b1;
  q.b2 = b2;
  q.steps_per_epoch = steps_per_epoch;
  q.seed = v->seed;
  for (int k = 0; k < 2; ++k) {
    q.neg_ent[k] = v->neg_ent[k];
    q.neg_side[k] = v->neg_side[k];
  }
  q.A = apply_table(v->ent, v->ent_acc, v->lr);
  q.B = apply_table(v->rel, v->rel_acc, v->lr);
  q.sync = (uint32_t*)(ws + L.sync);
  q.trace = (unsigned long long*)(ws + L.trace);
  static const int apply_mode = getenv("MKE_PERSIST_APPLY") ? atoi(getenv("MKE_PERSIST_APPLY")) : 1;
  static const int apply_chunk = getenv("MKE_PERSIST_CHUNK") ? atoi(getenv("MKE_PERSIST_CHUNK")) : 32;
  static const int samp_phase = getenv("MKE_PERSIST_SAMP_PHASE") ? atoi(getenv("MKE_PERSIST_SAMP_PHASE")) : 1;
  q.samp_phase = samp_phase;
  static const int fence_mode = getenv("MKE_PERSIST_FENCE") ? atoi(getenv("MKE_PERSIST_FENCE")) : 0;
  q.fence_mode = fence_mode;

  q.apply_mode = apply_mode;
  q.apply_chunk = apply_chunk == 32 ? 32 : 16;
  double* host_loss_dev = nullptr;  // the pinned loss buffer as the device sees it
  if (v->host_step_loss != nullptr) {
    void* d = nullptr;
    if (cudaHostGetDevicePointer(&d, (void*)v->host_step_loss, 0) == cudaSuccess)
      host_loss_dev = (double*)d;
    else
      cudaGetLastError();
  }
  int dev = 0;
  cudaGetDevice(&dev);
  EventPool& pool = g_events[dev & 15];
  cudaEvent_t done[2] = {nullptr, nullptr};
  if (host_fed) {  // earlier work on main may still read the staging buffers
    cudaEvent_t e = pool.get();
    cudaEventRecord(e, main);
    cudaStreamWaitEvent(side, e, 0);
  }
  long long positives = 0;
  for (int c0 = 0; c0 < n_steps; c0 += chunk) {
    const int len = n_steps - c0 < chunk ? n_steps - c0 : chunk;
    const unsigned seq = ++g_persist_seq;
    const int buf = (int)(seq & 1u);
    q.first_step = (first_step + c0) % steps_per_epoch;
    q.n_steps = len;
    q.first_global_step = first_global_step + (uint64_t)c0;
    q.step_loss = v->step_loss + c0;
    q.host_loss = host_loss_dev ? host_loss_dev + c0 : nullptr;
    for (int k = 0; k < len; ++k) {
      const Slice sl = step_slice(v, (q.first_step + k) % steps_per_epoch);
      positives += sl.len1 + sl.len2;
    }
    // Host fed: the batches of this launch travel on `side` while the kernel runs; the kernel waits for flags[k]
    // before it reads step k.  Only step 0 is copied before the launch (the kernel's prologue needs it), the other
    // steps after it, so that a short run does not wait for len x 3 copy calls before its kernel starts.
    int32_t *s1 = nullptr, *s2 = nullptr;
    uint32_t* fl = nullptr;
    auto copy_steps = [&](int k0, int k1) -> cudaError_t {
      for (int k = k0; k < k1; ++k) {
        const Slice sl = step_slice(v, (q.first_step + k) % steps_per_epoch);
        cudaError_t e = cudaSuccess;
        if (sl.len1 > 0)
          e = cudaMemcpyAsync(s1 + 3 * (size_t)k * b1, v->host_triples1 + 3 * (size_t)sl.a1, (size_t)sl.len1 * 12,
                              cudaMemcpyHostToDevice, side);
        if (e == cudaSuccess && sl.len2 > 0)
          e = cudaMemcpyAsync(s2 + 3 * (size_t)k * b2, v->host_triples2 + 3 * (size_t)sl.a2, (size_t)sl.len2 * 12,
                              cudaMemcpyHostToDevice, side);
        if (e == cudaSuccess)  // lands after the batch (stream order): "step k is here"
          e = cudaMemcpyAsync(fl + k, v->persist_flag_src + (seq & 255u), sizeof(uint32_t), cudaMemcpyHostToDevice, side);
        if (e != cudaSuccess) return e;
      }
      return cudaSuccess;
    };
    if (host_fed) {
      s1 = (int32_t*)(ws + L.stage1[buf]);
      s2 = (int32_t*)(ws + L.stage2[buf]);
      fl = (uint32_t*)(ws + L.flags[buf]);
      if (done[buf] != nullptr) cudaStreamWaitEvent(side, done[buf], 0);  // the launch that last read this buffer
      if (cudaError_t e = copy_steps(0, 1)) return cuda_fail(e, "H2D batch");
      q.t1 = q.t2 = nullptr;
      q.st1 = s1;
      q.st2 = s2;
      q.flags = fl;
      q.flag_value = seq & 255u;
    } else {
      q.t1 = v->triples1;
      q.t2 = v->triples2;
      q.st1 = q.st2 = nullptr;
      q.flags = nullptr;
    }
    if (cudaError_t e = cudaMemsetAsync(q.sync, 0, kSyncWords * sizeof(uint32_t), main)) return cuda_fail(e, "memset sync");
    // debug: MKE_PERSIST_BLOCKTRACE=<file> dumps the per-block barrier stamps of launch number MKE_PERSIST_BLOCKTRACE_AT
    static const char* bt_path = getenv("MKE_PERSIST_BLOCKTRACE");
    static const int bt_at = getenv("MKE_PERSIST_BLOCKTRACE_AT") ? atoi(getenv("MKE_PERSIST_BLOCKTRACE_AT")) : 3;
    static int bt_calls = 0;
    unsigned long long* bt_buf = nullptr;
    const size_t bt_words = (size_t)(2 * len + 2) * sm_count() * 4;
    q.block_trace = nullptr;
    if (bt_path != nullptr && bt_calls++ == bt_at && cudaMalloc(&bt_buf, bt_words * 8) == cudaSuccess) {
      cudaMemsetAsync(bt_buf, 0, bt_words * 8, main);
      q.block_trace = bt_buf;
    }
    const bool timed = g_timer.used + 2 <= (int)g_timer.ev.size() && (g_timer.seen++ % g_timer.every) == 0;
    if (timed) cudaEventRecord(g_timer.ev[g_timer.used], main);
    const int rc = launch_rel_persist(p, q, main);
    if (rc == 0 && host_fed)
      if (cudaError_t e = copy_steps(1, len)) return cuda_fail(e, "H2D batch");
    if (bt_buf != nullptr) {
      cudaStreamSynchronize(main);
      std::vector<unsigned long long> host(bt_words + 2);
      host[0] = (unsigned long long)(2 * len + 2);
      host[1] = (unsigned long long)sm_count();
      cudaMemcpy(host.data() + 2, bt_buf, bt_words * 8, cudaMemcpyDeviceToHost);
      if (FILE* f = fopen(bt_path, "wb")) {
        fwrite(host.data(), 8, host.size(), f);
        fclose(f);
      }
      cudaFree(bt_buf);
    }
    if (timed) {
      cudaEventRecord(g_timer.ev[g_timer.used + 1], main);
      g_timer.used += 2;
    }
    if (rc != 0) {
      MKE_CHECK_ARG(rc < 0 || c0 == 0, "persistent launch shape changed between chunks");
      return rc;
    }

    if (host_fed) {
      done[buf] = pool.get();
      cudaEventRecord(done[buf], main);
    }
    if (v->host_step_loss != nullptr && host_loss_dev == nullptr)
      if (cudaError_t e = cudaMemcpyAsync(v->host_step_loss + c0, v->step_loss + c0, sizeof(double) * len,
                                          cudaMemcpyDeviceToHost, main))
        return cuda_fail(e, "D2H loss");
  }
  if (host_fed) {  // join: every copy on `side` has been consumed by a kernel on main, keep the streams ordered anyway
    cudaEvent_t e = pool.get();
    cudaEventRecord(e, side);
    cudaStreamWaitEvent(main, e, 0);
  }
  if (positives_out != nullptr) *positives_out = positives;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "mke_rel_train_steps (persistent)");
  return 0;
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
  if (n_steps > 0)  // every step ADDS its loss to step_loss[s]
    if (cudaError_t e = cudaMemsetAsync(v->step_loss, 0, sizeof(double) * (size_t)n_steps, (cudaStream_t)main_))
      return cuda_fail(e, "memset step_loss");
  if (v->variant == 4) {  // persistent step kernel; shapes it does not cover run one launch per phase (variant 3)
    const int rc = train_steps_persistent(v, first_step, n_steps, first_global_step, host_fed, positives_out,
                                          (cudaStream_t)main_, (cudaStream_t)side_);
    if (rc <= 0) return rc;
  }
  const int variant = v->variant == 4 ? 3 : v->variant;
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
                                      v->neg_side[s & 1], nullptr, 1.0f, v->step_loss + s, variant, main);
      else
        rc = mke_rel_step_sampled(v->ent, v->rel, c1, cur.len1, v->kg1, c2, cur.len2, v->kg2, v->K, v->seed,
                                  first_global_step + (uint64_t)s, nullptr, 1.0f, v->step_loss + s, nullptr,
                                  variant, main);
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

extern "C" int64_t mke_rel_persist_workspace_bytes(int32_t n1, int32_t n2, int32_t batch_size, int32_t chunk_steps) {
  if (n1 < 0 || n2 < 0 || (long long)n1 + n2 <= 0 || batch_size < 1 || chunk_steps < 1) return 0;
  mke_rel_view_t v{};
  v.n1 = n1;
  v.n2 = n2;
  v.batch_size = batch_size;
  int b1, b2;
  split_b(&v, b1, b2);
  return (int64_t)persist_layout(chunk_steps, b1, b2).total;
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
