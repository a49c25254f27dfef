// Multi-GPU relation view, step driver: the sequence of multike_b200/sharded.py's step --
// negatives, phase 1 over peer-mapped shards, sum of the replicated relation table's gradient bucket, phase 2 --
// issued from C for a run of steps, with NO collective library in the data path: the bucket (176 KB) is
// summed by a kernel that reads every rank's copy through its peer mapping, and the three points at which the
// ranks must wait for each other are flag barriers in peer memory.
//
//   K1 sampler [+ ownership filter]          (side stream, one step ahead, when a side stream is given)
//   K2 phase 1                                gathers rows / reduces gradient rows through peer pointers
//   K3 rel_exchange_kernel (cooperative,      a. my replicas -> my exchange buffer of this step's parity
//      32 blocks)                             b. barrier A (the last block to finish a. does it, the others wait on a
//                                                local flag): every rank has finished K2 (its peer reductions into
//                                                my shard have landed: a kernel's writes are complete when it ends)
//                                                and filled its exchange buffer
//                                             c. sum of all exchange buffers, in rank order (identical bits on
//                                                every rank) -> my gradient replica 0
//                                             (the buffers are double-buffered by step parity: a buffer is rewritten
//                                             two steps later, after barrier A of the step in between, at which
//                                             every rank has long finished reading it -- no barrier after c.)
//   K4 phase 2                                local shard + relation replica (MultiKE_model.py:15-31)
//   K5 peer_barrier_kernel                    barrier C: nobody starts the next phase 1 (peer reads of var, peer
//                                             reductions into grad) before everybody's phase 2 has ended
// A barrier: rank r stores the sequence number into slot r of every rank's flag array (st.release.sys after a
// system fence) and waits until all slots of its own array have reached it (ld.acquire.sys).
#include <cstdlib>
#include <utility>
#include "mke_common.cuh"

namespace mke {

bool timer_begin(cudaStream_t st);
void timer_end(cudaStream_t st);

constexpr unsigned long long kPeerWaitNs = 20000000000ull;  // two ranks may time-slice ONE GPU in the tests

struct PeerSync {
  uint32_t* flags[MKE_MAX_SHARDS];  // flags[k] = rank k's array of MKE_MAX_SHARDS words (peer-mapped)
  int world, rank;
};

__device__ __forceinline__ void peer_barrier(const PeerSync& ps, uint32_t seq) {
  __syncthreads();
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
      if (t1 - t0 > kPeerWaitNs) __trap();  // a rank that never arrives is a bug: fail, do not hang
    }
  }
  __syncthreads();
}

struct ExchangeParams {
  PeerSync ps;
  float* xchg[MKE_MAX_SHARDS];  // every rank's exchange buffers: 2 (step parity) x n4 float4
  float* grad;                  // my relation gradient: `replicas` copies of n4 float4 each
  uint32_t* local;              // my words [0]: block arrivals (monotonic), [1]: released-up-to sequence number
  int replicas;
  int n4;
  uint32_t seq;                 // sequence number of barrier A
  uint32_t launch_no;           // exchange launches before this one
};

constexpr int kXBlocks = 32, kXThreads = 256;

__global__ void __launch_bounds__(kXThreads) rel_exchange_kernel(const ExchangeParams p) {
  const size_t half = (size_t)(p.launch_no & 1u) * p.n4;
  float4* const mine = reinterpret_cast<float4*>(p.xchg[p.ps.rank]) + half;
  float4* const g = reinterpret_cast<float4*>(p.grad);
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  for (int i = tid; i < p.n4; i += nth) {
    float4 s = g[i];
    for (int r = 1; r < p.replicas; ++r) {
      s = f4_add(s, g[(size_t)r * p.n4 + i]);
      g[(size_t)r * p.n4 + i] = f4_zero();
    }
    mine[i] = s;
  }
  // grid-wide: the last block to get here talks to the peers, the others wait for its word
  __shared__ int s_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const uint32_t old = atomicAdd(p.local, 1u);
    s_last = (old + 1u == (p.launch_no + 1u) * gridDim.x) ? 1 : 0;
  }
  __syncthreads();
  if (s_last) {
    peer_barrier(p.ps, p.seq);
    if (threadIdx.x == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p.local + 1), "r"(p.seq) : "memory");
  } else if (threadIdx.x == 0) {
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0)::"memory");
    while (true) {
      uint32_t v;
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p.local + 1) : "memory");
      if ((int32_t)(v - p.seq) >= 0) break;
      __nanosleep(100);
      unsigned long long t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1)::"memory");
      if (t1 - t0 > kPeerWaitNs) __trap();
    }
  }
  __syncthreads();
  for (int i = tid; i < p.n4; i += nth) {
    float4 s = f4_zero();
    for (int k = 0; k < p.ps.world; ++k)  // rank order: the same sum on every rank
      s = f4_add(s, __ldcg(reinterpret_cast<const float4*>(p.xchg[k]) + half + i));
    g[i] = s;
  }
}

// barrier C; host fed: this rank's share of the step loss goes to pinned host memory by a store of the kernel (a
// cudaMemcpyAsync of 8 bytes per step on the main stream costs the step a copy-engine round trip)
__global__ void peer_barrier_kernel(const PeerSync ps, uint32_t seq, const double* loss_src, double* loss_host) {
  if (threadIdx.x == 0 && loss_host != nullptr) *reinterpret_cast<volatile double*>(loss_host) = *loss_src;
  peer_barrier(ps, seq);
}

// ---- which positives a rank trains in a global step (sharded.py: rank_parts / group_parts) ------------------------
struct Plan {
  const int32_t *p1, *p2;  // resolved by resolve(): device list + offset, or a staging buffer (host fed)
  int o1, o2;              // offsets (rows) into the kg1 / kg2 triple lists
  int l1, l2, base, lo, hi, mine;
};
static std::pair<int, int> clip(long long start, int bs, int n) {
  const int s = (int)(start < n ? start : n);
  const int e = (int)(start + bs < n ? start + bs : n);
  return {s, e - s};
}
static Plan make_plan(const mke_rel_sharded_view_t* v, int step) {
  // base/batch.py:36-37, 45-54 on the GLOBAL batch
  const int b1 = (int)((double)v->n1 / ((double)v->n1 + (double)v->n2) * (double)v->global_batch);
  const int b2 = v->global_batch - b1;
  const auto s1 = clip((long long)step * b1, b1, v->n1);
  const auto s2 = clip((long long)step * b2, b2, v->n2);
  const int a1 = s1.first, len1 = s1.second, a2 = s2.first, len2 = s2.second;
  auto range = [](int n, int r, int g, int& lo, int& hi) {
    lo = (int)((long long)r * n / g);
    hi = (int)((long long)(r + 1) * n / g);
  };
  Plan p{};
  const int half = v->world / 2;
  if (v->owner_negs) {  // every rank of a KG's half walks the KG's whole slice
    if (v->rank < half) {
      range(len1, v->rank, half, p.lo, p.hi);
      p.o1 = a1;
      p.l1 = len1;
      p.base = 0;
    } else {
      range(len2, v->rank - half, half, p.lo, p.hi);
      p.o2 = a2;
      p.l2 = len2;
      p.base = len1;
    }
    p.mine = p.hi - p.lo;
    return p;
  }
  p.lo = 0;
  p.hi = 0x7fffffff;
  int lo, hi;
  if (v->by_kg) {
    if (v->rank < half) {
      range(len1, v->rank, half, lo, hi);
      p.o1 = a1 + lo;
      p.l1 = hi - lo;
      p.o2 = a2;
      p.base = lo;
    } else {
      range(len2, v->rank - half, half, lo, hi);
      p.o1 = a1 + len1;
      p.o2 = a2 + lo;
      p.l2 = hi - lo;
      p.base = len1 + lo;
    }
  } else {
    range(len1 + len2, v->rank, v->world, lo, hi);
    const int u1 = lo < len1 ? lo : len1, t1 = hi < len1 ? hi : len1;
    const int u2 = (lo > len1 ? lo : len1) - len1, t2 = (hi > len1 ? hi : len1) - len1;
    p.o1 = a1 + u1;
    p.l1 = t1 - u1;
    p.o2 = a2 + u2;
    p.l2 = t2 - u2;
    p.base = lo;
  }
  p.mine = p.l1 + p.l2;
  return p;
}

}  // namespace mke

using namespace mke;

extern "C" int mke_rel_sharded_train_steps(const mke_rel_sharded_view_t* v, int32_t first_step, int32_t n_steps,
                                           uint64_t first_global_step, uint32_t* barrier_seq, int64_t* positives_out,
                                           mke_stream_t main_, mke_stream_t side_) {
  MKE_CHECK_ARG(v && v->ent && v->rel && v->ent_acc && v->rel_acc, "view needs tables and Adagrad slots");
  MKE_CHECK_ARG(v->world == 2 || v->world == 4 || v->world == 8, "world=%d (2, 4 or 8)", v->world);
  MKE_CHECK_ARG(v->rank >= 0 && v->rank < v->world && v->ent->n_shards == v->world, "rank / shard count mismatch");
  MKE_CHECK_ARG(v->n1 >= 0 && v->n2 >= 0 && (long long)v->n1 + v->n2 > 0 && v->global_batch > 0, "bad lists / batch");
  MKE_CHECK_ARG(v->K >= 0 && v->K <= MKE_MAX_NEG && n_steps >= 0 && first_step >= 0, "bad K / step range");
  MKE_CHECK_ARG(barrier_seq && v->step_loss, "barrier_seq / step_loss are null");
  MKE_CHECK_ARG(v->K == 0 || (v->neg_ent[0] && v->neg_ent[1] && v->neg_side[0] && v->neg_side[1]), "negative buffers");
  MKE_CHECK_ARG(!v->owner_negs || (v->neg_valid[0] && v->neg_valid[1] && v->by_kg && v->world >= 4),
                "negatives-where-they-live needs ownership words, KG-block placement and >= 4 ranks");
  MKE_CHECK_ARG(v->rel->stride % 4 == 0, "relation stride");
  for (int k = 0; k < v->world; ++k) MKE_CHECK_ARG(v->xchg[k] && v->sync[k], "peer exchange / flag buffers");
  cudaStream_t main = (cudaStream_t)main_;
  const bool ahead = v->K > 0 && side_ != nullptr && side_ != main_;
  cudaStream_t side = ahead ? (cudaStream_t)side_ : main;
  const int steps_per_epoch = (int)(((long long)v->n1 + v->n2 + v->global_batch - 1) / v->global_batch);
  ExchangeParams x{};
  x.ps.world = v->world;
  x.ps.rank = v->rank;
  for (int k = 0; k < v->world; ++k) {
    x.ps.flags[k] = v->sync[k];
    x.xchg[k] = v->xchg[k];
  }
  x.local = v->sync[v->rank] + MKE_MAX_SHARDS;
  x.grad = v->rel->grad;
  x.replicas = v->rel->grad_replicas > 1 ? v->rel->grad_replicas : 1;
  x.n4 = v->rel->rows * v->rel->stride / 4;
  cudaEvent_t ev_free = nullptr, ev_ready = nullptr;  // buffers of the other parity free / negatives of the next step drawn
  if (ahead) {
    if (cudaEventCreateWithFlags(&ev_free, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ev_ready, cudaEventDisableTiming) != cudaSuccess)
      return cuda_fail(cudaGetLastError(), "cudaEventCreate");
  }
  const bool host_fed = v->host_triples1 != nullptr || v->host_triples2 != nullptr;
  MKE_CHECK_ARG(!host_fed || ((v->n1 == 0 || (v->host_triples1 && v->stage1[0] && v->stage1[1])) &&
                              (v->n2 == 0 || (v->host_triples2 && v->stage2[0] && v->stage2[1]))),
                "host-fed batches need pinned lists and two staging buffers per KG");
  MKE_CHECK_ARG(host_fed || ((v->n1 == 0 || v->triples1) && (v->n2 == 0 || v->triples2)), "device triple lists are null");
  // this rank's positives of a step: pointers into the resident lists, or an H2D copy into staging set `buf`
  auto resolve = [&](Plan& p, int buf, cudaStream_t st) -> int {
    if (!host_fed) {
      p.p1 = v->triples1 ? v->triples1 + 3 * (size_t)p.o1 : nullptr;
      p.p2 = v->triples2 ? v->triples2 + 3 * (size_t)p.o2 : nullptr;
      return 0;
    }
    cudaError_t e = cudaSuccess;
    if (p.l1 > 0)
      e = cudaMemcpyAsync(v->stage1[buf], v->host_triples1 + 3 * (size_t)p.o1, (size_t)p.l1 * 12, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && p.l2 > 0)
      e = cudaMemcpyAsync(v->stage2[buf], v->host_triples2 + 3 * (size_t)p.o2, (size_t)p.l2 * 12, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return cuda_fail(e, "H2D batch");
    p.p1 = v->stage1[buf];
    p.p2 = v->stage2[buf];
    return 0;
  };
  auto draw = [&](Plan& p, uint64_t gstep, int buf, cudaStream_t st) -> int {
    if (int rc0 = resolve(p, buf, st)) return rc0;
    if (v->K == 0 || p.l1 + p.l2 == 0) return 0;
    if (int rc = mke_sample_structured_at(p.p1, p.l1, v->kg1, p.p2, p.l2, v->kg2, v->K, v->seed, gstep, p.base,
                                          v->neg_ent[buf], v->neg_side[buf], st))
      return rc;
    if (v->owner_negs)
      return mke_neg_keep_owned2(v->neg_ent[buf], v->neg_side[buf], p.l1 + p.l2, v->K, v->world, v->ent->shard_split,
                                 v->rank, v->dummy_row, v->neg_valid[buf], st);
    return 0;
  };
  double* host_loss_dev = nullptr;  // the pinned loss buffer as the device sees it
  if (v->host_step_loss != nullptr) {
    void* d = nullptr;
    if (cudaHostGetDevicePointer(&d, (void*)v->host_step_loss, 0) == cudaSuccess)
      host_loss_dev = (double*)d;
    else
      cudaGetLastError();
  }
  long long positives = 0;
  int rc = 0;
  Plan cur = make_plan(v, first_step % steps_per_epoch);
  if (n_steps > 0) rc = draw(cur, first_global_step, 0, main);
  for (int s = 0; s < n_steps && rc == 0; ++s) {
    const int buf = s & 1;
    const int n = cur.l1 + cur.l2;
    Plan nxt{};
    const bool have_next = s + 1 < n_steps;
    if (have_next) nxt = make_plan(v, (first_step + s + 1) % steps_per_epoch);
    if (n > 0) {
      const bool timed = timer_begin(main);
      rc = mke_rel_step_structured4(v->ent, v->rel, cur.p1, cur.l1, cur.p2, cur.l2, v->K, v->neg_ent[buf],
                                    v->neg_side[buf], v->owner_negs ? v->neg_valid[buf] : nullptr, 1, cur.lo, cur.hi,
                                    nullptr, 1.0f, v->step_loss + s, v->variant, main);
      if (timed) timer_end(main);
      if (rc) break;
      if (v->host_step_loss != nullptr && host_loss_dev == nullptr)  // (pinned memory the device cannot address: copy)
        if (cudaError_t e = cudaMemcpyAsync(v->host_step_loss + s, v->step_loss + s, sizeof(double), cudaMemcpyDeviceToHost, main)) {
          rc = cuda_fail(e, "D2H loss");
          break;
        }
    }
    if (have_next) {  // the next step's negatives: under the exchange and phase 2 when there is a side stream
      if (ahead) {
        cudaEventRecord(ev_free, main);  // phase 1 of step s-1, the last reader of the other buffers, precedes this
        cudaStreamWaitEvent(side, ev_free, 0);
      }
      rc = draw(nxt, first_global_step + (uint64_t)(s + 1), buf ^ 1, side);
      if (rc) break;
      if (ahead) cudaEventRecord(ev_ready, side);
    }
    x.seq = *barrier_seq + 1u;     // barrier A; barrier C is seq + 1
    x.launch_no = *barrier_seq / 2u;
    *barrier_seq += 2u;
    {
      void* args[] = {(void*)&x};
      cudaError_t e = cudaLaunchCooperativeKernel((const void*)rel_exchange_kernel, dim3(kXBlocks), dim3(kXThreads), args, 0, main);
      count_launch();
      if (e != cudaSuccess) { rc = cuda_fail(e, "rel_exchange_kernel"); break; }
    }
    rc = mke_rows_apply_adagrad_pair(v->ent, v->ent_acc, v->lr, v->rel, v->rel_acc, v->lr, main);
    if (rc) break;
    // this rank's share of the step loss, back to the host every step: stored by the barrier kernel
    peer_barrier_kernel<<<1, 32, 0, main>>>(x.ps, x.seq + 1u, v->step_loss + s,
                                            (host_loss_dev != nullptr && n > 0) ? host_loss_dev + s : nullptr);
    count_launch();
    if (cudaError_t e = cudaGetLastError()) { rc = cuda_fail(e, "peer_barrier_kernel"); break; }
    if (have_next && ahead) cudaStreamWaitEvent(main, ev_ready, 0);
    positives += cur.mine;
    cur = nxt;
  }
  if (ev_free) cudaEventDestroy(ev_free);
  if (ev_ready) cudaEventDestroy(ev_ready);
  if (rc) return rc;
  if (positives_out != nullptr) *positives_out = positives;
  return 0;
}
