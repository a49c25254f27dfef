// Relation view, PERSISTENT STEP KERNEL: the whole inner loop of train_relation_view_1epo
// (MultiKE_model.py:302-313) -- negatives (base/batch.py:86-116), phase 1 (MultiKE_model.py:123-131,
// losses.py:4-12), phase 2 (MultiKE_model.py:15-31) -- for a run of consecutive steps in ONE
// cooperative launch.  Why: at the reference's batch of 20 000 a step is two 30-45 us kernels, and a
// quarter of each launch is grid ramp-up, drain and tail, plus the gaps between launches and the
// cross-stream fence of the sampler (profiles/r1_phase1_trace.md: 75 us of kernels in an 89-93 us step).
// Here the SMs stay loaded; the step boundaries that Adagrad's non-linearity forces (sum, then apply)
// are grid barriers instead of launches.  One block of W warps per SM (default: all of them workers):
//
//   prologue : negatives of step 0 (tickets of 4 positives)                                 | barrier
//   step s   : phase 1 = row stream of mke_rel_q8p.cuh, static split                       | barrier 1
//              phase 2 = flagged rows as a second cp.async row stream (ApplyStream), by ticket,
//              then the negatives of step s+1 (sampling reads no table)                    | barrier 2
//
// A barrier = bar.sync, one fence + one atomic + ONE poller per block (relaxed loads), one fence
// (MEMBAR.SC.GPU + CCTL.IVALL), bar.sync; about 1.5 us.  That acquire is what makes the rows another SM updated
// in phase 2 visible to this SM's L1-allocating cp.async in the next phase 1.
// Host-fed steps: the batches arrive by cudaMemcpyAsync on another stream while the kernel runs; flags[k] (a
// 4-byte copy issued after the batch copy) tells the kernel that step k has landed, and the step loss is stored
// straight into pinned host memory: the end-to-end rate equals the device-resident one.
//
// Measured on the way (B200, cfg 2, us per step; profiles/r2_persistent_kernel.md):
//   * every warp arriving / polling for itself: fences of 2-7 us each while the SM still has reductions in flight,
//     and +10 us per phase from the load on the counter's L2 slice -> one arrival and one poller per block;
//   * warps specialised on sampling (MKE_PERSIST_SPLIT, kept as an experiment): 2 extra warps cost phase 1 4-5 us
//     even when they sit idle in the hardware barrier, and phase 2 gains nothing;
//   * sampling as the waiting time of barrier 1 (the SMs finish phase 1 between 24 and 35 us, the same SMs slow
//     in every step): -3.7 us on phase 2, +2.4 us on phase 1;
//   * phase 2 as load-compute-store per row: 57.6 us; as a row stream through the ring: 53.8 us; 18 warps per SM,
//     not 16 or 20 (the batch of 20 000 splits worse over their quarters);
//   * the kernel sits exactly at its register budget (18 warps = 5 per scheduler = 96 registers): one more live
//     value in the step loop (a debug stamp) cost 4.5 us per step -- keep the step loop as it is.
#include <cstdlib>
#include "mke_rel_persist.cuh"

namespace mke {

// One block per SM: a grid barrier then costs one fence, one atomic and one poller per SM, and the block's
// warps (18 at stride <= 80: the register file holds 18 x 32 x 112) share one phase schedule.
constexpr unsigned long long kWaitLimitNs = 4000000000ull;  // a bounded wait that runs out is a bug: trap, do not hang

constexpr int ps_max_regs(int warps) {
  const int per_smsp = (warps + 3) / 4;
  const int r = ((16384 / (per_smsp * 32)) / 8) * 8;
  return r > 255 ? 248 : r;
}

__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ void fence_acq_rel() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }

__device__ __noinline__ void wait_failed(uint32_t* sync, uint32_t code) {
  atomicExch(sync + kSyncError, code);
  __threadfence_system();
  __trap();
}

// ---- grid barrier (all blocks of the cooperative, hence co-resident, grid): one arrival per block, every
// warp waits on its own ----------------------------------------------------------------------------------
// The poll of the barrier counter is a RELAXED load (an acquire load is followed by CCTL.IVALL each time);
// the one acquire fence (+ L1 invalidation) comes after the poll succeeded.
__device__ __forceinline__ uint32_t ld_relaxed_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
struct StepSlice {
  const int32_t* pos1;
  int len1;
  const int32_t* pos2;
  int len2;
};
// base/batch.py:36-37, 45-54: step e of an epoch is [e b1, (e+1) b1) of list 1 ++ [e b2, (e+1) b2) of list 2, clipped
__device__ __forceinline__ StepSlice step_slice_dev(const PersistParams& q, int s) {
  const int e = (q.first_step + s) % q.steps_per_epoch;
  auto clip = [](long long start, int bs, int n, int& a, int& len) {
    a = (int)(start < n ? start : n);
    const int end = (int)(start + bs < n ? start + bs : n);
    len = end - a;
  };
  int a1, a2;
  StepSlice sl;
  clip((long long)e * q.b1, q.b1, q.n1, a1, sl.len1);
  clip((long long)e * q.b2, q.b2, q.n2, a2, sl.len2);
  if (q.st1 != nullptr || q.st2 != nullptr) {  // host fed: step s of the launch sits at its own staging offset
    sl.pos1 = q.st1 + 3 * (size_t)s * q.b1;
    sl.pos2 = q.st2 + 3 * (size_t)s * q.b2;
  } else {
    sl.pos1 = q.t1 + 3 * (size_t)a1;
    sl.pos2 = q.t2 + 3 * (size_t)a2;
  }
  return sl;
}

// one ticket of the sampling queue: positives 4 item .. 4 item + 3, one per quarter (same draws as sample_kernel)
__device__ __forceinline__ void sample_item(const RelStepParams& p, const StepSlice& sl, uint64_t skey, int item,
                                            int lane, volatile int32_t* pick, int32_t* __restrict__ neg_ent,
                                            uint32_t* __restrict__ neg_side) {
  const int sub = lane & 7;
  const int i = item * kQPerWarp + (lane >> 3);
  if (i < sl.len1 + sl.len2) {
    const bool first = i < sl.len1;
    const int32_t* row = first ? sl.pos1 + 3 * (size_t)i : sl.pos2 + 3 * (size_t)(i - sl.len1);
    const int32_t h = __ldcg(row), r = __ldcg(row + 1), t = __ldcg(row + 2);
    const KgView kg = kg_view(p, first);
    const uint32_t side = sample_negs_quarter(kg, h, r, t, p.K, skey, (uint32_t)i, lane, pick);
    for (int c = sub; c < p.K; c += 8) neg_ent[(size_t)i * p.K + c] = pick[c];
    if (sub == 0) neg_side[i] = side;
    __syncwarp(0xffu << (lane & 24));  // pick[] is rewritten by this quarter's next positive
  }
}

__device__ __forceinline__ uint32_t take_ticket(uint32_t* ctr, int lane) {
  uint32_t v = 0;
  if (lane == 0) v = atomicAdd(ctr, 1u);
  return __shfl_sync(kFull, v, 0);
}

// one ticket of the apply queue: `chunk` (16 or 32) flag bytes (large tables) or four rows (tables without flags)
template <int FPL>
__device__ __forceinline__ void apply_item(const ApplyTable& T, int item, int chunk, int lane) {
  if (T.touched != nullptr) {
    if (chunk == 16)
      apply_flag_chunk<FPL, 16>(T, item * 16, lane);
    else
      apply_flag_chunk<FPL, 32>(T, item * 32, lane);
  } else {
    apply_row4<FPL>(T, item * 4, lane);
  }
}
__host__ __device__ __forceinline__ int apply_items(const ApplyTable& T, int chunk) {
  if (T.rows <= 0) return 0;
  return T.touched != nullptr ? (T.rows + chunk - 1) / chunk : (T.rows + 3) / 4;
}

// The three phase bodies are separate functions so that each gets its own register allocation (the
// row stream of phase 1 was tuned on its own; inlined next to the sampler it spilt inside the K loop).
template <int FPL, int D>
static __device__ __noinline__ float phase1_rows(const RelStepParams& p, const StepBatch b, const int passes, const int Q,
                                                 const int g, unsigned char* ring_w, int32_t* ids_q, float* rel_grad,
                                                 const int lane) {
  return q8p_stream<FPL, D, false>(p, b, passes, Q, g, ring_w, ids_q, rel_grad, lane);
}

// Phase 2 as a ROW STREAM (large flagged table): the g, v, acc rows of the next round of four flagged rows
// travel through the warp's shared-memory ring (2 x 3 slots, cp.async) while the current round is computed,
// across chunk and ticket boundaries -- twice the bytes in flight of the load-compute-store loop of
// apply_flag_chunk without holding a register for them.  Chunks of `chunk` flag bytes come by ticket;
// the ticket after the next and the flag bytes of the next chunk are requested one chunk ahead, so the chain
// ticket -> flags -> rows never stalls the stream.
template <int FPL>
struct ApplyStream {
  static constexpr int stride = FPL * 8;
  const ApplyTable& T;
  uint32_t* ctr;
  const int item0, n_items;  // tickets item0 .. item0 + n_items - 1 are chunks 0 .. n_items - 1 of T
  const int chunk;           // flag bytes per chunk: 16 or 32
  const int lane;
  uint32_t cur_m = 0;
  int cur_base = 0;
  int nx_item;       // chunk whose flag byte this lane already holds (or >= n_items: none)
  uint8_t nx_flag = 0;
  __device__ __forceinline__ ApplyStream(const ApplyTable& t, uint32_t* c, int i0, int n, int first_item, int ch, int ln)
      : T(t), ctr(c), item0(i0), n_items(n), chunk(ch), lane(ln), nx_item(first_item) {
    load_flag();
  }
  __device__ __forceinline__ void load_flag() {
    const int my = nx_item * chunk + lane;
    nx_flag = (nx_item < n_items && lane < chunk && my < T.rows) ? T.touched[my] : (uint8_t)0;
  }
  // next round of (up to) four flagged rows: this quarter's row offset, or on == false.  Returns false at the end.
  __device__ __forceinline__ bool next(size_t& off, bool& on) {
    while (cur_m == 0u) {
      if (nx_item >= n_items) return false;
      const bool flag = nx_flag != 0;
      cur_m = __ballot_sync(kFullMask, flag);
      cur_base = nx_item * chunk;
      if (flag) T.touched[cur_base + lane] = 0;
      const uint32_t t = take_ticket(ctr, lane);
      nx_item = (int)t - item0;
      load_flag();
    }
    const int q = lane >> 3;
    uint32_t mm = cur_m;
    int bit = -1;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int b = mm ? (__ffs(mm) - 1) : -1;
      if (k == q) bit = b;
      mm &= mm - 1;
    }
    cur_m = mm;
    on = bit >= 0;
    off = (size_t)(cur_base + (on ? bit : 0)) * stride;
    return true;
  }
};

template <int FPL>
static __device__ __noinline__ void phase2_apply(const PersistParams& q, uint32_t* ctr, const int items_b,
                                                 const int items_all, unsigned char* ring_w, const int lane) {
  using Ring = Stage<FPL, 6>;
  const int sub = lane & 7;
  uint32_t it = take_ticket(ctr, lane);
  // the small table first (no flags, gradient replicas): one item = four rows, load-compute-store
  while ((int)it < items_b) {
    apply_item<FPL>(q.B, (int)it, q.apply_chunk, lane);
    it = take_ticket(ctr, lane);
  }
  if ((int)it >= items_all) return;
  const ApplyTable& T = q.A;
  if (T.touched == nullptr || T.replicas > 1 || q.apply_mode == 0) {  // (or not asked for): item by item
    while ((int)it < items_all) {
      const uint32_t nx = take_ticket(ctr, lane);
      apply_item<FPL>(T, (int)it - items_b, q.apply_chunk, lane);
      it = nx;
    }
    return;
  }
  Ring stg;
  stg.base = (uint32_t)__cvta_generic_to_shared(ring_w);
  stg.lane = lane;
  ApplyStream<FPL> st(T, ctr, items_b, items_all - items_b, (int)it - items_b, q.apply_chunk, lane);
  auto issue = [&](int set, size_t off) {
    stg.issue(3 * set + 0, T.grad + off, sub);
    stg.issue(3 * set + 1, T.var + off, sub);
    stg.issue(3 * set + 2, T.acc + off, sub);
    cp_async_commit();
  };
  size_t off_a, off_b = 0;
  bool on_a, on_b = false;
  if (!st.next(off_a, on_a)) return;
  issue(0, off_a);
  int set = 0;
  while (true) {
    const bool more = st.next(off_b, on_b);
    if (more) {
      issue(set ^ 1, off_b);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    float g[FPL], v[FPL], a[FPL];
    stg.read(3 * set + 0, g);
    stg.read(3 * set + 1, v);
    stg.read(3 * set + 2, a);
    float inv = 1.f, coef = 0.f;
    if (T.normalised) {
      float ss = 0.f, vg = 0.f;
#pragma unroll
      for (int k = 0; k < FPL; ++k) {
        ss = fmaf(v[k], v[k], ss);
        vg = fmaf(v[k], g[k], vg);
      }
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) {
        ss += __shfl_xor_sync(kFullMask, ss, o);
        vg += __shfl_xor_sync(kFullMask, vg, o);
      }
      // y = v * rsqrt(max(|v|^2, eps)); the max() routes no gradient to |v|^2 below eps (as apply_one_row)
      inv = rsqrtf(fmaxf(ss, kNormEps));
      coef = (ss >= kNormEps) ? vg * inv * inv : 0.f;
    }
#pragma unroll
    for (int k = 0; k < FPL; ++k) {
      const float gv = (g[k] - v[k] * coef) * inv;
      a[k] = fmaf(gv, gv, a[k]);
      v[k] -= gv * T.lr * (a[k] > 0.f ? rsqrtf(a[k]) : 0.f);  // ApplyAdagrad, no epsilon [TF semantics]
      g[k] = 0.f;
    }
    if (on_a) {
      q_store<FPL>(T.var + off_a, sub, v);
      q_store<FPL>(T.acc + off_a, sub, a);
      q_store<FPL>(T.grad + off_a, sub, g);
    }
    if (!more) break;
    off_a = off_b;
    on_a = on_b;
    set ^= 1;
  }
}

// Negatives of step s of the launch (tickets of 4 positives), after its batch has landed (host fed): drains the queue.
static __device__ __noinline__ void sample_queue(const RelStepParams& p, const PersistParams& q, const int s,
                                                 const StepSlice sl, uint32_t* ctr, volatile int32_t* pick,
                                                 const int lane) {
  const int items = (sl.len1 + sl.len2 + kQPerWarp - 1) / kQPerWarp;
  if (items <= 0) return;
  uint32_t cur = take_ticket(ctr, lane);
  if ((int)cur >= items) return;
  if (q.flags != nullptr) {
    if (lane == 0) {
      const unsigned long long t0 = gtimer_raw();
      while (ld_acquire_sys_u32(q.flags + s) != q.flag_value) {
        __nanosleep(200);
        if (gtimer_raw() - t0 > kWaitLimitNs) wait_failed(q.sync, 2u);
      }
    }
    __syncwarp();
  }
  const uint64_t skey = stream_key(q.seed, q.first_global_step + (uint64_t)s);
  while ((int)cur < items) {
    const uint32_t nxt = take_ticket(ctr, lane);
    sample_item(p, sl, skey, (int)cur, lane, pick, q.neg_ent[s & 1], q.neg_side[s & 1]);
    cur = nxt;
  }
}

template <int FPL, int D, int W, int S>
__global__ void __launch_bounds__((W + S) * 32, 1) __maxnreg__(ps_max_regs(W + S))
    rel_step_persist_kernel(const __grid_constant__ RelStepParams p, const __grid_constant__ PersistParams q) {
  using Ring = Stage<FPL, (D > 6 ? D : 6)>;  // phase 1 uses D slots of it, phase 2 two sets of three
  extern __shared__ __align__(128) unsigned char s_dyn[];
  unsigned char* const s_ring = s_dyn;                                           // [W][Ring::kBytes]
  int32_t* const s_ids = reinterpret_cast<int32_t*>(s_dyn + W * Ring::kBytes);  // [W + S][4][2][kIdStride]
  float* const s_loss = reinterpret_cast<float*>(s_ids + (W + S) * kQPerWarp * 2 * kIdStride);  // [W]
  const int lane = threadIdx.x & 31;
  const int sub = lane & 7;
  const int qi = lane >> 3;
  const int wib = threadIdx.x >> 5;
  const bool sampler = wib >= W;
  const int gw = blockIdx.x * W + wib;  // workers only
  const int Q = gridDim.x * W * kQPerWarp;
  const int g = gw * kQPerWarp + qi;
  int32_t* const ids_q = s_ids + (wib * kQPerWarp + qi) * 2 * kIdStride;
  volatile int32_t* const pick = ids_q;  // sampler scratch (a worker samples only while its row stream is idle)
  const bool leader = blockIdx.x == 0 && threadIdx.x == 0;
  const bool tracer = threadIdx.x == 0 && q.block_trace != nullptr;  // debug: thread 0 of every block
  uint32_t round = 0;  // barriers completed
  auto bt = [&](int k) -> unsigned long long& { return q.block_trace[((size_t)round * gridDim.x + blockIdx.x) * 4 + k]; };
  // One arrival and ONE poller per block (thread 0), between two bar.syncs of the warps that take part (all of
  // them, or the workers): waiting warps sit in the hardware barrier and put no load on the L2 slice of the
  // counter (every warp polling for itself was measured: +10 us per phase).
  auto barrier = [&](bool workers_only) {
    auto block_sync = [&]() {
      if (workers_only)
        asm volatile("bar.sync 1, %0;" ::"n"(W * 32) : "memory");
      else
        __syncthreads();
    };
    block_sync();
    if (threadIdx.x == 0) {
      if (tracer) bt(0) = gtimer_raw();
      if (q.fence_mode) fence_acq_rel(); else __threadfence();  // release: the block's writes and reductions (ordered before by bar.sync)
      atomicAdd(q.sync + kSyncBarrier, 1u);
      if (tracer) bt(1) = gtimer_raw();
      const uint32_t target = (round + 1) * gridDim.x;
      const unsigned long long t0 = gtimer_raw();
      while (ld_relaxed_u32(q.sync + kSyncBarrier) < target) {
        __nanosleep(100);
        if (gtimer_raw() - t0 > kWaitLimitNs) wait_failed(q.sync, 1u);
      }
      if (q.fence_mode) fence_acq_rel(); else __threadfence();  // acquire + L1 invalidation for the whole SM (MEMBAR + CCTL.IVALL)
      if (tracer) bt(3) = gtimer_raw();
    }
    block_sync();
    ++round;
    if (leader && q.trace != nullptr) q.trace[round] = gtimer_raw();
  };
  if (leader && q.trace != nullptr) q.trace[0] = gtimer_raw();
  const StepSlice none{nullptr, 0, nullptr, 0};

  // ---- prologue: negatives of the first step, by everybody -------------------------------------------
  StepSlice cur = step_slice_dev(q, 0);
  sample_queue(p, q, 0, cur, q.sync + kSyncQueues + 1, pick, lane);
  barrier(false);

  if (sampler) {
#pragma unroll 1
    for (int s = 0; s < q.n_steps; ++s) {
      // samp_phase 1: sample only under phase 2 (idle in the hardware barrier during phase 1, where extra
      // warps were measured to cost the row stream 5 us); 0: all through the step
      if (q.samp_phase == 1) barrier(false);
      if (s + 1 < q.n_steps)
        sample_queue(p, q, s + 1, step_slice_dev(q, s + 1), q.sync + kSyncQueues + 2 * (s + 1) + 1, pick, lane);
      if (q.samp_phase != 1) ++round;  // barrier 1 is the workers'
      barrier(false);
    }
    return;
  }

  unsigned char* const ring_w = s_ring + wib * Ring::kBytes;
  float* const rel_grad = rel_grad_replica(p);
  const int items_b = apply_items(q.B, q.apply_chunk), items_all = items_b + apply_items(q.A, q.apply_chunk);
#pragma unroll 1
  for (int s = 0; s < q.n_steps; ++s) {
    const int n = cur.len1 + cur.len2;
    const bool has_next = s + 1 < q.n_steps;
    const StepSlice nxt = has_next ? step_slice_dev(q, s + 1) : none;
    // ---- phase 1 ---------------------------------------------------------------------------------
    if (n > 0) {
      const int passes = (n + Q - 1) / Q;
      const StepBatch b{cur.pos1, cur.len1, cur.pos2, cur.len2, q.neg_ent[s & 1], q.neg_side[s & 1]};
      const float loss_local = phase1_rows<FPL, D>(p, b, passes, Q, g, ring_w, ids_q, rel_grad, lane);
      float v = (sub == 0) ? loss_local : 0.f;
      v = warp_sum(v);
      if (lane == 0) s_loss[wib] = v;
      asm volatile("bar.sync 1, %0;" ::"n"(W * 32) : "memory");
      if (threadIdx.x == 0) {  // one fp64 atomic per block and step
        double a = 0.0;
#pragma unroll 1
        for (int w = 0; w < W; ++w) a += (double)s_loss[w];
        if (a != 0.0) atomicAdd(q.step_loss + s, a);
      }
    }
    barrier(S == 0 || q.samp_phase != 1);
    if (leader && q.host_loss != nullptr && n > 0) {
      const double v = __ldcg(q.step_loss + s);
      *reinterpret_cast<volatile double*>(q.host_loss + s) = v;
    }
    // ---- phase 2, then whatever the sampler warps have left of the next step's negatives --------------
    if (n > 0) phase2_apply<FPL>(q, q.sync + kSyncQueues + 2 * s, items_b, items_all, ring_w, lane);
    if (has_next) sample_queue(p, q, s + 1, nxt, q.sync + kSyncQueues + 2 * (s + 1) + 1, pick, lane);
    barrier(false);
    cur = nxt;
  }
}

template <int FPL, int D, int W, int S>
static int launch_persist(const RelStepParams& p, const PersistParams& q, cudaStream_t stream) {
  auto kern = rel_step_persist_kernel<FPL, D, W, S>;
  using Ring = Stage<FPL, (D > 6 ? D : 6)>;
  constexpr size_t smem = (size_t)W * Ring::kBytes + (size_t)(W + S) * kQPerWarp * 2 * kIdStride * sizeof(int32_t) +
                          W * sizeof(float);
  static bool configured = false;
  if (!configured) {
    if (cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem))
      return cuda_fail(e, "cudaFuncSetAttribute(rel_step_persist_kernel)");
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, (W + S) * 32, smem) != cudaSuccess || per_sm < 1) {
      set_error("rel_step_persist_kernel does not fit an SM (%zu bytes of shared memory)", smem);
      return MKE_EINVAL;
    }
    configured = true;
  }
  int blocks = sm_count();
  if (const char* fg = getenv("MKE_PERSIST_GRID")) {  // test knob: tiny grids make small fixtures run many passes
    const int v = atoi(fg);
    if (v > 0 && v < blocks) blocks = v;
  }
  void* args[] = {(void*)&p, (void*)&q};
  cudaError_t e = cudaLaunchCooperativeKernel((const void*)kern, dim3(blocks), dim3((W + S) * 32), args, smem, stream);
  count_launch();
  if (e != cudaSuccess) return cuda_fail(e, "rel_step_persist_kernel");
  return 0;
}

// workers + samplers per block: up to 20 warps at stride <= 80 (5 per scheduler: 96 registers each), 16 above.
// MKE_PERSIST_SPLIT = 10 W' + S picks W = WARPS - W' workers and S samplers (experiments; default below).
template <int FPL, int D, int WARPS>
static int launch_persist_split(const RelStepParams& p, const PersistParams& q, cudaStream_t stream) {
  static const int split = getenv("MKE_PERSIST_SPLIT") ? atoi(getenv("MKE_PERSIST_SPLIT")) : 20;
  switch (split) {
    case 0: return launch_persist<FPL, D, WARPS, 0>(p, q, stream);
    case 2: return launch_persist<FPL, D, WARPS - 2, 2>(p, q, stream);
    case 22: return launch_persist<FPL, D, WARPS - 4, 2>(p, q, stream);
    case 40: return launch_persist<FPL, D, WARPS - 4, 0>(p, q, stream);
    default: return launch_persist<FPL, D, WARPS - 2, 0>(p, q, stream);
  }
}

int launch_rel_persist(const RelStepParams& p, const PersistParams& q, cudaStream_t stream) {
  if (p.sharded || p.K > MKE_MAX_NEG || p.K < 1 || p.w != nullptr || p.neg_valid != nullptr) return 1;
  const int R = 3 + p.K;
  if (2 * 4 > R) return 1;
  const bool deep = 2 * 6 <= R;
  switch (p.stride) {
#define MKE_PS_CASE(STRIDE, FPL, WARPS)                                 \
  case STRIDE:                                                          \
    if (deep) return launch_persist_split<FPL, 6, WARPS>(p, q, stream); \
    return launch_persist_split<FPL, 4, WARPS>(p, q, stream);
    MKE_PS_CASE(32, 4, 20)
    MKE_PS_CASE(64, 8, 20)
    MKE_PS_CASE(80, 10, 20)
    MKE_PS_CASE(104, 13, 16)
    MKE_PS_CASE(128, 16, 16)
#undef MKE_PS_CASE
    default: return 1;
  }
}

}  // namespace mke
