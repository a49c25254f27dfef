// Relation view, PERSISTENT STEP KERNEL: the whole inner loop of train_relation_view_1epo
// (MultiKE_model.py:302-313) -- negatives (base/batch.py:86-116), phase 1 (MultiKE_model.py:123-131,
// losses.py:4-12), phase 2 (MultiKE_model.py:15-31) -- for a run of consecutive steps in ONE
// cooperative launch.  Why: at the reference's batch of 20 000 a step is two 30-45 us kernels, and a
// quarter of each launch is grid ramp-up, drain and tail, plus the gaps between launches and the
// cross-stream fence of the sampler (profiles/r1_phase1_trace.md: 75 us of kernels in an 89-93 us step).
// Here the SMs stay loaded; the step boundaries that Adagrad's non-linearity forces (sum, then apply)
// are grid barriers instead of launches.  One block per SM, WARP-SPECIALISED:
//
//   W worker warps : phase 1 of step s (row stream, mke_rel_q8p.cuh) | barrier 1 | phase 2 of step s
//                    (flagged rows by ticket, mke_apply.cuh), leftovers of the sampling queue | barrier 2
//   S sampler warps: negatives of step s+1 (tickets of 4 positives) all through step s        | barrier 2
//
// Sampling reads only the triple lists, the filter set and the counter-based RNG -- never a table -- so it
// needs no ordering against the phases except its buffers: the negatives of step s+1 go to buffer (s+1)&1,
// last read by phase 1 of step s-1, and must be complete when phase 1 of step s+1 starts (barrier 2 of step s,
// at which a block arrives only when the sampling queue is drained).  It is latency-bound work with a few MB of
// traffic; giving it warps of its own (instead of time slices of the workers: measured 58-66 us for phase 2
// + sampling against 45 us for phase 2 alone) lets it hide under both phases.
// A barrier = bar.sync of the participating warps, one fence + one atomic per block, then every warp polls
// with RELAXED loads and fences once (MEMBAR.SC.GPU + CCTL.IVALL): that acquire is what makes the rows another
// SM updated in phase 2 visible to this SM's L1-allocating cp.async in the next phase 1.
// Host-fed steps: the batches arrive by cudaMemcpyAsync on another stream while the kernel runs;
// flags[k] (a 4-byte copy issued after the batch copy) tells the kernel that step k has landed, and
// the step loss is stored straight into pinned host memory.
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

// one ticket of the apply queue: kApplyChunk flag bytes (large tables) or four rows (tables without flags)
constexpr int kApplyChunk = 16;
template <int FPL>
__device__ __forceinline__ void apply_item(const ApplyTable& T, int item, int lane) {
  if (T.touched != nullptr)
    apply_flag_chunk<FPL, kApplyChunk>(T, item * kApplyChunk, lane);
  else
    apply_row4<FPL>(T, item * 4, lane);
}
__host__ __device__ __forceinline__ int apply_items(const ApplyTable& T) {
  if (T.rows <= 0) return 0;
  return T.touched != nullptr ? (T.rows + kApplyChunk - 1) / kApplyChunk : (T.rows + 3) / 4;
}

// The three phase bodies are separate functions so that each gets its own register allocation (the
// row stream of phase 1 was tuned on its own; inlined next to the sampler it spilt inside the K loop).
template <int FPL, int D>
static __device__ __noinline__ float phase1_rows(const RelStepParams& p, const StepBatch b, const int passes, const int Q,
                                                 const int g, unsigned char* ring_w, int32_t* ids_q, float* rel_grad,
                                                 const int lane) {
  return q8p_stream<FPL, D, false>(p, b, passes, Q, g, ring_w, ids_q, rel_grad, lane);
}

template <int FPL>
static __device__ __noinline__ void phase2_apply(const PersistParams& q, uint32_t* ctr, const int items_b,
                                                 const int items_all, const int lane) {
  uint32_t it = take_ticket(ctr, lane);
  while ((int)it < items_all) {
    const uint32_t nx = take_ticket(ctr, lane);
    if ((int)it < items_b)
      apply_item<FPL>(q.B, (int)it, lane);
    else
      apply_item<FPL>(q.A, (int)it - items_b, lane);
    it = nx;
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
  using Ring = Stage<FPL, D>;
  extern __shared__ __align__(128) unsigned char s_dyn[];
  unsigned char* const s_ring = s_dyn;                                           // [W][Ring::kBytes]
  int32_t* const s_ids = reinterpret_cast<int32_t*>(s_dyn + W * Ring::kBytes);  // [W + S][4][2][kIdStride]
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
      __threadfence();  // release: the block's writes and reductions (ordered before by bar.sync)
      atomicAdd(q.sync + kSyncBarrier, 1u);
      if (tracer) bt(1) = gtimer_raw();
      const uint32_t target = (round + 1) * gridDim.x;
      const unsigned long long t0 = gtimer_raw();
      while (ld_relaxed_u32(q.sync + kSyncBarrier) < target) {
        __nanosleep(100);
        if (gtimer_raw() - t0 > kWaitLimitNs) wait_failed(q.sync, 1u);
      }
      if (tracer) bt(2) = gtimer_raw();
      __threadfence();  // acquire + L1 invalidation for the whole SM
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
      if (s + 1 < q.n_steps)
        sample_queue(p, q, s + 1, step_slice_dev(q, s + 1), q.sync + kSyncQueues + 2 * (s + 1) + 1, pick, lane);
      ++round;  // barrier 1 is the workers'
      barrier(false);
    }
    return;
  }

  unsigned char* const ring_w = s_ring + wib * Ring::kBytes;
  float* const rel_grad = rel_grad_replica(p);
  const int items_b = apply_items(q.B), items_all = items_b + apply_items(q.A);
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
      if (lane == 0 && v != 0.f) atomicAdd(q.step_loss + s, (double)v);
    }
    barrier(true);
    if (leader && q.host_loss != nullptr && n > 0) {
      const double v = __ldcg(q.step_loss + s);
      *reinterpret_cast<volatile double*>(q.host_loss + s) = v;
    }
    // ---- phase 2, then whatever the sampler warps have left of the next step's negatives --------------
    if (n > 0) phase2_apply<FPL>(q, q.sync + kSyncQueues + 2 * s, items_b, items_all, lane);
    if (has_next) sample_queue(p, q, s + 1, nxt, q.sync + kSyncQueues + 2 * (s + 1) + 1, pick, lane);
    barrier(false);
    cur = nxt;
  }
}

template <int FPL, int D, int W, int S>
static int launch_persist(const RelStepParams& p, const PersistParams& q, cudaStream_t stream) {
  auto kern = rel_step_persist_kernel<FPL, D, W, S>;
  using Ring = Stage<FPL, D>;
  constexpr size_t smem = (size_t)W * Ring::kBytes + (size_t)(W + S) * kQPerWarp * 2 * kIdStride * sizeof(int32_t);
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

// workers + samplers per block: 20 warps at stride <= 80 (5 per scheduler: 96 registers each), 16 above
template <int FPL, int D, int WARPS>
static int launch_persist_split(const RelStepParams& p, const PersistParams& q, cudaStream_t stream) {
  static const int samplers = getenv("MKE_PERSIST_SAMPLERS") ? atoi(getenv("MKE_PERSIST_SAMPLERS")) : 2;
  switch (samplers) {
    case 1: return launch_persist<FPL, D, WARPS - 1, 1>(p, q, stream);
    case 3: return launch_persist<FPL, D, WARPS - 3, 3>(p, q, stream);
    case 4: return launch_persist<FPL, D, WARPS - 4, 4>(p, q, stream);
    default: return launch_persist<FPL, D, WARPS - 2, 2>(p, q, stream);
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
