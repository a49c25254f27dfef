// Relation view, PERSISTENT STEP KERNEL: the whole inner loop of train_relation_view_1epo
// (MultiKE_model.py:302-313) -- negatives (base/batch.py:86-116), phase 1 (MultiKE_model.py:123-131,
// losses.py:4-12), phase 2 (MultiKE_model.py:15-31) -- for a run of consecutive steps in ONE
// cooperative launch.  Why: at the reference's batch of 20 000 a step is two 30-45 us kernels, and a
// quarter of each launch is grid ramp-up, drain and tail, plus the gaps between launches and the
// cross-stream fence of the sampler (profiles/r1_phase1_trace.md: 75 us of kernels in an 89-93 us step).
// Here the SMs stay loaded; the step boundaries that Adagrad's non-linearity forces (sum, then apply)
// are grid barriers (one atomic + one acquire poll per block, ~1 us) instead of launches:
//
//   prologue : negatives of step 0                                             | barrier
//   step s   : phase 1 (row stream, mke_rel_q8p.cuh)                           | barrier
//              phase 2 (flagged rows, mke_apply.cuh) || negatives of step s+1  | barrier
//
// Phase 2 is HBM-bound and sampling is latency-bound with a few MB of traffic, so they share the
// SMs: two work queues (atomic tickets); every samp_mod-th warp starts on the sampling queue and
// moves to the apply queue when it is empty, the other warps the other way round.  The barrier's
// fence (MEMBAR.SC.GPU + CCTL.IVALL) is what makes the rows another SM updated in phase 2 visible
// to this SM's L1-allocating cp.async in the next phase 1.
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

// All blocks of the (cooperative, hence co-resident) grid.  `round` counts this block's barriers.
// bt (debug, may be NULL): four stamps of this block -- arrived, fenced, released, done.
__device__ __forceinline__ void grid_barrier(uint32_t* sync, uint32_t& round, unsigned long long* bt) {
  __syncthreads();
  if (threadIdx.x == 0) {
    ++round;
    const uint32_t target = round * gridDim.x;
    if (bt) bt[0] = gtimer_raw();
    __threadfence();  // release: this block's writes and reductions (ordered before by bar.sync)
    if (bt) bt[1] = gtimer_raw();
    atomicAdd(sync + kSyncBarrier, 1u);
    const unsigned long long t0 = gtimer_raw();
    while (ld_acquire_u32(sync + kSyncBarrier) < target) {
      if (gtimer_raw() - t0 > kWaitLimitNs) wait_failed(sync, 1u);
    }
    if (bt) bt[2] = gtimer_raw();
    __threadfence();  // acquire + L1 invalidation for the whole SM
    if (bt) bt[3] = gtimer_raw();
  }
  __syncthreads();
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

template <int FPL>
__device__ __forceinline__ void apply_item(const ApplyTable& T, int item, int lane) {
  if (T.touched != nullptr)
    apply_flag_chunk<FPL>(T, item * 32, lane);
  else
    apply_row4<FPL>(T, item * 4, lane);
}
__host__ __device__ __forceinline__ int apply_items(const ApplyTable& T) {
  if (T.rows <= 0) return 0;
  return T.touched != nullptr ? (T.rows + 31) / 32 : (T.rows + 3) / 4;
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

// negatives of step s of the launch (tickets of 4 positives), after its batch has landed (host fed)
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

template <int FPL, int D, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) __maxnreg__(ps_max_regs(WARPS))
    rel_step_persist_kernel(const __grid_constant__ RelStepParams p, const __grid_constant__ PersistParams q) {
  using Ring = Stage<FPL, D>;
  extern __shared__ __align__(128) unsigned char s_dyn[];
  unsigned char* const s_ring = s_dyn;                                               // [WARPS][Ring::kBytes]
  int32_t* const s_ids = reinterpret_cast<int32_t*>(s_dyn + WARPS * Ring::kBytes);  // [WARPS][4][2][kIdStride]
  float* const s_loss = reinterpret_cast<float*>(s_ids + WARPS * kQPerWarp * 2 * kIdStride);  // [WARPS]
  const int lane = threadIdx.x & 31;
  const int sub = lane & 7;
  const int qi = lane >> 3;
  const int wib = threadIdx.x >> 5;
  const int gw = blockIdx.x * WARPS + wib;
  const int Q = gridDim.x * WARPS * kQPerWarp;
  const int g = gw * kQPerWarp + qi;
  unsigned char* const ring_w = s_ring + wib * Ring::kBytes;
  int32_t* const ids_q = s_ids + (wib * kQPerWarp + qi) * 2 * kIdStride;
  volatile int32_t* const pick = ids_q;  // sampler scratch (phase 1 is not running then)
  float* const rel_grad = rel_grad_replica(p);
  const bool leader = blockIdx.x == 0 && threadIdx.x == 0;
  uint32_t round = 0;
  int stamp = 0;
  auto barrier = [&]() {
    unsigned long long* bt =
        q.block_trace ? q.block_trace + ((size_t)stamp * gridDim.x + blockIdx.x) * 4 : nullptr;
    grid_barrier(q.sync, round, bt);
    ++stamp;
    if (leader && q.trace != nullptr) q.trace[stamp] = gtimer_raw();
  };
  if (leader && q.trace != nullptr) q.trace[0] = gtimer_raw();

  // ---- prologue: negatives of the first step -------------------------------------------------------
  StepSlice cur = step_slice_dev(q, 0);
  sample_queue(p, q, 0, cur, q.sync + kSyncPrologue, pick, lane);
  barrier();

  const int items_b = apply_items(q.B), items_all = items_b + apply_items(q.A);
#pragma unroll 1
  for (int s = 0; s < q.n_steps; ++s) {
    const int n = cur.len1 + cur.len2;
    // ---- phase 1 ---------------------------------------------------------------------------------
    if (leader) {  // the queues of the step after this one (last used two barriers ago)
      q.sync[kSyncQueues + 2 * ((s + 1) & 1)] = 0u;
      q.sync[kSyncQueues + 2 * ((s + 1) & 1) + 1] = 0u;
    }
    if (n > 0) {
      const int passes = (n + Q - 1) / Q;
      const StepBatch b{cur.pos1, cur.len1, cur.pos2, cur.len2, q.neg_ent[s & 1], q.neg_side[s & 1]};
      const float loss_local = phase1_rows<FPL, D>(p, b, passes, Q, g, ring_w, ids_q, rel_grad, lane);
      float v = (sub == 0) ? loss_local : 0.f;
      v = warp_sum(v);
      if (lane == 0) s_loss[wib] = v;
      __syncthreads();
      if (threadIdx.x == 0) {
        double a = 0.0;
#pragma unroll 1
        for (int w = 0; w < WARPS; ++w) a += (double)s_loss[w];
        if (a != 0.0) atomicAdd(q.step_loss + s, a);
      }
    }
    barrier();
    if (leader && q.host_loss != nullptr && n > 0) {
      const double v = __ldcg(q.step_loss + s);
      *reinterpret_cast<volatile double*>(q.host_loss + s) = v;
    }
    // ---- phase 2 || negatives of the next step -----------------------------------------------------
    const bool has_next = s + 1 < q.n_steps;
    StepSlice nxt{nullptr, 0, nullptr, 0};
    if (has_next) nxt = step_slice_dev(q, s + 1);
    uint32_t* const ctr_apply = q.sync + kSyncQueues + 2 * (s & 1);
    uint32_t* const ctr_samp = ctr_apply + 1;
    const bool sampler_first = (gw % q.samp_mod) == 0;
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
      if ((half == 0) == sampler_first) {
        if (has_next) sample_queue(p, q, s + 1, nxt, ctr_samp, pick, lane);
      } else if (n > 0) {
        phase2_apply<FPL>(q, ctr_apply, items_b, items_all, lane);
      }
    }
    barrier();
    cur = nxt;
  }
}

template <int FPL, int D, int WARPS>
static int launch_persist(const RelStepParams& p, const PersistParams& q, cudaStream_t stream) {
  auto kern = rel_step_persist_kernel<FPL, D, WARPS>;
  using Ring = Stage<FPL, D>;
  constexpr size_t smem = (size_t)WARPS * Ring::kBytes + (size_t)WARPS * kQPerWarp * 2 * kIdStride * sizeof(int32_t) +
                          WARPS * sizeof(float);
  static bool configured = false;
  if (!configured) {
    if (cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem))
      return cuda_fail(e, "cudaFuncSetAttribute(rel_step_persist_kernel)");
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, WARPS * 32, smem) != cudaSuccess || per_sm < 1) {
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
  cudaError_t e = cudaLaunchCooperativeKernel((const void*)kern, dim3(blocks), dim3(WARPS * 32), args, smem, stream);
  count_launch();
  if (e != cudaSuccess) return cuda_fail(e, "rel_step_persist_kernel");
  return 0;
}

int launch_rel_persist(const RelStepParams& p, const PersistParams& q, cudaStream_t stream) {
  if (p.sharded || p.K > MKE_MAX_NEG || p.K < 1 || p.w != nullptr || p.neg_valid != nullptr) return 1;
  const int R = 3 + p.K;
  if (2 * 4 > R) return 1;
  const bool deep = 2 * 6 <= R;
  switch (p.stride) {
#define MKE_PS_CASE(STRIDE, FPL, WARPS)                                 \
  case STRIDE:                                                          \
    if (deep) return launch_persist<FPL, 6, WARPS>(p, q, stream);       \
    return launch_persist<FPL, 4, WARPS>(p, q, stream);
    MKE_PS_CASE(32, 4, 18)
    MKE_PS_CASE(64, 8, 18)
    MKE_PS_CASE(80, 10, 18)
    MKE_PS_CASE(104, 13, 15)
    MKE_PS_CASE(128, 16, 15)
#undef MKE_PS_CASE
    default: return 1;
  }
}

}  // namespace mke
