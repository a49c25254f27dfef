// Relation view, phase 1, quarter-warp layout, PERSISTENT ROW-STREAM schedule.
//
// Same arithmetic and per-lane row layout as rel_fused_q8_kernel (mke_rel_q8.cu; losses.py:4-12 on
// l2-normalised rows), different schedule.  The one-wave kernel leaves the memory system idle
// while every warp walks the dependent chain ids -> rows -> positive term at the same time, and
// again while the slowest warps finish (profiles/r1_phase1_trace.md).  Here a quarter warp owns
// SEVERAL positives and sees them as one continuous stream of rows
//     h r t e0 .. eK-1 | h r t e0 .. eK-1 | ...
// pulled through a ring of D shared-memory slots with cp.async: while row s is scored, rows
// s+1 .. s+D-1 are in flight, across positive boundaries.  The ids of the next positive
// (3 + K + side word) are themselves fetched with 4-byte cp.async into a double-buffered id list,
// riding in the commit group of an earlier row, so no register waits on them either.  The
// prologue/epilogue bubbles of a positive are filled by its neighbours in the stream, and the grid
// is sized so that every quarter gets the same number of positives (no second-wave tail).
//
// Used for pre-drawn negatives (mke_rel_step_structured*, the pipelined step driver) when
// 2 D <= 3 + K; everything else (fused sampler, K = 0, neg_out) stays on rel_fused_q8_kernel.
#include <cstdlib>
#include "mke_rel_q8p.cuh"

namespace mke {

constexpr int kQ8pThreads = 96;
constexpr int kQ8pWarps = kQ8pThreads / 32;

// register cap for MINB resident blocks: each of the 4 SM sub-partitions owns 16 384 registers and
// gets ceil(3 MINB / 4) of the warps
constexpr int q8p_max_regs(int minb) {
  const int per_smsp = (kQ8pWarps * minb + 3) / 4;
  const int r = ((16384 / (per_smsp * 32)) / 8) * 8;
  return r > 255 ? 248 : r;
}

template <int FPL, int D, int MINB, bool SHARDED>
__global__ void __launch_bounds__(kQ8pThreads) __maxnreg__(q8p_max_regs(MINB)) rel_fused_q8p_kernel(const RelStepParams p, const int passes) {
  constexpr int WARPS = kQ8pWarps;
  using Ring = Stage<FPL, D>;
  __shared__ __align__(128) unsigned char s_ring[WARPS][Ring::kBytes];
  __shared__ int32_t s_ids[WARPS][kQPerWarp][2][kIdStride];
  __shared__ float s_loss[WARPS];
  const int lane = threadIdx.x & 31;
  const int sub = lane & 7;
  const int q = lane >> 3;
  const int wib = threadIdx.x >> 5;
  const int Q = gridDim.x * WARPS * kQPerWarp;
  const int g = (blockIdx.x * WARPS + wib) * kQPerWarp + q;
  const StepBatch b{p.pos1, p.len1, p.pos2, p.len2, p.neg_ent, p.neg_side};
  const float loss_local = q8p_stream<FPL, D, SHARDED>(p, b, passes, Q, g, &s_ring[wib][0], &s_ids[wib][q][0][0],
                                                       rel_grad_replica(p), lane);
  // ---- loss: quarter leaders -> warp -> block -> one fp64 atomic ------------------------------
  float v = (sub == 0) ? loss_local : 0.f;
  v = warp_sum(v);
  if (lane == 0) s_loss[wib] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) a += (double)s_loss[w];
    if (a != 0.0) atomicAdd(p.loss, a);
  }
}

// Grid: b blocks per SM with b chosen so that the quarters (12 per block) divide the batch into
// whole passes as evenly as possible; ties go to the larger b (more rows in flight).
template <int FPL, int D, int MINB, bool SHARDED>
static int launch_q8p(const RelStepParams& p, cudaStream_t stream) {
  auto kern = rel_fused_q8p_kernel<FPL, D, MINB, SHARDED>;
  constexpr int per_block = kQ8pWarps * kQPerWarp;
  static int per_sm_cached = 0;
  if (per_sm_cached == 0) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kQ8pThreads, 0) != cudaSuccess || per_sm < 1)
      per_sm = 1;
    per_sm_cached = per_sm;
  }
  static const int forced_b = getenv("MKE_Q8P_BLOCKS") ? atoi(getenv("MKE_Q8P_BLOCKS")) : 0;
  const int n = p.len1 + p.len2;
  const int need = (n + per_block - 1) / per_block;
  int best_b = per_sm_cached;
  if (forced_b > 0) {
    best_b = forced_b < per_sm_cached ? forced_b : per_sm_cached;
  } else {
    double best_u = -1.0;
    for (int b = 1; b <= per_sm_cached; ++b) {
      const long long Q = (long long)sm_count() * b * per_block;
      const long long passes = (n + Q - 1) / Q;
      const double u = (double)n / (double)(passes * Q);
      if (u >= best_u - 1e-9) {
        best_u = u > best_u ? u : best_u;
        best_b = b;
      }
    }
  }
  int blocks = sm_count() * best_b;
  if (need < blocks) blocks = need;
  // test knob: a tiny grid makes small fixtures run many passes per quarter (read per launch)
  if (const char* forced_grid = getenv("MKE_Q8P_GRID")) {
    const int fg = atoi(forced_grid);
    if (fg > 0 && fg < blocks) blocks = fg;
  }
  const long long Q = (long long)blocks * per_block;
  const int passes = (int)((n + Q - 1) / Q);
  kern<<<blocks, kQ8pThreads, 0, stream>>>(p, passes);
  MKE_CHECK_LAUNCH("rel_fused_q8p_kernel");
  return 0;
}

// returns 1 when this launch is not covered (caller falls back to rel_fused_q8_kernel)
int launch_rel_q8p(const RelStepParams& p, int cfg, cudaStream_t stream) {
  if (p.sampled || p.neg_out != nullptr || p.trace != nullptr || p.dbg != 0) return 1;
  if (p.K > MKE_MAX_NEG || p.neg_ent == nullptr || p.neg_side == nullptr) return 1;
  const int R = 3 + p.K;
  const bool deep = 2 * 6 <= R, shallow = 2 * 4 <= R;
  if (!shallow) return 1;
  const bool d4 = (cfg & 2) != 0 || !deep;
  const bool roomy = (cfg & 1) != 0;  // experiment: 4 blocks per SM at most, no register cap to speak of
  const bool sh = p.sharded != 0;
  switch (p.stride) {
#define MKE_Q8P_CASE(STRIDE, FPL, MINB)                                                                         \
  case STRIDE:                                                                                                  \
    if (roomy && !sh) return launch_q8p<FPL, 6, 4, false>(p, stream);                                           \
    if (d4) return sh ? launch_q8p<FPL, 4, MINB, true>(p, stream) : launch_q8p<FPL, 4, MINB, false>(p, stream); \
    return sh ? launch_q8p<FPL, 6, MINB, true>(p, stream) : launch_q8p<FPL, 6, MINB, false>(p, stream);
    MKE_Q8P_CASE(32, 4, 6)
    MKE_Q8P_CASE(64, 8, 6)
    MKE_Q8P_CASE(80, 10, 6)
    MKE_Q8P_CASE(104, 13, 5)
    MKE_Q8P_CASE(128, 16, 5)
#undef MKE_Q8P_CASE
    default: return 1;
  }
}

}  // namespace mke
