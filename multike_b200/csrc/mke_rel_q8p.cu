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
#include "mke_q8.cuh"

namespace mke {

constexpr int kIdStride = 36;  // h r t + MKE_MAX_NEG ids + side word; 2*36 = 8 (mod 32): quarters on distinct banks
constexpr int kQ8pThreads = 96;
constexpr int kQ8pWarps = kQ8pThreads / 32;

// register cap for MINB resident blocks: each of the 4 SM sub-partitions owns 16 384 registers and
// gets ceil(3 MINB / 4) of the warps
constexpr int q8p_max_regs(int minb) {
  const int per_smsp = (kQ8pWarps * minb + 3) / 4;
  const int r = ((16384 / (per_smsp * 32)) / 8) * 8;
  return r > 255 ? 248 : r;
}

template <int FPL, int D, int MINB, bool HOLD>
__global__ void __launch_bounds__(kQ8pThreads) __maxnreg__(q8p_max_regs(MINB)) rel_fused_q8p_kernel(const RelStepParams p, const int passes) {
  constexpr int WARPS = kQ8pWarps;
  constexpr int stride = FPL * 8;
  using Ring = Stage<FPL, D>;
  __shared__ __align__(128) unsigned char s_ring[WARPS][Ring::kBytes];
  __shared__ int32_t s_ids[WARPS][kQPerWarp][2][kIdStride];
  __shared__ float s_loss[WARPS];
  const int lane = threadIdx.x & 31;
  const int sub = lane & 7;
  const int q = lane >> 3;
  const int wib = threadIdx.x >> 5;
  Ring stg;
  stg.base = (uint32_t)__cvta_generic_to_shared(&s_ring[wib][0]);
  stg.lane = lane;
  RowScatter<FPL, false> out;
  out.buf = 0;
  out.qmask = 0xffu << (lane & 24);
  out.sub = sub;
  volatile int32_t* const ids0 = s_ids[wib][q][0];
  const uint32_t ids_base = (uint32_t)__cvta_generic_to_shared(&s_ids[wib][q][0][0]);
  const int total = p.len1 + p.len2;
  const int K = p.K;
  const int R = 3 + K;  // rows per positive in the stream
  const int Q = gridDim.x * WARPS * kQPerWarp;
  const int g = (blockIdx.x * WARPS + wib) * kQPerWarp + q;
  float* const rel_grad = rel_grad_replica(p);
  float loss_local = 0.f;

  // id list of positive i -> buffer `buf`, asynchronously (the copies join the next commit group)
  auto ids_issue = [&](int buf, int i) {
    const uint32_t dst = ids_base + (uint32_t)buf * (kIdStride * 4);
    if (i < total) {
      const int32_t* prow = i < p.len1 ? p.pos1 + 3 * (size_t)i : p.pos2 + 3 * (size_t)(i - p.len1);
      const int32_t* nrow = p.neg_ent + (size_t)i * K;
      for (int c = sub; c <= R; c += 8) {
        const void* src = c < 3 ? (const void*)(prow + c)
                                : (c < R ? (const void*)(nrow + (c - 3)) : (const void*)(p.neg_side + i));
        cp_async4(dst + 4u * c, src);
      }
    } else {  // idle quarter of the last pass: row 0 of each table, nothing is written back
      for (int c = sub; c <= R; c += 8) asm volatile("st.shared.u32 [%0], %1;" ::"r"(dst + 4u * c), "r"(0) : "memory");
    }
  };

  // producer side of the row stream
  int iss_n = 0, iss_c = 0, iss_slot = 0, cons_slot = 0;
  auto issue_next = [&]() {
    if (iss_n < passes) {
      if (iss_c == 0) __syncwarp();  // the id list of positive iss_n was landed by other lanes' copies
      const int32_t id = ids0[(iss_n & 1) * kIdStride + iss_c];
      const float* row = (iss_c == 1) ? p.rel_var + (size_t)id * stride : ent_var_row(p, id, stride);
      stg.issue(iss_slot, row, sub);
    }
    cp_async_commit();
    iss_slot = (iss_slot + 1 == D) ? 0 : iss_slot + 1;
    if (++iss_c == R) {
      iss_c = 0;
      ++iss_n;
    }
  };

  ids_issue(0, g);
  cp_async_commit();
  cp_async_wait<0>();
  __syncwarp();
#pragma unroll 1
  for (int k = 0; k < D; ++k) issue_next();

#pragma unroll 1
  for (int n = 0; n < passes; ++n) {
    const int i = g + n * Q;
    const bool active = i < total;
    volatile int32_t* const ids = ids0 + (n & 1) * kIdStride;
    const int32_t h = ids[0], r = ids[1], t = ids[2];
    const uint32_t side = (uint32_t)ids[R];
    if (n + 1 < passes) ids_issue((n + 1) & 1, i + Q);
    // ---- positive term ---------------------------------------------------------------------
    const bool side0 = (side & 1u) != 0u;  // side of negative 0: true = head replaced
    const float sgn = side0 ? 1.f : -1.f;
    float base[FPL], acc[FPL];
    float bb = 0.f;  // |base|^2
    {
      float sp = 0.f;
      if constexpr (!HOLD) {
        // rows leave the ring one by one: each slot is refilled as soon as its row is in registers
      float xh[FPL], xr[FPL], xt[FPL];
      cp_async_wait<D - 1>();
      stg.read(cons_slot, xh);
      cons_slot = (cons_slot + 1 == D) ? 0 : cons_slot + 1;
      float sh = sumsq<FPL>(xh);
      issue_next();
      cp_async_wait<D - 1>();
      stg.read(cons_slot, xr);
      cons_slot = (cons_slot + 1 == D) ? 0 : cons_slot + 1;
      float sr = sumsq<FPL>(xr);
      issue_next();
      cp_async_wait<D - 1>();
      stg.read(cons_slot, xt);
      cons_slot = (cons_slot + 1 == D) ? 0 : cons_slot + 1;
      float st = sumsq<FPL>(xt);
      issue_next();
      qsum3(sh, sr, st);
      const float ih = p.ent_norm ? rsqrtf(fmaxf(sh, kNormEps)) : 1.f;
      const float ir = p.rel_norm ? rsqrtf(fmaxf(sr, kNormEps)) : 1.f;
      const float it = p.ent_norm ? rsqrtf(fmaxf(st, kNormEps)) : 1.f;
#pragma unroll
      for (int k = 0; k < FPL; ++k) {
        const float hh = xh[k] * ih, tt = xt[k] * it;
        const float pd = fmaf(xr[k], ir, hh) - tt;  // pos_distance (losses.py:5)
        sp = fmaf(pd, pd, sp);
        acc[k] = pd;
        // head side: nd = e^ + (r^ - t^) = e^ + (pd - h^);  tail side: nd = (h^ + r^) - e^ = (pd + t^) - e^
        base[k] = side0 ? (pd - hh) : (pd + tt);
        bb = fmaf(base[k], base[k], bb);
      }
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) {
        sp += __shfl_xor_sync(kFull, sp, o);
        bb += __shfl_xor_sync(kFull, bb, o);
      }
      } else {
      // the three rows stay in their ring slots and are read twice (norms first, then the
      // distance piece by piece): only base/acc and one piece of each row are ever live
      constexpr int NV4 = FPL / 4, REM = FPL % 4;
      cp_async_wait<D - 3>();
      const int s0 = cons_slot;
      const int s1 = (s0 + 1 == D) ? 0 : s0 + 1;
      const int s2 = (s1 + 1 == D) ? 0 : s1 + 1;
      cons_slot = (s2 + 1 == D) ? 0 : s2 + 1;
      float sh, sr, st;
      {
        float x[FPL];
        stg.read(s0, x);
        sh = sumsq<FPL>(x);
        stg.read(s1, x);
        sr = sumsq<FPL>(x);
        stg.read(s2, x);
        st = sumsq<FPL>(x);
      }
      qsum3(sh, sr, st);
      const float ih = p.ent_norm ? rsqrtf(fmaxf(sh, kNormEps)) : 1.f;
      const float ir = p.rel_norm ? rsqrtf(fmaxf(sr, kNormEps)) : 1.f;
      const float it = p.ent_norm ? rsqrtf(fmaxf(st, kNormEps)) : 1.f;
      auto piece = [&](int k, float xh, float xr, float xt) {
        const float hh = xh * ih, tt = xt * it;
        const float pd = fmaf(xr, ir, hh) - tt;  // pos_distance (losses.py:5)
        sp = fmaf(pd, pd, sp);
        acc[k] = pd;
        // head side: nd = e^ + (r^ - t^) = e^ + (pd - h^);  tail side: nd = (h^ + r^) - e^ = (pd + t^) - e^
        base[k] = side0 ? (pd - hh) : (pd + tt);
        bb = fmaf(base[k], base[k], bb);
      };
#pragma unroll
      for (int c = 0; c < NV4; ++c) {
        float vh[4], vr[4], vt[4];
        stg.read4(s0, c, vh);
        stg.read4(s1, c, vr);
        stg.read4(s2, c, vt);
#pragma unroll
        for (int k = 0; k < 4; ++k) piece(4 * c + k, vh[k], vr[k], vt[k]);
      }
      if constexpr (REM > 0) {
        float vh[REM], vr[REM], vt[REM];
        stg.read_tail(s0, vh);
        stg.read_tail(s1, vr);
        stg.read_tail(s2, vt);
#pragma unroll
        for (int k = 0; k < REM; ++k) piece(4 * NV4 + k, vh[k], vr[k], vt[k]);
      }
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) {
        sp += __shfl_xor_sync(kFull, sp, o);
        bb += __shfl_xor_sync(kFull, bb, o);
      }
      // the three slots are free: the stream moves on by three rows
      issue_next();
      issue_next();
      issue_next();
      }
      float lpos, sg;
      softplus_sigmoid(sp, lpos, sg);  // log(1 + exp(-pos_score)), pos_score = -sp (losses.py:7,9)
      const float wgt = (p.w != nullptr && active ? __ldg(p.w + i) : 1.f) * p.pos_scale;
      if (active) loss_local += wgt * lpos;
      const float cp = 2.f * sg * wgt;
#pragma unroll
      for (int k = 0; k < FPL; ++k) acc[k] *= cp;  // d loss / d pd; the K-loop adds the negatives
      // the endpoint that no same-side negative shares gets its positive-term gradient now
      if (active) out.add(ent_grad_row(p, side0 ? h : t, stride), acc, sgn);
    }
    // ---- negatives ---------------------------------------------------------------------------
#pragma unroll 1
    for (int j = 0; j < K; ++j) {
      float x[FPL];
      cp_async_wait<D - 1>();
      stg.read(cons_slot, x);
      cons_slot = (cons_slot + 1 == D) ? 0 : cons_slot + 1;
      const int32_t e = ids[3 + j];
      // |nd|^2 = |base + s ie e|^2 = |base|^2 + 2 s ie (base.e) + ie^2 (e.e)
      float ee = 0.f, be = 0.f;
#pragma unroll
      for (int k = 0; k < FPL; ++k) {
        ee = fmaf(x[k], x[k], ee);
        be = fmaf(x[k], base[k], be);
      }
      issue_next();  // the slot is free: the sums above consumed x
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) {
        ee += __shfl_xor_sync(kFull, ee, o);
        be += __shfl_xor_sync(kFull, be, o);
      }
      const float ie = p.ent_norm ? rsqrtf(fmaxf(ee, kNormEps)) : 1.f;
      const float sie = sgn * ie;
      const float sn = fmaf(ie * ie, ee, fmaf(2.f * sie, be, bb));  // -neg_score (losses.py:8)
      float lneg, sg;
      softplus_sigmoid(-sn, lneg, sg);  // log(1 + exp(neg_score)), neg_score = -sn
      const bool odd = (((side >> j) & 1u) != 0u) != side0;
      const bool on = active && !odd;
      const float cn = on ? -2.f * sg : 0.f;
      if (on) loss_local += lneg;
#pragma unroll
      for (int k = 0; k < FPL; ++k) {
        x[k] = fmaf(x[k], sie, base[k]);  // neg_distance (losses.py:6)
        acc[k] = fmaf(cn, x[k], acc[k]);
      }
      if (on) out.add(ent_grad_row(p, e, stride), x, cn * sgn);
    }
    // ---- r gets every same-side term, the shared endpoint likewise ---------------------------
    if (active) {
      out.add(rel_grad + (size_t)r * stride, acc, 1.f);
      out.add(ent_grad_row(p, side0 ? t : h, stride), acc, -sgn);
      for (int c = sub; c < K; c += 8) ent_mark(p, ids[3 + c]);
      if (sub == 0) {
        ent_mark(p, h);
        ent_mark(p, t);
        mark_touched(p.rel_touched, r);
      }
    }
    // ---- negatives on the other side than negative 0 (rare), once base/acc are dead ----------
    const bool mixed = active && side != 0u && side != low_ones(K);
    if (__any_sync(kFull, mixed)) {
      for (int j = 1; j < K; ++j) {
        const bool odd = active && ((((side >> j) & 1u) != 0u) != side0);
        if (__any_sync(kFull, odd))
          loss_local += odd_negative<FPL>(
              ent_var_row(p, h, stride), p.rel_var + (size_t)r * stride, ent_var_row(p, t, stride),
              ent_var_row(p, ids[3 + j], stride), ent_grad_row(p, h, stride), rel_grad + (size_t)r * stride,
              ent_grad_row(p, t, stride), ent_grad_row(p, ids[3 + j], stride), p.ent_norm, p.rel_norm, !side0, odd,
              sub);
      }
    }
    __syncwarp();  // this positive's id buffer is rewritten two positives from now
  }
  cp_async_wait<0>();
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
template <int FPL, int D, int MINB, bool HOLD>
static int launch_q8p(const RelStepParams& p, cudaStream_t stream) {
  auto kern = rel_fused_q8p_kernel<FPL, D, MINB, HOLD>;
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
  const bool hold = (cfg & 1) != 0;  // positive rows held in the ring and read twice (fewer registers)
  const bool d4 = (cfg & 2) != 0 || !deep;
  switch (p.stride) {
#define MKE_Q8P_CASE(STRIDE, FPL, MINB)                                          \
  case STRIDE:                                                                   \
    if (d4) return hold ? launch_q8p<FPL, 4, MINB, true>(p, stream) : launch_q8p<FPL, 4, MINB, false>(p, stream); \
    return hold ? launch_q8p<FPL, 6, MINB, true>(p, stream) : launch_q8p<FPL, 6, MINB, false>(p, stream);
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
