// Relation view, phase 1, quarter-warp layout (the default data path).
//
// One positive triple with its K single-side-corrupted negatives is owned by a QUARTER warp
// (8 lanes); a row of `stride` floats is spread over the 8 lanes as FPL = stride/8 floats per
// lane (stride 80: two float4 + one float2 per lane, i.e. one 128-byte segment per 16-byte
// load instruction and quarter).  Against the warp-per-positive kernel (mke_rel.cu) this
//   * keeps all 32 lanes busy (a 75-float row only fills 19 lanes of float4),
//   * amortises the per-row overhead (two reductions, exp/log, addressing) over 4 rows,
//   * shortens reductions to 3 shuffle steps,
//   * holds only two live row vectors per positive (base, acc) so that >= 128 positives are in
//     flight per SM -- the kernel is a single wave at batch 20 000 and is bounded by L2/HBM
//     row traffic, not by instruction issue (profiles/ has the ncu evidence).
// Arithmetic follows losses.py:4-12 on l2-normalised rows (base/initializers.py:26):
//   pd = h^ + r^ - t^                      loss += w * log(1 + exp(|pd|^2))
//   nd = e^ + r^ - t^ (head side) or h^ + r^ - e^ (tail side)   loss += log(1 + exp(-|nd|^2))
// and accumulates d loss / d row into the gradient tables with red.global.add (vectorised).
#include <cstdlib>
#include "mke_q8.cuh"

namespace mke {


__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)::"memory");
  return t;
}
#define MKE_TRACE(slot)                                                                        \
  do {                                                                                         \
    if (p.trace != nullptr && lane == 0)                                                       \
      p.trace[(size_t)(blockIdx.x * WARPS + wib) * 32 + (slot)] = gtimer();                   \
  } while (0)

template <int FPL, int THREADS, int MINB, bool BULK>
__global__ void __launch_bounds__(THREADS, MINB) rel_fused_q8_kernel(const RelStepParams p) {
  constexpr int WARPS = THREADS / 32;
  constexpr int stride = FPL * 8;
  __shared__ __align__(128) unsigned char s_stage[WARPS][Stage<FPL>::kBytes];
  __shared__ __align__(128) float s_out[BULK ? WARPS : 1][kQPerWarp][BULK ? stride : 1];
  __shared__ int32_t s_pick_all[WARPS][kQPerWarp][kPickStride];
  __shared__ float s_loss[WARPS];
  const int lane = threadIdx.x & 31;
  const int sub = lane & 7;
  const int q = lane >> 3;
  const int wib = threadIdx.x >> 5;
  volatile int32_t* pick = s_pick_all[wib][q];
  Stage<FPL> stg;
  stg.base = (uint32_t)__cvta_generic_to_shared(&s_stage[wib][0]);
  stg.lane = lane;
  RowScatter<FPL, BULK> out;
  out.buf = (uint32_t)__cvta_generic_to_shared(&s_out[BULK ? wib : 0][q][0]);
  out.qmask = 0xffu << (lane & 24);
  out.sub = sub;
  const int n = p.len1 + p.len2;
  const int K = p.K;
  const int per_pass = gridDim.x * WARPS * kQPerWarp;
  float* const rel_grad = rel_grad_replica(p);
  float loss_local = 0.f;
  MKE_TRACE(0);

  for (int i0 = (blockIdx.x * WARPS + wib) * kQPerWarp; i0 < n; i0 += per_pass) {
    const int i = i0 + q;
    const bool valid = i < n;
    int32_t h = 0, r = 0, t = 0;
    const bool first = i < p.len1;
    if (valid) {
      const int32_t* row = first ? p.pos1 + 3 * (size_t)i : p.pos2 + 3 * (size_t)(i - p.len1);
      h = __ldg(row);
      r = __ldg(row + 1);
      t = __ldg(row + 2);
    }
    {
      uint32_t side = 0u;
      uint32_t mine = 0xffffffffu;  // negatives of this positive that are this launch's (RelStepParams.neg_valid)
      int Kw = K;                   // rows the K-loop of this warp walks
      const bool active = valid;
      const bool pos_on = valid && i >= p.pos_own_lo && i < p.pos_own_hi;
      if (h + r + t == -3) MKE_TRACE(15);  // (forces the id loads to have landed)
      MKE_TRACE(1);
      // the three rows of the positive travel to shared memory while the sampler probes
      stg.issue(0, ent_var_row(p, h, stride), sub);
      stg.issue(1, p.rel_var + (size_t)r * stride, sub);
      stg.issue(2, ent_var_row(p, t, stride), sub);
      cp_async_commit();
      if (K > 0) {
        if (p.sampled) {
          // idle quarters of the last warp sample from a harmless pool; nothing they draw is used
          KgView kg = kg_view(p, valid ? first : (p.len1 > 0));
          if (!valid || (p.dbg & 16)) {
            kg.neighbours = nullptr;
            kg.set.slots = nullptr;
          }
          side = sample_negs_quarter(kg, h, r, t, K, p.skey, (uint32_t)(p.index_base + i), lane, pick,
                                     p.trace ? p.trace + (size_t)(blockIdx.x * WARPS + wib) * 32 : nullptr);
          if (!valid) {
            side = 0u;
            for (int c = sub; c < K; c += 8) pick[c] = 0;
          }
        } else {
          for (int c = sub; c < K; c += 8) pick[c] = valid ? __ldg(p.neg_ent + (size_t)i * K + c) : 0;
          side = valid ? __ldg(p.neg_side + i) : 0u;
          if (p.neg_valid != nullptr && valid) mine = __ldg(p.neg_valid + i);
        }
        __syncwarp();
        // "negatives where they live", compacted (mke_neg_keep_owned2: this launch's negatives first): the
        // K-loop only runs as far as the longest list of the warp's four positives
        if (p.neg_compact) {
          int cnt = valid ? __popc(mine & low_ones(K)) : 0;
          cnt = max(cnt, __shfl_xor_sync(kFull, cnt, 8));
          cnt = max(cnt, __shfl_xor_sync(kFull, cnt, 16));
          Kw = cnt;
        }
        if (p.neg_out != nullptr && active) {
          for (int c = sub; c < K; c += 8) {
            const bool hs = (side >> c) & 1u;
            const int32_t e = pick[c];
            int32_t* o = p.neg_out + ((size_t)i * K + c) * 3;
            o[0] = hs ? e : h;
            o[1] = r;
            o[2] = hs ? t : e;
          }
        }
      }
      MKE_TRACE(2);
      // ---- positive term -----------------------------------------------------------------
      const bool side0 = (side & 1u) != 0u;  // side of negative 0: true = head replaced
      const float sgn = side0 ? 1.f : -1.f;
      float base[FPL], acc[FPL];
      float bb = 0.f;  // |base|^2
      {
        float xh[FPL], xr[FPL], xt[FPL];
        cp_async_wait<0>();
        stg.read(0, xh);
        stg.read(1, xr);
        stg.read(2, xt);
        float sh = sumsq<FPL>(xh), sr = sumsq<FPL>(xr), st = sumsq<FPL>(xt);
        qsum3(sh, sr, st);
        if (sh == -1.f) MKE_TRACE(15);
        MKE_TRACE(3);
        // the slots are free again (their contents fed the sums above): first negatives go out
#pragma unroll
        for (int j = 0; j < kSlots; ++j) {
          if (j < Kw) stg.issue(j, ent_var_row(p, pick[j], stride), sub);
          cp_async_commit();
        }
        const float ih = p.ent_norm ? rsqrtf(fmaxf(sh, kNormEps)) : 1.f;
        const float ir = p.rel_norm ? rsqrtf(fmaxf(sr, kNormEps)) : 1.f;
        const float it = p.ent_norm ? rsqrtf(fmaxf(st, kNormEps)) : 1.f;
        float sp = 0.f;
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
        float lpos, sg;
        softplus_sigmoid(sp, lpos, sg);  // log(1 + exp(-pos_score)), pos_score = -sp (losses.py:7,9)
        const float wgt = pos_on ? (p.w != nullptr ? __ldg(p.w + i) : 1.f) * p.pos_scale : 0.f;
        if (pos_on) loss_local += wgt * lpos;
        const float cp = 2.f * sg * wgt;
#pragma unroll
        for (int k = 0; k < FPL; ++k) acc[k] *= cp;  // d loss / d pd; the K-loop adds the negatives
        // the endpoint that no same-side negative shares gets its positive-term gradient now
        if (pos_on && !(p.dbg & 8)) out.add(ent_grad_row(p, side0 ? h : t, stride), acc, sgn);
      }
      MKE_TRACE(4);
      // ---- negatives: rows j+1, j+2 are in flight while row j is scored ------------------------
      int slot = 0;
      for (int j = 0; j < Kw; ++j) {
        if (j < 8) MKE_TRACE(5 + j);
        float x[FPL];
        cp_async_wait<kSlots - 1>();
        stg.read(slot, x);
        const int32_t e = pick[j];
        // |nd|^2 = |base + s ie e|^2 = |base|^2 + 2 s ie (base.e) + ie^2 (e.e): both dot products
        // are reduced together, so a negative costs one shuffle chain instead of two
        float ee = 0.f, be = 0.f;
#pragma unroll
        for (int k = 0; k < FPL; ++k) {
          ee = fmaf(x[k], x[k], ee);
          be = fmaf(x[k], base[k], be);
        }
        // slot is free: request row j + kSlots (the sums above consumed x)
        if (j + kSlots < Kw) stg.issue(slot, ent_var_row(p, pick[j + kSlots], stride), sub);
        cp_async_commit();
        slot = (slot + 1 == kSlots) ? 0 : slot + 1;
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
        const bool on = active && !odd && ((mine >> j) & 1u) != 0u;
        const float cn = on ? -2.f * sg : 0.f;
        if (on) loss_local += lneg;
#pragma unroll
        for (int k = 0; k < FPL; ++k) {
          x[k] = fmaf(x[k], sie, base[k]);  // neg_distance (losses.py:6)
          acc[k] = fmaf(cn, x[k], acc[k]);
        }
        if (on && !(p.dbg & 8)) out.add(ent_grad_row(p, e, stride), x, cn * sgn);
      }
      cp_async_wait<0>();
      MKE_TRACE(13);
      // ---- r gets every same-side term, the shared endpoint likewise -----------------------
      if (active && (pos_on || (mine & low_ones(K)) != 0u)) {
        if (!(p.dbg & 1)) out.add(rel_grad + (size_t)r * stride, acc, 1.f);
        if (!(p.dbg & 8)) out.add(ent_grad_row(p, side0 ? t : h, stride), acc, -sgn);
        if (!(p.dbg & 2)) {
        for (int c = sub; c < K; c += 8)
          if ((mine >> c) & 1u) ent_mark(p, pick[c]);
        if (sub == 0) {
          ent_mark(p, h);
          ent_mark(p, t);
          mark_touched(p.rel_touched, r);
        }
        }
      }
      MKE_TRACE(19);
      // ---- negatives on the other side than negative 0 (rare), once base/acc are dead ------
      const bool mixed = active && side != 0u && side != low_ones(K);
      if (__any_sync(kFull, mixed)) {
        for (int j = 1; j < K; ++j) {
          const bool odd = active && ((mine >> j) & 1u) != 0u && ((((side >> j) & 1u) != 0u) != side0);
          if (__any_sync(kFull, odd))
            loss_local += odd_negative<FPL>(
                ent_var_row(p, h, stride), p.rel_var + (size_t)r * stride, ent_var_row(p, t, stride),
                ent_var_row(p, pick[j], stride), ent_grad_row(p, h, stride), rel_grad + (size_t)r * stride,
                ent_grad_row(p, t, stride), ent_grad_row(p, pick[j], stride), p.ent_norm, p.rel_norm, !side0,
                odd, sub);
        }
      }
      __syncwarp();  // pick[] is rewritten by the next positive
      MKE_TRACE(14);
    }
  }
  out.drain();
  // ---- loss: quarter leaders -> warp -> block -> one fp64 atomic ------------------------------
  float v = (sub == 0) ? loss_local : 0.f;
  v = warp_sum(v);
  if (lane == 0) s_loss[wib] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) a += (double)s_loss[w];
    if (a != 0.0 && !(p.dbg & 4)) atomicAdd(p.loss, a);
  }
}

template <int FPL, int THREADS, int MINB, bool BULK>
static int launch_q8(const RelStepParams& p, cudaStream_t stream) {
  auto kern = rel_fused_q8_kernel<FPL, THREADS, MINB, BULK>;
  constexpr int per_block = (THREADS / 32) * kQPerWarp;
  const int n = p.len1 + p.len2;
  static int per_sm_cached = 0;
  if (per_sm_cached == 0) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, THREADS, 0) != cudaSuccess || per_sm < 1)
      per_sm = 1;
    per_sm_cached = per_sm;
  }
  const int full = sm_count() * per_sm_cached;
  const int need = (n + per_block - 1) / per_block;
  kern<<<need < full ? need : full, THREADS, 0, stream>>>(p);
  MKE_CHECK_LAUNCH("rel_fused_q8_kernel");
  return 0;
}

// strides with a quarter-warp instantiation: 32 (dim<=32), 64, 80 (dim 75), 104 (dim 100), 128
int launch_rel_q8(const RelStepParams& p, cudaStream_t stream) {
  static const int cfg = getenv("MKE_Q8_CFG") ? atoi(getenv("MKE_Q8_CFG")) : 0;  // tuning knob
  // <floats per lane, threads per block, min blocks per SM, TMA bulk-reduce scatter>.  The defaults
  // come from the sweep in profiles/r1_phase1_tuning.md: fewer, fatter blocks (more registers, no
  // spills) beat maximum occupancy because the kernel is bound by row traffic, not by issue.
  // 10..13: persistent row-stream schedule (mke_rel_q8p.cu): bit0 = build without a register cap, bit1 = ring of 4
  if (cfg >= 10 && cfg <= 13) {
    const int rc = launch_rel_q8p(p, cfg - 10, stream);
    if (rc != 1) return rc;  // 1 = launch shape not covered there
  }
  switch (p.stride) {
    case 32: return launch_q8<4, 128, 8, false>(p, stream);
    case 64: return launch_q8<8, 128, 6, false>(p, stream);
    case 80:
      if (cfg == 1) return launch_q8<10, 64, 18, false>(p, stream);
      if (cfg == 2) return launch_q8<10, 64, 16, true>(p, stream);
      if (cfg == 3) return launch_q8<10, 64, 9, false>(p, stream);
      if (cfg == 4) return launch_q8<10, 128, 6, true>(p, stream);
      if (cfg == 6) return launch_q8<10, 256, 2, false>(p, stream);
      return launch_q8<10, 128, 6, false>(p, stream);
    case 104: return launch_q8<13, 128, 5, false>(p, stream);
    case 128: return launch_q8<16, 128, 4, false>(p, stream);
    default: return 1;
  }
}

}  // namespace mke
