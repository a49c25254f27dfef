// Row-stream phase-1 body shared by rel_fused_q8p_kernel (one launch per step, mke_rel_q8p.cu) and the
// persistent step kernel (mke_rel_persist.cu).  The schedule is described in mke_rel_q8p.cu.
#pragma once
#include "mke_q8.cuh"

namespace mke {

constexpr int kIdStride = 37;  // h r t + MKE_MAX_NEG ids + side word + ownership word; quarters 74 = 10 (mod 32) banks apart
// MUFU forms without the denormal / range guards of rsqrtf, __expf and __logf: every argument here is
// >= 1e-12 (clamped sums of squares) or in [-9, 9] (squared distances of unit rows), so the .ftz
// forms are exact replacements and save ~12 instructions per row.
__device__ __forceinline__ float mufu_rsqrt(float x) {
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float mufu_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float mufu_lg2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float mufu_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// log(1 + exp(x)) and sigmoid(x) as the reference writes them (losses.py:9-10), cf. softplus_sigmoid
__device__ __forceinline__ void softplus_sigmoid_mufu(float x, float& sp, float& sg) {
  const float ex = mufu_ex2(x * 1.4426950408889634f);
  const float one_p = 1.0f + ex;
  sp = mufu_lg2(one_p) * 0.6931471805599453f;
  sg = ex * mufu_rcp(one_p);
}

// the positives and pre-drawn negatives of one step (what varies from step to step)
struct StepBatch {
  const int32_t* pos1;
  int len1;
  const int32_t* pos2;
  int len2;
  const int32_t* neg_ent;
  const uint32_t* neg_side;
};

// One quarter's share of a step: positives g, g + Q, ..., `passes` of them (indices >= len1 + len2 are
// idle passes that read row 0 and write nothing).  ring_w: this warp's Stage<FPL, D>::kBytes of shared
// memory; ids_q: this quarter's [2][kIdStride] id lists; rel_grad: the relation-gradient replica this
// warp reduces into.  Returns the lane's loss share (meaningful in lane sub == 0 of each quarter).
template <int FPL, int D, bool SHARDED>
__device__ __forceinline__ float q8p_stream(const RelStepParams& p, const StepBatch& b, const int passes, const int Q,
                                            const int g, unsigned char* ring_w, int32_t* ids_q, float* const rel_grad,
                                            const int lane) {
  constexpr int stride = FPL * 8;
  constexpr int H = FPL / 2;           // packed fp32 pairs per lane (fma.rn.f32x2: one issue slot, two FMAs)
  constexpr bool ODD = (FPL & 1) != 0;  // + one scalar tail element
  using Ring = Stage<FPL, D>;
  const int sub = lane & 7;
  Ring stg;
  stg.base = (uint32_t)__cvta_generic_to_shared(ring_w);
  stg.lane = lane;
  volatile int32_t* const ids0 = ids_q;
  const uint32_t ids_base = (uint32_t)__cvta_generic_to_shared(ids_q);
  const int total = b.len1 + b.len2;
  const int K = p.K;
  const int R = 3 + K;  // rows per positive in the stream
  float loss_local = 0.f;

  // entity rows: one local table, or the owner's shard through its peer mapping (compile-time choice)
  auto evar = [&](int32_t id) -> const float* {
    if constexpr (!SHARDED) return p.ent_var + (size_t)id * stride;
    int s;
    int32_t l;
    p.smap.locate(id, s, l);
    return p.sh.var[s] + (size_t)l * stride;
  };
  auto egrad = [&](int32_t id) -> float* {
    if constexpr (!SHARDED) return p.ent_grad + (size_t)id * stride;
    int s;
    int32_t l;
    p.smap.locate(id, s, l);
    return p.sh.grad[s] + (size_t)l * stride;
  };
  auto emark = [&](int32_t id) {
    if constexpr (!SHARDED) {
      mark_touched(p.ent_touched, id);
    } else {
      int s;
      int32_t l;
      p.smap.locate(id, s, l);
      mark_touched(p.sh.touched[s], l);
    }
  };

  // id list of positive i -> buffer `buf`, asynchronously (the copies join the next commit group)
  auto ids_issue = [&](int buf, int i) {
    const uint32_t dst = ids_base + (uint32_t)buf * (kIdStride * 4);
    if (i < total) {
      const int32_t* prow = i < b.len1 ? b.pos1 + 3 * (size_t)i : b.pos2 + 3 * (size_t)(i - b.len1);
      const int32_t* nrow = b.neg_ent + (size_t)i * K;
      for (int c = sub; c <= R; c += 8) {
        const void* src = c < 3 ? (const void*)(prow + c)
                                : (c < R ? (const void*)(nrow + (c - 3)) : (const void*)(b.neg_side + i));
        cp_async4(dst + 4u * c, src);
      }
      if (p.neg_valid != nullptr && sub == 7) cp_async4(dst + 4u * (R + 1), p.neg_valid + i);
    } else {  // idle quarter of the last pass: row 0 of each table, nothing is written back
      for (int c = sub; c <= R + 1; c += 8) asm volatile("st.shared.u32 [%0], %1;" ::"r"(dst + 4u * c), "r"(0) : "memory");
    }
  };

  // producer side of the row stream
  // One ring position serves both sides: take() reads the oldest slot, issue_next() refills that
  // very slot with the row D positions further down the stream and moves on.
  int iss_n = 0, iss_c = 0, slot = 0;
  auto issue_next = [&]() {
    if (iss_n < passes) {
      if (iss_c == 0) __syncwarp();  // the id list of positive iss_n was landed by other lanes' copies
      const int32_t id = ids0[(iss_n & 1) * kIdStride + iss_c];
      const float* row = (iss_c == 1) ? p.rel_var + (size_t)id * stride : evar(id);
      stg.issue(slot, row, sub);
    }
    cp_async_commit();
    slot = (slot + 1 == D) ? 0 : slot + 1;
    if (++iss_c == R) {
      iss_c = 0;
      ++iss_n;
    }
  };
  auto take = [&](float (&x)[FPL]) {  // oldest row of the ring -> registers
    cp_async_wait<D - 1>();
    stg.read(slot, x);
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
    // negatives of this positive that are this launch's, and whether its positive term is
    const uint32_t mine = p.neg_valid != nullptr ? (uint32_t)ids[R + 1] : 0xffffffffu;
    const bool pos_on = active && i >= p.pos_own_lo && i < p.pos_own_hi;
    if (n + 1 < passes) ids_issue((n + 1) & 1, i + Q);
    // ---- positive term ---------------------------------------------------------------------
    const bool side0 = (side & 1u) != 0u;  // side of negative 0: true = head replaced
    const float sgn = side0 ? 1.f : -1.f;
    // base = the part of a same-side negative's distance that does not depend on the negative;
    // accs = sgn * (d loss / d pos_distance + sum over negatives of d loss / d neg_distance): kept
    // pre-multiplied by sgn so that a negative costs one packed multiply and one packed add
    float2 base2[H > 0 ? H : 1], accs2[H > 0 ? H : 1];
    float base_t = 0.f, accs_t = 0.f;
    float bb = 0.f;  // |base|^2
    {
      float xh[FPL], xr[FPL], xt[FPL];
      take(xh);
      float sh = sumsq<FPL>(xh);
      issue_next();
      take(xr);
      float sr = sumsq<FPL>(xr);
      issue_next();
      take(xt);
      float st = sumsq<FPL>(xt);
      issue_next();
      qsum3(sh, sr, st);
      const float ih = p.ent_norm ? mufu_rsqrt(fmaxf(sh, kNormEps)) : 1.f;
      const float ir = p.rel_norm ? mufu_rsqrt(fmaxf(sr, kNormEps)) : 1.f;
      const float it = p.ent_norm ? mufu_rsqrt(fmaxf(st, kNormEps)) : 1.f;
      float pd[FPL], bs[FPL];
      float sp = 0.f;
#pragma unroll
      for (int k = 0; k < FPL; ++k) {
        const float hh = xh[k] * ih, tt = xt[k] * it;
        pd[k] = fmaf(xr[k], ir, hh) - tt;  // pos_distance (losses.py:5)
        sp = fmaf(pd[k], pd[k], sp);
        // head side: nd = e^ + (r^ - t^) = e^ + (pd - h^);  tail side: nd = (h^ + r^) - e^ = (pd + t^) - e^
        bs[k] = side0 ? (pd[k] - hh) : (pd[k] + tt);
        bb = fmaf(bs[k], bs[k], bb);
      }
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) {
        sp += __shfl_xor_sync(kFull, sp, o);
        bb += __shfl_xor_sync(kFull, bb, o);
      }
      float lpos, sg;
      softplus_sigmoid_mufu(sp, lpos, sg);  // log(1 + exp(-pos_score)), pos_score = -sp (losses.py:7,9)
      const float wgt = pos_on ? (p.w != nullptr ? __ldg(p.w + i) : 1.f) * p.pos_scale : 0.f;
      if (pos_on) loss_local += wgt * lpos;
      const float cps = 2.f * sg * wgt * sgn;
      float first[FPL];  // sgn * d loss / d pos_distance
#pragma unroll
      for (int k = 0; k < FPL; ++k) first[k] = pd[k] * cps;
#pragma unroll
      for (int k = 0; k < H; ++k) {
        base2[k] = make_float2(bs[2 * k], bs[2 * k + 1]);
        accs2[k] = make_float2(first[2 * k], first[2 * k + 1]);
      }
      if constexpr (ODD) {
        base_t = bs[FPL - 1];
        accs_t = first[FPL - 1];
      }
      // the endpoint that no same-side negative shares gets its positive-term gradient now
      if (pos_on) red_row<FPL>(egrad(side0 ? h : t), sub, first, 1.f);
    }
    // ---- negatives ---------------------------------------------------------------------------
#pragma unroll 1
    for (int j = 0; j < K; ++j) {
      float x[FPL];
      take(x);
      const int32_t e = ids[3 + j];
      // |nd|^2 = |base + s ie e|^2 = |base|^2 + 2 s ie (base.e) + ie^2 (e.e)
      float2 ee2 = make_float2(0.f, 0.f), be2 = make_float2(0.f, 0.f);
#pragma unroll
      for (int k = 0; k < H; ++k) {
        const float2 xv = make_float2(x[2 * k], x[2 * k + 1]);
        ee2 = __ffma2_rn(xv, xv, ee2);
        be2 = __ffma2_rn(xv, base2[k], be2);
      }
      float ee = ee2.x + ee2.y, be = be2.x + be2.y;
      if constexpr (ODD) {
        ee = fmaf(x[FPL - 1], x[FPL - 1], ee);
        be = fmaf(x[FPL - 1], base_t, be);
      }
      issue_next();  // the slot is free: the sums above consumed x
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) {
        ee += __shfl_xor_sync(kFull, ee, o);
        be += __shfl_xor_sync(kFull, be, o);
      }
      const float ie = p.ent_norm ? mufu_rsqrt(fmaxf(ee, kNormEps)) : 1.f;
      const float sie = sgn * ie;
      const float sn = fmaf(ie * ie, ee, fmaf(2.f * sie, be, bb));  // -neg_score (losses.py:8)
      float lneg, sg;
      softplus_sigmoid_mufu(-sn, lneg, sg);  // log(1 + exp(neg_score)), neg_score = -sn
      const bool odd = (((side >> j) & 1u) != 0u) != side0;
      const bool on = active && !odd && ((mine >> j) & 1u) != 0u;
      const float cs = on ? -2.f * sg * sgn : 0.f;  // sgn * d loss / d |nd|^2 * 2
      if (on) loss_local += lneg;
      const float2 sie2 = make_float2(sie, sie), cs2 = make_float2(cs, cs);
      float y[FPL];  // sgn * d loss / d neg_distance = the gradient row of the corrupted entity
#pragma unroll
      for (int k = 0; k < H; ++k) {
        const float2 nd = __ffma2_rn(make_float2(x[2 * k], x[2 * k + 1]), sie2, base2[k]);  // neg_distance (losses.py:6)
        const float2 yy = __fmul2_rn(nd, cs2);
        accs2[k] = __fadd2_rn(accs2[k], yy);
        y[2 * k] = yy.x;
        y[2 * k + 1] = yy.y;
      }
      if constexpr (ODD) {
        y[FPL - 1] = fmaf(x[FPL - 1], sie, base_t) * cs;
        accs_t += y[FPL - 1];
      }
      if (on) red_row<FPL>(egrad(e), sub, y, 1.f);
    }
    // ---- r gets every same-side term, the shared endpoint likewise ---------------------------
    if (active && (pos_on || (mine & low_ones(K)) != 0u)) {
      float a[FPL];
#pragma unroll
      for (int k = 0; k < H; ++k) {
        a[2 * k] = accs2[k].x;
        a[2 * k + 1] = accs2[k].y;
      }
      if constexpr (ODD) a[FPL - 1] = accs_t;
      red_row<FPL>(rel_grad + (size_t)r * stride, sub, a, sgn);
      red_row<FPL>(egrad(side0 ? t : h), sub, a, -1.f);
      for (int c = sub; c < K; c += 8)
        if ((mine >> c) & 1u) emark(ids[3 + c]);
      if (sub == 0) {
        emark(h);
        emark(t);
        mark_touched(p.rel_touched, r);
      }
    }
    // ---- negatives on the other side than negative 0 (rare), once base/accs are dead ---------
    const bool mixed = active && side != 0u && side != low_ones(K);
    if (__any_sync(kFull, mixed)) {
      for (int j = 1; j < K; ++j) {
        const bool odd = active && ((mine >> j) & 1u) != 0u && ((((side >> j) & 1u) != 0u) != side0);
        if (__any_sync(kFull, odd))
          loss_local += odd_negative<FPL>(evar(h), p.rel_var + (size_t)r * stride, evar(t), evar(ids[3 + j]), egrad(h),
                                          rel_grad + (size_t)r * stride, egrad(t), egrad(ids[3 + j]), p.ent_norm,
                                          p.rel_norm, !side0, odd, sub);
      }
    }
    __syncwarp();  // this positive's id buffer is rewritten two positives from now
  }
  cp_async_wait<0>();
  return loss_local;
}

}  // namespace mke
