// Relation view, phase 1: fused gather -> ||h+r-t||^2 score -> logistic loss -> gradient ->
// scatter-add, one warp per positive triple with its K single-side-corrupted negatives.
//
// Replaces the TF graph of MultiKE_model.py:123-131 + losses.py:4-12 and (sampled form) the
// Python sampler base/batch.py:86-116.  Two data paths are kept, selected per call:
//   variant 0  rows gathered with 16-byte LDG straight into registers, gradient rows scattered
//              with red.global.add.v4.f32 (REDG.E.ADD.F32x4);
//   variant 1  rows staged through shared memory by the TMA engine (cp.async.bulk + mbarrier,
//              SASS UBLKCP) and gradient rows returned with cp.reduce.async.bulk .add.f32
//              (SASS UBLKRED) -- one bulk op per 300-byte row instead of 19 lanes of LSU work.
// The arithmetic is identical in both; DESIGN.md has the measurements behind the default.
#include <cstdio>
#include <cstdlib>
#include "mke_rel.cuh"

namespace mke {

constexpr int kRelThreads = 256;
constexpr int kRelWarps = kRelThreads / 32;

__device__ __forceinline__ void block_loss_commit(float loss_local, double* loss) {
  __shared__ float s_loss[kRelWarps];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) s_loss[w] = loss_local;
  __syncthreads();
  if (threadIdx.x == 0) {
    double acc = 0.0;
#pragma unroll
    for (int q = 0; q < kRelWarps; ++q) acc += (double)s_loss[q];
    if (acc != 0.0) atomicAdd(loss, acc);
  }
}

// Fetch positive i, its KG and its negatives (drawn or supplied).
__device__ __forceinline__ void fetch_work(const RelStepParams& p, int i, int lane,
                                           volatile int32_t* s_pick, int32_t& h, int32_t& r,
                                           int32_t& t, int32_t& e, uint32_t& side) {
  const bool first = i < p.len1;
  const int32_t* row = first ? p.pos1 + 3 * (size_t)i : p.pos2 + 3 * (size_t)(i - p.len1);
  h = __ldg(row);
  r = __ldg(row + 1);
  t = __ldg(row + 2);
  e = -1;
  side = 0;
  if (p.K > 0) {
    if (p.sampled) {
      sample_negs_warp(first ? p.kg1 : p.kg2, h, r, t, p.K, p.skey, (uint32_t)i, lane, s_pick, e,
                       side);
    } else {
      if (lane < p.K) e = __ldg(p.neg_ent + (size_t)i * p.K + lane);
      side = __ldg(p.neg_side + i);
    }
    if (p.neg_out != nullptr && lane < p.K) {
      const bool hs = (side >> lane) & 1u;
      int32_t* o = p.neg_out + ((size_t)i * p.K + lane) * 3;
      o[0] = hs ? e : h;
      o[1] = r;
      o[2] = hs ? t : e;
    }
  }
}

// --------------------------------------------------------------------------------------------
// variant 0: register path.  NV = float4 pieces per lane (1: stride <= 128, 2: stride <= 256).
// --------------------------------------------------------------------------------------------
template <int NV, int KC>
__global__ void __launch_bounds__(kRelThreads) rel_fused_ldg_kernel(const RelStepParams p) {
  __shared__ int32_t s_pick_all[kRelWarps][32];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  volatile int32_t* s_pick = s_pick_all[wib];
  const int gwarp = blockIdx.x * kRelWarps + wib;
  const int nwarps = gridDim.x * kRelWarps;
  const int n = p.len1 + p.len2;
  bool act[NV];
  int off[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    act[v] = (lane + 32 * v) < p.nchunk;
    off[v] = (lane + 32 * v) * 4;
  }
  float loss_local = 0.f;

  for (int i = gwarp; i < n; i += nwarps) {
    int32_t h, r, t, e;
    uint32_t side;
    fetch_work(p, i, lane, s_pick, h, r, t, e, side);

    const float* ph = p.ent_var + (size_t)h * p.stride;
    const float* pr = p.rel_var + (size_t)r * p.stride;
    const float* pt = p.ent_var + (size_t)t * p.stride;
    float4 xh[NV], xr[NV], xt[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      xh[v] = act[v] ? ldg_f4(ph + off[v]) : f4_zero();
      xr[v] = act[v] ? ldg_f4(pr + off[v]) : f4_zero();
      xt[v] = act[v] ? ldg_f4(pt + off[v]) : f4_zero();
    }
    float sh = 0.f, sr = 0.f, st = 0.f;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      sh += dot4(xh[v], xh[v]);
      sr += dot4(xr[v], xr[v]);
      st += dot4(xt[v], xt[v]);
    }
    warp_sum3(sh, sr, st);
    const float ih = p.ent_norm ? rsqrtf(fmaxf(sh, kNormEps)) : 1.f;
    const float ir = p.rel_norm ? rsqrtf(fmaxf(sr, kNormEps)) : 1.f;
    const float it = p.ent_norm ? rsqrtf(fmaxf(st, kNormEps)) : 1.f;
    // bt = h^ + r^  (tail-corrupted negatives: nd = bt - e^);  bh = r^ - t^  (head: nd = e^ + bh)
    float4 bt[NV], bh[NV], gp[NV], accA[NV], accB[NV];
    float sp = 0.f;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const float4 rr = f4_scale(xr[v], ir);
      const float4 tt = f4_scale(xt[v], it);
      bt[v] = f4_fma(xh[v], ih, rr);
      bh[v] = f4_sub(rr, tt);
      gp[v] = f4_sub(bt[v], tt);  // pos_distance (losses.py:5)
      sp += dot4(gp[v], gp[v]);
      accA[v] = f4_zero();
      accB[v] = f4_zero();
    }
    sp = warp_sum(sp);
    {
      float lpos, sg;
      softplus_sigmoid(sp, lpos, sg);  // log(1+exp(-pos_score)), pos_score = -sp (losses.py:7,9)
      const float wgt = (p.w ? __ldg(p.w + i) : 1.f) * p.pos_scale;
      loss_local += wgt * lpos;
      const float cp = 2.f * sg * wgt;
#pragma unroll
      for (int v = 0; v < NV; ++v) gp[v] = f4_scale(gp[v], cp);
    }

    for (int j0 = 0; j0 < p.K; j0 += KC) {
      float4 xe[KC][NV];
      int32_t ej[KC];
#pragma unroll
      for (int jj = 0; jj < KC; ++jj) {
        const int j = j0 + jj;
        ej[jj] = __shfl_sync(0xffffffffu, e, j & 31);
        const bool valid = j < p.K;
        const float* pe = p.ent_var + (size_t)(valid ? ej[jj] : h) * p.stride;
#pragma unroll
        for (int v = 0; v < NV; ++v) xe[jj][v] = (valid && act[v]) ? ldg_f4(pe + off[v]) : f4_zero();
      }
#pragma unroll
      for (int jj = 0; jj < KC; ++jj) {
        const int j = j0 + jj;
        if (j < p.K) {  // warp-uniform
          float se = 0.f;
#pragma unroll
          for (int v = 0; v < NV; ++v) se += dot4(xe[jj][v], xe[jj][v]);
          se = warp_sum(se);
          const float ie = p.ent_norm ? rsqrtf(fmaxf(se, kNormEps)) : 1.f;
          const bool hs = (side >> j) & 1u;
          float4 nd[NV];
          float sn = 0.f;
#pragma unroll
          for (int v = 0; v < NV; ++v) {
            // neg_distance (losses.py:6): head side e^ + r^ - t^, tail side h^ + r^ - e^
            nd[v] = hs ? f4_fma(xe[jj][v], ie, bh[v]) : f4_fma(xe[jj][v], -ie, bt[v]);
            sn += dot4(nd[v], nd[v]);
          }
          sn = warp_sum(sn);
          float lneg, sg;
          softplus_sigmoid(-sn, lneg, sg);  // log(1+exp(neg_score)), neg_score = -sn
          loss_local += lneg;
          const float cn = -2.f * sg;
          float* ge = p.ent_grad + (size_t)ej[jj] * p.stride;
#pragma unroll
          for (int v = 0; v < NV; ++v) {
            const float4 gn = f4_scale(nd[v], cn);
            if (hs) {
              accB[v] = f4_add(accB[v], gn);
              if (act[v]) red_add_f4(ge + off[v], gn);
            } else {
              accA[v] = f4_add(accA[v], gn);
              if (act[v]) red_add_f4(ge + off[v], f4_scale(gn, -1.f));
            }
          }
        }
      }
    }
    // h += gp + A ; t -= gp + B ; r += gp + A + B
    float* gh = p.ent_grad + (size_t)h * p.stride;
    float* gt = p.ent_grad + (size_t)t * p.stride;
    float* gr = rel_grad_replica(p) + (size_t)r * p.stride;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      if (act[v]) {
        const float4 a = f4_add(gp[v], accA[v]);
        const float4 b = f4_add(gp[v], accB[v]);
        red_add_f4(gh + off[v], a);
        red_add_f4(gt + off[v], f4_scale(b, -1.f));
        red_add_f4(gr + off[v], f4_add(a, accB[v]));
      }
    }
    if (lane < p.K) mark_touched(p.ent_touched, e);
    if (lane == 0) {
      mark_touched(p.ent_touched, h);
      mark_touched(p.ent_touched, t);
      mark_touched(p.rel_touched, r);
    }
  }
  block_loss_commit(loss_local, p.loss);
}

// --------------------------------------------------------------------------------------------
// variant 1: TMA path.  Per warp: two input stages of (3+K) rows filled by cp.async.bulk, two
// output stages drained by cp.reduce.async.bulk.add.f32.
// --------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
// global -> shared bulk copy, completion reported as bytes on the mbarrier (SASS UBLKCP)
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes,
                                         uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(dst),
      "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}
// shared -> global bulk reduction, element-wise fp32 add performed at L2 (SASS UBLKRED)
__device__ __forceinline__ void bulk_red_add_s2g(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(
                   dst),
               "r"(src), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

constexpr int kTmaWarps = 4;
constexpr int kTmaThreads = kTmaWarps * 32;

// dynamic smem per warp: in[2][rows][stride] + out[2][rows][stride] floats, rows = 3 + K
template <int NV>
__global__ void __launch_bounds__(kTmaThreads) rel_fused_tma_kernel(const RelStepParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t s_bar[kTmaWarps][2];
  __shared__ int32_t s_pick_all[kTmaWarps][32];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  volatile int32_t* s_pick = s_pick_all[wib];
  const int gwarp = blockIdx.x * kTmaWarps + wib;
  const int nwarps = gridDim.x * kTmaWarps;
  const int n = p.len1 + p.len2;
  const int rows = 3 + p.K;
  const uint32_t row_bytes = (uint32_t)p.stride * 4u;
  const uint32_t data_bytes = (uint32_t)p.nchunk * 16u;  // bytes of a row that carry data
  float* warp_base = reinterpret_cast<float*>(smem_raw) + (size_t)wib * 4 * rows * p.stride;
  const size_t stage_floats = (size_t)rows * p.stride;  // in stages at 0,1; out stages at 2,3
  const uint32_t bar0 = smem_u32(&s_bar[wib][0]);
  if (lane == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar0 + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  bool act[NV];
  int off[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    act[v] = (lane + 32 * v) < p.nchunk;
    off[v] = (lane + 32 * v) * 4;
  }
  float loss_local = 0.f;
  uint32_t phase_bits = 0u;

  struct Work {
    int32_t h, r, t, e;
    uint32_t side;
  };
  // lane L of the warp owns row L of a stage: 0=h 1=r 2=t 3+j=e_j
  auto row_owner = [&](const Work& w, int32_t& row, bool& is_rel) {
    const int32_t ej = __shfl_sync(0xffffffffu, w.e, (lane - 3) & 31);
    is_rel = (lane == 1);
    row = (lane == 0) ? w.h : (lane == 1) ? w.r : (lane == 2) ? w.t : ej;
  };
  auto issue = [&](int slot, int i, Work& w) {
    fetch_work(p, i, lane, s_pick, w.h, w.r, w.t, w.e, w.side);
    const uint32_t bar = bar0 + 8u * slot;
    if (lane == 0) mbar_expect_tx(bar, (uint32_t)rows * data_bytes);
    __syncwarp();
    int32_t row;
    bool is_rel;
    row_owner(w, row, is_rel);
    if (lane < rows) {
      const float* src = (is_rel ? p.rel_var : p.ent_var) + (size_t)row * p.stride;
      bulk_g2s(smem_u32(warp_base + slot * stage_floats + (size_t)lane * p.stride), src, data_bytes,
               bar);
    }
  };

  int i = gwarp;
  int slot = 0;
  Work cur{}, nxt{};
  if (i < n) issue(0, i, cur);
  for (; i < n; i += nwarps, slot ^= 1, cur = nxt) {
    const int inext = i + nwarps;
    if (inext < n) issue(slot ^ 1, inext, nxt);  // prefetch the next positive's rows
    mbar_wait(bar0 + 8u * slot, (phase_bits >> slot) & 1u);
    phase_bits ^= 1u << slot;
    const float* in = warp_base + slot * stage_floats;
    float* out = warp_base + (2 + slot) * stage_floats;
    // the bulk reductions that read this output stage two iterations ago must have drained
    bulk_wait_read<1>();
    __syncwarp();

    float4 xh[NV], xr[NV], xt[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      xh[v] = act[v] ? *reinterpret_cast<const float4*>(in + off[v]) : f4_zero();
      xr[v] = act[v] ? *reinterpret_cast<const float4*>(in + p.stride + off[v]) : f4_zero();
      xt[v] = act[v] ? *reinterpret_cast<const float4*>(in + 2 * p.stride + off[v]) : f4_zero();
    }
    float sh = 0.f, sr = 0.f, st = 0.f;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      sh += dot4(xh[v], xh[v]);
      sr += dot4(xr[v], xr[v]);
      st += dot4(xt[v], xt[v]);
    }
    warp_sum3(sh, sr, st);
    const float ih = p.ent_norm ? rsqrtf(fmaxf(sh, kNormEps)) : 1.f;
    const float ir = p.rel_norm ? rsqrtf(fmaxf(sr, kNormEps)) : 1.f;
    const float it = p.ent_norm ? rsqrtf(fmaxf(st, kNormEps)) : 1.f;
    float4 bt[NV], bh[NV], gp[NV], accA[NV], accB[NV];
    float sp = 0.f;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const float4 rr = f4_scale(xr[v], ir);
      const float4 tt = f4_scale(xt[v], it);
      bt[v] = f4_fma(xh[v], ih, rr);
      bh[v] = f4_sub(rr, tt);
      gp[v] = f4_sub(bt[v], tt);
      sp += dot4(gp[v], gp[v]);
      accA[v] = f4_zero();
      accB[v] = f4_zero();
    }
    sp = warp_sum(sp);
    {
      float lpos, sg;
      softplus_sigmoid(sp, lpos, sg);
      const float wgt = (p.w ? __ldg(p.w + i) : 1.f) * p.pos_scale;
      loss_local += wgt * lpos;
      const float cp = 2.f * sg * wgt;
#pragma unroll
      for (int v = 0; v < NV; ++v) gp[v] = f4_scale(gp[v], cp);
    }
    for (int j = 0; j < p.K; ++j) {
      const float* xrow = in + (size_t)(3 + j) * p.stride;
      float4 xe[NV];
      float se = 0.f;
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        xe[v] = act[v] ? *reinterpret_cast<const float4*>(xrow + off[v]) : f4_zero();
        se += dot4(xe[v], xe[v]);
      }
      se = warp_sum(se);
      const float ie = p.ent_norm ? rsqrtf(fmaxf(se, kNormEps)) : 1.f;
      const bool hs = (cur.side >> j) & 1u;
      float4 nd[NV];
      float sn = 0.f;
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        nd[v] = hs ? f4_fma(xe[v], ie, bh[v]) : f4_fma(xe[v], -ie, bt[v]);
        sn += dot4(nd[v], nd[v]);
      }
      sn = warp_sum(sn);
      float lneg, sg;
      softplus_sigmoid(-sn, lneg, sg);
      loss_local += lneg;
      const float cn = -2.f * sg;
      float* orow = out + (size_t)(3 + j) * p.stride;
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const float4 gn = f4_scale(nd[v], cn);
        if (hs) {
          accB[v] = f4_add(accB[v], gn);
          if (act[v]) *reinterpret_cast<float4*>(orow + off[v]) = gn;
        } else {
          accA[v] = f4_add(accA[v], gn);
          if (act[v]) *reinterpret_cast<float4*>(orow + off[v]) = f4_scale(gn, -1.f);
        }
      }
    }
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      if (act[v]) {
        const float4 a = f4_add(gp[v], accA[v]);
        const float4 b = f4_add(gp[v], accB[v]);
        *reinterpret_cast<float4*>(out + off[v]) = a;
        *reinterpret_cast<float4*>(out + p.stride + off[v]) = f4_add(a, accB[v]);
        *reinterpret_cast<float4*>(out + 2 * p.stride + off[v]) = f4_scale(b, -1.f);
      }
    }
    fence_async_smem();  // make the generic-proxy stores visible to the TMA engine
    __syncwarp();
    {
      int32_t row;
      bool is_rel;
      row_owner(cur, row, is_rel);
      if (lane < rows) {
        float* dst = (is_rel ? rel_grad_replica(p) : p.ent_grad) + (size_t)row * p.stride;
        bulk_red_add_s2g(dst, smem_u32(out + (size_t)lane * p.stride), data_bytes);
        mark_touched(is_rel ? p.rel_touched : p.ent_touched, row);
      }
      bulk_commit();
    }
  }
  bulk_wait_read<0>();
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  __shared__ float s_loss[kTmaWarps];
  if (lane == 0) s_loss[wib] = loss_local;
  __syncthreads();
  if (threadIdx.x == 0) {
    double acc = 0.0;
#pragma unroll
    for (int q = 0; q < kTmaWarps; ++q) acc += (double)s_loss[q];
    if (acc != 0.0) atomicAdd(p.loss, acc);
  }
}

// --------------------------------------------------------------------------------------------
// host side
// --------------------------------------------------------------------------------------------
static int validate_tables(const mke_table_t* ent, const mke_table_t* rel) {
  MKE_CHECK_ARG(ent && rel, "null table");
  MKE_CHECK_ARG(ent->var && ent->grad, "entity table needs var/grad");
  MKE_CHECK_ARG(rel->var && rel->grad, "relation table needs var/grad");
  MKE_CHECK_ARG(ent->stride == rel->stride && ent->dim == rel->dim,
                "fused relation kernel needs equal dim/stride for both tables (%d/%d vs %d/%d)",
                ent->dim, ent->stride, rel->dim, rel->stride);
  MKE_CHECK_ARG(ent->stride % 4 == 0 && ent->dim <= ent->stride && ent->dim > 0, "bad stride/dim");
  MKE_CHECK_ARG(ent->stride <= 256, "fused relation kernel supports stride <= 256 floats");
  MKE_CHECK_ARG(ent->grad_replicas <= 1, "the entity table of the fused relation kernel takes no gradient replicas");
  MKE_CHECK_ARG(rel->n_shards <= 1, "the relation table is replicated, not sharded");
  if (ent->n_shards > 1) {
    MKE_CHECK_ARG(ent->n_shards == 2 || ent->n_shards == 4 || ent->n_shards == 8, "n_shards must be 2, 4 or 8");
    MKE_CHECK_ARG(ent->shard_rank >= 0 && ent->shard_rank < ent->n_shards, "bad shard_rank");
    MKE_CHECK_ARG(ent->shard_split >= 0 && ent->shard_split <= ent->rows, "bad shard_split");
    for (int k = 0; k < ent->n_shards; ++k)
      MKE_CHECK_ARG(ent->peer_var[k] && ent->peer_grad[k], "peer pointer %d of the sharded entity table is null", k);
  }
  return 0;
}

template <typename Kern>
static int grid_for(Kern kern, int threads, size_t smem, int warps_per_block, int n) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem) != cudaSuccess ||
      per_sm < 1)
    per_sm = 1;
  const int full = sm_count() * per_sm;
  const int need = (n + warps_per_block - 1) / warps_per_block;
  return need < full ? (need < 1 ? 1 : need) : full;
}

static int launch_rel(RelStepParams& p, int variant, cudaStream_t stream) {
  static const int dbg = getenv("MKE_DEBUG_SKIP") ? atoi(getenv("MKE_DEBUG_SKIP")) : 0;
  p.dbg = dbg;
  // debug: MKE_TRACE=<file> dumps per-warp milestone timestamps of launch number MKE_TRACE_AT
  static const char* trace_path = getenv("MKE_TRACE");
  static const int trace_at = getenv("MKE_TRACE_AT") ? atoi(getenv("MKE_TRACE_AT")) : 50;
  static int trace_calls = 0;
  p.trace = nullptr;
  unsigned long long* trace_buf = nullptr;
  const size_t trace_words = (size_t)1 << 20;
  if (trace_path != nullptr && trace_calls++ == trace_at) {
    if (cudaMalloc(&trace_buf, trace_words * 8) == cudaSuccess) {
      cudaMemsetAsync(trace_buf, 0, trace_words * 8, stream);
      p.trace = trace_buf;
    }
  }
  struct TraceDump {
    unsigned long long* buf; size_t words; const char* path; cudaStream_t stream;
    ~TraceDump() {
      if (!buf) return;
      cudaStreamSynchronize(stream);
      unsigned long long* host = (unsigned long long*)malloc(words * 8);
      cudaMemcpy(host, buf, words * 8, cudaMemcpyDeviceToHost);
      if (FILE* f = fopen(path, "wb")) { fwrite(host, 8, words, f); fclose(f); }
      free(host);
      cudaFree(buf);
    }
  } trace_dump{trace_buf, trace_words, trace_path, stream};
  const int n = p.len1 + p.len2;
  if (n <= 0) return 0;
  const int nv = (p.nchunk + 31) / 32;
  if (variant == 3) {  // persistent row-stream schedule; launch shapes it does not cover use variant 0
    const int rc = launch_rel_q8p(p, 0, stream);
    if (rc <= 0) return rc;
    variant = 0;
  }
  if (variant == 0) {
    const int rc = launch_rel_q8(p, stream);
    if (rc <= 0) return rc;  // launched (0) or failed (<0); 1 = no instantiation for this stride
  }
  MKE_CHECK_ARG(!p.sharded, "row-sharded entity tables need the quarter-warp kernel (variant 0, stride 32/64/80/104/128)");
  MKE_CHECK_ARG(p.neg_valid == nullptr && p.pos_own_lo <= 0 && p.pos_own_hi >= n,
                "ownership masks need the quarter-warp kernels (variant 0 or 3, stride 32/64/80/104/128)");
  if (variant == 1) {
    const size_t smem = (size_t)kTmaWarps * 4 * (3 + p.K) * p.stride * sizeof(float);
    if (smem <= 200 * 1024) {
      if (nv == 1) {
        auto kern = rel_fused_tma_kernel<1>;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(tma<1>)");
        kern<<<grid_for(kern, kTmaThreads, smem, kTmaWarps, n), kTmaThreads, smem, stream>>>(p);
      } else {
        auto kern = rel_fused_tma_kernel<2>;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(tma<2>)");
        kern<<<grid_for(kern, kTmaThreads, smem, kTmaWarps, n), kTmaThreads, smem, stream>>>(p);
      }
      MKE_CHECK_LAUNCH("rel_fused_tma_kernel");
      return 0;
    }
    // stage too large for shared memory: the register path handles any K
  }
  if (nv == 1) {
    auto kern = rel_fused_ldg_kernel<1, 5>;
    kern<<<grid_for(kern, kRelThreads, 0, kRelWarps, n), kRelThreads, 0, stream>>>(p);
  } else {
    auto kern = rel_fused_ldg_kernel<2, 4>;
    kern<<<grid_for(kern, kRelThreads, 0, kRelWarps, n), kRelThreads, 0, stream>>>(p);
  }
  MKE_CHECK_LAUNCH("rel_fused_ldg_kernel");
  return 0;
}

void fill_tables(RelStepParams& p, const mke_table_t* ent, const mke_table_t* rel) {
  p.ent_var = ent->var;
  p.ent_grad = ent->grad;
  p.ent_touched = ent->touched;
  p.rel_var = rel->var;
  p.rel_grad = rel->grad;
  p.sharded = ent->n_shards > 1 ? 1 : 0;
  p.smap = shard_map(ent);
  for (int k = 0; k < MKE_MAX_SHARDS; ++k) {
    const bool on = ent->n_shards > 1 && k < ent->n_shards;
    p.sh.var[k] = on ? ent->peer_var[k] : nullptr;
    p.sh.grad[k] = on ? ent->peer_grad[k] : nullptr;
    p.sh.touched[k] = on ? ent->peer_touched[k] : nullptr;
  }
  p.rel_rep = rel->grad_replicas > 1 ? rel->grad_replicas : 1;
  p.rel_rep_floats = (size_t)rel->rows * (size_t)rel->stride;
  p.rel_touched = rel->touched;
  p.stride = ent->stride;
  p.nchunk = (ent->dim + 3) / 4;
  p.ent_norm = ent->normalised;
  p.rel_norm = rel->normalised;
}

}  // namespace mke

using namespace mke;

extern "C" int mke_rel_step_sampled(const mke_table_t* ent, const mke_table_t* rel,
                                    const int32_t* pos1, int32_t len1, const mke_kg_sampler_t* kg1,
                                    const int32_t* pos2, int32_t len2, const mke_kg_sampler_t* kg2,
                                    int32_t K, uint64_t seed, uint64_t step, const float* w_or_null,
                                    float pos_scale, double* loss_accum, int32_t* neg_out_or_null,
                                    int32_t variant, mke_stream_t stream) {
  if (int rc = validate_tables(ent, rel)) return rc;
  MKE_CHECK_ARG(K >= 0 && K <= MKE_MAX_NEG, "K=%d outside [0,%d]", K, MKE_MAX_NEG);
  MKE_CHECK_ARG(len1 >= 0 && len2 >= 0, "negative batch length");
  MKE_CHECK_ARG((uint64_t)len1 + (uint64_t)len2 < (1ull << 31), "batch too large");
  MKE_CHECK_ARG(len1 == 0 || (pos1 && (K == 0 || kg1)), "kg1 slice needs positives and a sampler");
  MKE_CHECK_ARG(len2 == 0 || (pos2 && (K == 0 || kg2)), "kg2 slice needs positives and a sampler");
  MKE_CHECK_ARG(loss_accum, "loss_accum is null");
  RelStepParams p{};
  p.pos_own_lo = 0;
  p.pos_own_hi = 0x7fffffff;
  fill_tables(p, ent, rel);
  p.pos1 = pos1;
  p.len1 = len1;
  p.pos2 = pos2;
  p.len2 = len2;
  if (kg1) p.kg1 = *kg1;
  if (kg2) p.kg2 = *kg2;
  if (K > 0) {
    for (const mke_kg_sampler_t* kg : {len1 ? kg1 : nullptr, len2 ? kg2 : nullptr}) {
      if (!kg) continue;
      MKE_CHECK_ARG(kg->n_entities >= K, "KG has fewer entities (%d) than K=%d", kg->n_entities, K);
      MKE_CHECK_ARG(!kg->neighbours || kg->n_neighbours >= K, "n_neighbours < K");
      MKE_CHECK_ARG(!kg->set.slots || (kg->set.capacity & (kg->set.capacity - 1)) == 0,
                    "triple-set capacity must be a power of two");
    }
  }
  p.K = K;
  p.sampled = 1;
  p.skey = stream_key(seed, step);
  p.w = w_or_null;
  p.pos_scale = pos_scale;
  p.loss = loss_accum;
  p.neg_out = neg_out_or_null;
  return launch_rel(p, variant, (cudaStream_t)stream);
}

extern "C" int mke_rel_step_structured2(const mke_table_t* ent, const mke_table_t* rel,
                                        const int32_t* pos1, int32_t len1, const int32_t* pos2,
                                        int32_t len2, int32_t K, const int32_t* neg_ent,
                                        const uint32_t* neg_side, const float* w_or_null,
                                        float pos_scale, double* loss_accum, int32_t variant,
                                        mke_stream_t stream) {
  return mke_rel_step_structured3(ent, rel, pos1, len1, pos2, len2, K, neg_ent, neg_side, nullptr, 0, 0x7fffffff,
                                  w_or_null, pos_scale, loss_accum, variant, stream);
}

extern "C" int mke_rel_step_structured3(const mke_table_t* ent, const mke_table_t* rel,
                                        const int32_t* pos1, int32_t len1, const int32_t* pos2,
                                        int32_t len2, int32_t K, const int32_t* neg_ent,
                                        const uint32_t* neg_side, const uint32_t* neg_valid_or_null,
                                        int32_t pos_own_lo, int32_t pos_own_hi, const float* w_or_null,
                                        float pos_scale, double* loss_accum, int32_t variant,
                                        mke_stream_t stream) {
  return mke_rel_step_structured4(ent, rel, pos1, len1, pos2, len2, K, neg_ent, neg_side, neg_valid_or_null, 0,
                                  pos_own_lo, pos_own_hi, w_or_null, pos_scale, loss_accum, variant, stream);
}

extern "C" int mke_rel_step_structured4(const mke_table_t* ent, const mke_table_t* rel,
                                        const int32_t* pos1, int32_t len1, const int32_t* pos2,
                                        int32_t len2, int32_t K, const int32_t* neg_ent,
                                        const uint32_t* neg_side, const uint32_t* neg_valid_or_null,
                                        int32_t compact, int32_t pos_own_lo, int32_t pos_own_hi,
                                        const float* w_or_null, float pos_scale, double* loss_accum,
                                        int32_t variant, mke_stream_t stream) {
  if (int rc = validate_tables(ent, rel)) return rc;
  MKE_CHECK_ARG(pos_own_lo >= 0 && pos_own_hi >= pos_own_lo, "bad owner range [%d, %d)", pos_own_lo, pos_own_hi);
  MKE_CHECK_ARG(K >= 0 && K <= MKE_MAX_NEG, "K=%d outside [0,%d]", K, MKE_MAX_NEG);
  MKE_CHECK_ARG(len1 >= 0 && len2 >= 0, "negative batch length");
  MKE_CHECK_ARG((uint64_t)len1 + (uint64_t)len2 < (1ull << 31), "batch too large");
  MKE_CHECK_ARG((len1 == 0 || pos1) && (len2 == 0 || pos2), "pos is null");
  MKE_CHECK_ARG(K == 0 || len1 + len2 == 0 || (neg_ent && neg_side), "neg_ent/neg_side are null");
  MKE_CHECK_ARG(loss_accum, "loss_accum is null");
  RelStepParams p{};
  p.pos_own_lo = 0;
  p.pos_own_hi = 0x7fffffff;
  fill_tables(p, ent, rel);
  p.pos1 = pos1;
  p.len1 = len1;
  p.pos2 = pos2;
  p.len2 = len2;
  p.K = K;
  p.sampled = 0;
  p.neg_ent = neg_ent;
  p.neg_side = neg_side;
  p.neg_valid = neg_valid_or_null;
  p.neg_compact = (compact != 0 && neg_valid_or_null != nullptr) ? 1 : 0;
  p.pos_own_lo = pos_own_lo;
  p.pos_own_hi = pos_own_hi;
  p.w = w_or_null;
  p.pos_scale = pos_scale;
  p.loss = loss_accum;
  p.neg_out = nullptr;
  return launch_rel(p, variant, (cudaStream_t)stream);
}

extern "C" int mke_rel_step_structured(const mke_table_t* ent, const mke_table_t* rel,
                                       const int32_t* pos, int32_t n, int32_t K,
                                       const int32_t* neg_ent, const uint32_t* neg_side,
                                       const float* w_or_null, float pos_scale, double* loss_accum,
                                       int32_t variant, mke_stream_t stream) {
  return mke_rel_step_structured2(ent, rel, pos, n, nullptr, 0, K, neg_ent, neg_side, w_or_null,
                                  pos_scale, loss_accum, variant, stream);
}
