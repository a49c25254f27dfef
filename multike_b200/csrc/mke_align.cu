// Cross-view alignment step of ITC (MultiKE_model.py:225-239, losses.py:66-69):
//   loss = scale * ( w_name |F_i - N_i|^2 + |F_i - R_i|^2 + |F_i - A_i|^2 )   summed over a batch
// of entity ids i, with F = ent_embeds (shared), N = name_embeds (constant), R = rv_ent_embeds,
// A = av_ent_embeds -- each trainable table read through l2_normalize(.,1) -- plus its backward:
// gradient rows are accumulated into the three tables' grad buffers (phase 2 = the usual
// mke_rows_apply_adagrad with the ITC learning rate and this graph's accumulator slots).
// One warp per entity; any stride.
#include "mke_common.cuh"

namespace mke {

constexpr int kAlignThreads = 256;
constexpr int kAlignWarps = kAlignThreads / 32;
constexpr int kAlignNV = 8;  // rows up to 1024 floats

struct AlignTable {
  const float* var;
  float* grad;
  uint8_t* touched;
  int stride, normalised;
};
struct AlignParams {
  AlignTable f, n, r, a;
  int nchunk;
  const int32_t* idx;
  int count;
  float w_name, scale;
  double* loss;
};

__global__ void __launch_bounds__(kAlignThreads) align_fwd_bwd_kernel(const AlignParams p) {
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  float loss_local = 0.f;
  for (int i = blockIdx.x * kAlignWarps + wib; i < p.count; i += gridDim.x * kAlignWarps) {
    const int32_t e = __ldg(p.idx + i);
    const float* pf = p.f.var + (size_t)e * p.f.stride;
    const float* pn = p.n.var + (size_t)e * p.n.stride;
    const float* pr = p.r.var + (size_t)e * p.r.stride;
    const float* pa = p.a.var + (size_t)e * p.a.stride;
    float sf = 0.f, sn = 0.f, sr = 0.f, sa = 0.f;
    for (int c = lane; c < p.nchunk; c += 32) {
      const float4 xf = ldg_f4(pf + 4 * c), xn = ldg_f4(pn + 4 * c), xr = ldg_f4(pr + 4 * c), xa = ldg_f4(pa + 4 * c);
      sf += dot4(xf, xf);
      sn += dot4(xn, xn);
      sr += dot4(xr, xr);
      sa += dot4(xa, xa);
    }
    warp_sum3(sf, sr, sa);
    sn = warp_sum(sn);
    const float kf = p.f.normalised ? rsqrtf(fmaxf(sf, kNormEps)) : 1.f;
    const float kn = p.n.normalised ? rsqrtf(fmaxf(sn, kNormEps)) : 1.f;
    const float kr = p.r.normalised ? rsqrtf(fmaxf(sr, kNormEps)) : 1.f;
    const float ka = p.a.normalised ? rsqrtf(fmaxf(sa, kNormEps)) : 1.f;
    float l1 = 0.f, l2 = 0.f, l3 = 0.f;
    for (int c = lane; c < p.nchunk; c += 32) {  // second pass: rows are L1 hits
      const float4 F = f4_scale(ldg_f4(pf + 4 * c), kf);
      const float4 d1 = f4_sub(F, f4_scale(ldg_f4(pn + 4 * c), kn));
      const float4 d2 = f4_sub(F, f4_scale(ldg_f4(pr + 4 * c), kr));
      const float4 d3 = f4_sub(F, f4_scale(ldg_f4(pa + 4 * c), ka));
      l1 += dot4(d1, d1);
      l2 += dot4(d2, d2);
      l3 += dot4(d3, d3);
      const float s2 = 2.f * p.scale;
      const float4 gf = f4_scale(f4_add(f4_add(f4_scale(d1, p.w_name), d2), d3), s2);
      if (p.f.grad) red_add_f4(p.f.grad + (size_t)e * p.f.stride + 4 * c, gf);
      if (p.n.grad) red_add_f4(p.n.grad + (size_t)e * p.n.stride + 4 * c, f4_scale(d1, -s2 * p.w_name));
      if (p.r.grad) red_add_f4(p.r.grad + (size_t)e * p.r.stride + 4 * c, f4_scale(d2, -s2));
      if (p.a.grad) red_add_f4(p.a.grad + (size_t)e * p.a.stride + 4 * c, f4_scale(d3, -s2));
    }
    warp_sum3(l1, l2, l3);
    loss_local += p.scale * (p.w_name * l1 + l2 + l3);
    if (lane == 0) {
      if (p.f.grad) mark_touched(p.f.touched, e);
      if (p.n.grad) mark_touched(p.n.touched, e);
      if (p.r.grad) mark_touched(p.r.touched, e);
      if (p.a.grad) mark_touched(p.a.touched, e);
    }
  }
  __shared__ float s_loss[kAlignWarps];
  if (lane == 0) s_loss[wib] = loss_local;
  __syncthreads();
  if (threadIdx.x == 0 && p.loss != nullptr) {
    double acc = 0.0;
#pragma unroll
    for (int q = 0; q < kAlignWarps; ++q) acc += (double)s_loss[q];
    if (acc != 0.0) atomicAdd(p.loss, acc);
  }
}

}  // namespace mke

using namespace mke;

extern "C" int mke_align_fwd_bwd(const mke_table_t* shared, const mke_table_t* name, const mke_table_t* rv,
                                 const mke_table_t* av, const int32_t* idx, int32_t n, float name_weight,
                                 float scale, double* loss_accum, mke_stream_t stream) {
  MKE_CHECK_ARG(shared && name && rv && av, "null table");
  AlignParams p{};
  AlignTable* dst[4] = {&p.f, &p.n, &p.r, &p.a};
  const mke_table_t* src[4] = {shared, name, rv, av};
  for (int k = 0; k < 4; ++k) {
    const mke_table_t* t = src[k];
    MKE_CHECK_ARG(t->var && t->dim == shared->dim && t->stride % 4 == 0 && t->dim <= t->stride,
                  "table %d: bad var/dim/stride", k);
    MKE_CHECK_ARG(t->grad_replicas <= 1 && t->n_shards <= 1, "alignment step takes plain tables");
    *dst[k] = AlignTable{t->var, t->grad, t->grad ? t->touched : nullptr, t->stride, t->normalised};
  }
  MKE_CHECK_ARG(n >= 0, "negative n");
  if (n == 0) return 0;
  MKE_CHECK_ARG(idx, "idx is null");
  p.nchunk = (shared->dim + 3) / 4;
  MKE_CHECK_ARG(p.nchunk <= kAlignNV * 32, "dim too large");
  p.idx = idx;
  p.count = n;
  p.w_name = name_weight;
  p.scale = scale;
  p.loss = loss_accum;
  int blocks = (n + kAlignWarps - 1) / kAlignWarps;
  const int full = sm_count() * 8;
  if (blocks > full) blocks = full;
  align_fwd_bwd_kernel<<<blocks, kAlignThreads, 0, (cudaStream_t)stream>>>(p);
  MKE_CHECK_LAUNCH("align_fwd_bwd_kernel");
  return 0;
}
