// Generic TransE-style scored batch (reference API form: one index vector per role).
// One warp per triple; handles any stride, independent tables for head / middle / tail,
// optional per-triple weights, constant (grad == NULL) and un-normalised tables.
// Replaces losses.py:4-50 plus the embedding_lookup gathers and their backward
// (MultiKE_model.py:123-131, 164-170, 194-201).
#include "mke_common.cuh"

namespace mke {

struct TripleParams {
  const float *vh, *vm, *vt;
  float *gh, *gm, *gt;
  int rh, rm, rt;           // gradient replicas per table (>= 1)
  size_t fh_, fm_, ft_;     // floats per replica
  uint8_t *fh, *fm, *ft;
  int sh, sm, st;  // strides
  int nh, nm, nt;  // normalised flags
  int nchunk;      // ceil(dim/4)
  const int32_t *ih, *im, *it;
  int n;
  const float* w;
  int negative;
  float scale;
  double* loss;
  float* score;
};

constexpr int kTripleThreads = 256;
constexpr int kTripleWarps = kTripleThreads / 32;
constexpr int kMaxNV = 8;  // rows up to 8*32*4 = 1024 floats

template <int NV>
__global__ void __launch_bounds__(kTripleThreads) triple_fwd_bwd_kernel(const TripleParams p) {
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int gwarp = blockIdx.x * kTripleWarps + wib;
  const int nwarps = gridDim.x * kTripleWarps;
  float loss_local = 0.f;
  float* const gh = p.gh ? p.gh + (size_t)(blockIdx.x % (unsigned)p.rh) * p.fh_ : nullptr;
  float* const gm = p.gm ? p.gm + (size_t)(blockIdx.x % (unsigned)p.rm) * p.fm_ : nullptr;
  float* const gt = p.gt ? p.gt + (size_t)(blockIdx.x % (unsigned)p.rt) * p.ft_ : nullptr;
  for (int i = gwarp; i < p.n; i += nwarps) {
    const int32_t h = __ldg(p.ih + i), m = __ldg(p.im + i), t = __ldg(p.it + i);
    const float* ph = p.vh + (size_t)h * p.sh;
    const float* pm = p.vm + (size_t)m * p.sm;
    const float* pt = p.vt + (size_t)t * p.st;
    float4 xh[NV], xm[NV], xt[NV];
    float ssh = 0.f, ssm = 0.f, sst = 0.f;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int c = lane + 32 * v;
      const bool a = c < p.nchunk;
      xh[v] = a ? ldg_f4(ph + 4 * c) : f4_zero();
      xm[v] = a ? ldg_f4(pm + 4 * c) : f4_zero();
      xt[v] = a ? ldg_f4(pt + 4 * c) : f4_zero();
      ssh += dot4(xh[v], xh[v]);
      ssm += dot4(xm[v], xm[v]);
      sst += dot4(xt[v], xt[v]);
    }
    warp_sum3(ssh, ssm, sst);
    const float ih = p.nh ? rsqrtf(fmaxf(ssh, kNormEps)) : 1.f;
    const float im = p.nm ? rsqrtf(fmaxf(ssm, kNormEps)) : 1.f;
    const float it = p.nt ? rsqrtf(fmaxf(sst, kNormEps)) : 1.f;
    float4 d[NV];
    float s = 0.f;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      d[v] = f4_sub(f4_fma(xh[v], ih, f4_scale(xm[v], im)), f4_scale(xt[v], it));
      s += dot4(d[v], d[v]);
    }
    s = warp_sum(s);
    if (p.score != nullptr && lane == 0) p.score[i] = -s;
    const float x = p.negative ? -s : s;
    const float ex = expf(x);
    const float onep = 1.f + ex;
    const float wgt = (p.w ? __ldg(p.w + i) : 1.f) * p.scale;
    loss_local += wgt * logf(onep);
    // d loss / d distance = +-2 sigma(x) * distance
    const float c = (p.negative ? -2.f : 2.f) * (ex / onep) * wgt;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int cidx = lane + 32 * v;
      if (cidx < p.nchunk) {
        const float4 g = f4_scale(d[v], c);
        if (gh) red_add_f4(gh + (size_t)h * p.sh + 4 * cidx, g);
        if (gm) red_add_f4(gm + (size_t)m * p.sm + 4 * cidx, g);
        if (gt) red_add_f4(gt + (size_t)t * p.st + 4 * cidx, f4_scale(g, -1.f));
      }
    }
    if (lane == 0) {
      mark_touched(p.fh, h);
      mark_touched(p.fm, m);
      mark_touched(p.ft, t);
    }
  }
  __shared__ float s_loss[kTripleWarps];
  if (lane == 0) s_loss[wib] = loss_local;
  __syncthreads();
  if (threadIdx.x == 0 && p.loss != nullptr) {
    double acc = 0.0;
#pragma unroll
    for (int q = 0; q < kTripleWarps; ++q) acc += (double)s_loss[q];
    if (acc != 0.0) atomicAdd(p.loss, acc);
  }
}

template <int NV>
static int launch_triple(const TripleParams& p, cudaStream_t stream) {
  auto kern = triple_fwd_bwd_kernel<NV>;
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kTripleThreads, 0) != cudaSuccess ||
      per_sm < 1)
    per_sm = 1;
  const int full = sm_count() * per_sm;
  int need = (p.n + kTripleWarps - 1) / kTripleWarps;
  if (need > full) need = full;
  kern<<<need, kTripleThreads, 0, stream>>>(p);
  MKE_CHECK_LAUNCH("triple_fwd_bwd_kernel");
  return 0;
}

}  // namespace mke

using namespace mke;

extern "C" int mke_triple_fwd_bwd(const mke_table_t* head, const mke_table_t* mid,
                                  const mke_table_t* tail, const int32_t* ih, const int32_t* im,
                                  const int32_t* it, int32_t n, const float* w_or_null,
                                  int32_t negative, float scale, double* loss_accum,
                                  float* score_out_or_null, mke_stream_t stream) {
  MKE_CHECK_ARG(head && mid && tail, "null table");
  MKE_CHECK_ARG(head->var && mid->var && tail->var, "table without var");
  MKE_CHECK_ARG(head->dim == mid->dim && mid->dim == tail->dim && head->dim > 0,
                "tables disagree on dim (%d,%d,%d)", head->dim, mid->dim, tail->dim);
  for (const mke_table_t* tb : {head, mid, tail}) {
    MKE_CHECK_ARG(tb->stride % 4 == 0 && tb->dim <= tb->stride, "bad stride %d for dim %d",
                  tb->stride, tb->dim);
  }
  MKE_CHECK_ARG(n >= 0, "negative n");
  if (n == 0) return 0;
  MKE_CHECK_ARG(ih && im && it, "null index vector");
  TripleParams p{};
  p.vh = head->var; p.vm = mid->var; p.vt = tail->var;
  p.gh = head->grad; p.gm = mid->grad; p.gt = tail->grad;
  p.rh = head->grad_replicas > 1 ? head->grad_replicas : 1;
  p.rm = mid->grad_replicas > 1 ? mid->grad_replicas : 1;
  p.rt = tail->grad_replicas > 1 ? tail->grad_replicas : 1;
  p.fh_ = (size_t)head->rows * head->stride;
  p.fm_ = (size_t)mid->rows * mid->stride;
  p.ft_ = (size_t)tail->rows * tail->stride;
  p.fh = head->grad ? head->touched : nullptr;
  p.fm = mid->grad ? mid->touched : nullptr;
  p.ft = tail->grad ? tail->touched : nullptr;
  p.sh = head->stride; p.sm = mid->stride; p.st = tail->stride;
  p.nh = head->normalised; p.nm = mid->normalised; p.nt = tail->normalised;
  p.nchunk = (head->dim + 3) / 4;
  p.ih = ih; p.im = im; p.it = it;
  p.n = n;
  p.w = w_or_null;
  p.negative = negative ? 1 : 0;
  p.scale = scale;
  p.loss = loss_accum;
  p.score = score_out_or_null;
  const int nv = (p.nchunk + 31) / 32;
  MKE_CHECK_ARG(nv <= kMaxNV, "dim %d too large (max %d)", head->dim, kMaxNV * 128);
  cudaStream_t s = (cudaStream_t)stream;
  switch (nv) {
    case 1: return launch_triple<1>(p, s);
    case 2: return launch_triple<2>(p, s);
    case 3: case 4: return launch_triple<4>(p, s);
    default: return launch_triple<8>(p, s);
  }
}
