// losses.py on ALREADY GATHERED rows (the reference's free functions take [n, dim] matrices):
// forward value and the gradient with respect to every input row in one pass.  One warp per row.
// Used by multike_b200/refapi/losses.py (torch.autograd.Function wrappers); the training drivers
// use the fused index-based kernels instead.
#include "mke_common.cuh"

namespace mke {

constexpr int kDenseThreads = 256;
constexpr int kDenseWarps = kDenseThreads / 32;

// MODE 0: logistic TransE term  w * log(1 + exp(+-|h + m - t|^2))   (losses.py:4-50)
// MODE 1: squared distance      |a - b|^2                            (losses.py:66-69), m unused
template <int MODE>
__global__ void __launch_bounds__(kDenseThreads)
    dense_loss_kernel(const float* __restrict__ H, const float* __restrict__ M, const float* __restrict__ T,
                      int n, int dim, int ld, const float* __restrict__ w, int negative, float scale,
                      double* __restrict__ loss, float* __restrict__ gH, float* __restrict__ gM,
                      float* __restrict__ gT) {
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  float loss_local = 0.f;
  for (int i = blockIdx.x * kDenseWarps + wib; i < n; i += gridDim.x * kDenseWarps) {
    const size_t o = (size_t)i * ld;
    float s = 0.f;
    for (int c = lane; c < dim; c += 32) {
      const float d = (MODE == 0) ? (H[o + c] + M[o + c] - T[o + c]) : (H[o + c] - T[o + c]);
      s = fmaf(d, d, s);
    }
    s = warp_sum(s);
    float coef;
    if (MODE == 0) {
      const float x = negative ? -s : s;
      const float ex = expf(x), onep = 1.f + ex;
      const float wgt = (w ? __ldg(w + i) : 1.f) * scale;
      loss_local += wgt * logf(onep);
      coef = (negative ? -2.f : 2.f) * (ex / onep) * wgt;
    } else {
      loss_local += scale * s;
      coef = 2.f * scale;
    }
    for (int c = lane; c < dim; c += 32) {
      const float d = (MODE == 0) ? (H[o + c] + M[o + c] - T[o + c]) : (H[o + c] - T[o + c]);
      const float g = coef * d;
      if (gH) gH[o + c] = g;
      if (MODE == 0 && gM) gM[o + c] = g;
      if (gT) gT[o + c] = -g;
    }
  }
  __shared__ float s_loss[kDenseWarps];
  if (lane == 0) s_loss[wib] = loss_local;
  __syncthreads();
  if (threadIdx.x == 0 && loss != nullptr) {
    double a = 0.0;
#pragma unroll
    for (int q = 0; q < kDenseWarps; ++q) a += (double)s_loss[q];
    if (a != 0.0) atomicAdd(loss, a);
  }
}

template <int MODE>
static int launch_dense(const float* H, const float* M, const float* T, int n, int dim, int ld, const float* w,
                        int negative, float scale, double* loss, float* gH, float* gM, float* gT,
                        cudaStream_t stream) {
  int blocks = (n + kDenseWarps - 1) / kDenseWarps;
  const int full = sm_count() * 8;
  if (blocks > full) blocks = full;
  dense_loss_kernel<MODE><<<blocks, kDenseThreads, 0, stream>>>(H, M, T, n, dim, ld, w, negative, scale, loss, gH,
                                                                 gM, gT);
  MKE_CHECK_LAUNCH("dense_loss_kernel");
  return 0;
}

}  // namespace mke

using namespace mke;

extern "C" int mke_dense_logistic_fwd_bwd(const float* H, const float* M, const float* T, int32_t n,
                                          int32_t dim, int32_t ld, const float* w_or_null,
                                          int32_t negative, float scale, double* loss_accum, float* gH,
                                          float* gM, float* gT, mke_stream_t stream) {
  MKE_CHECK_ARG(n >= 0 && dim > 0 && ld >= dim, "bad shape n=%d dim=%d ld=%d", n, dim, ld);
  if (n == 0) return 0;
  MKE_CHECK_ARG(H && M && T, "null input matrix");
  return launch_dense<0>(H, M, T, n, dim, ld, w_or_null, negative ? 1 : 0, scale, loss_accum, gH, gM, gT,
                         (cudaStream_t)stream);
}

extern "C" int mke_dense_sqdist_fwd_bwd(const float* A, const float* B, int32_t n, int32_t dim, int32_t ld,
                                        float scale, double* loss_accum, float* gA, float* gB,
                                        mke_stream_t stream) {
  MKE_CHECK_ARG(n >= 0 && dim > 0 && ld >= dim, "bad shape n=%d dim=%d ld=%d", n, dim, ld);
  if (n == 0) return 0;
  MKE_CHECK_ARG(A && B, "null input matrix");
  return launch_dense<1>(A, nullptr, B, n, dim, ld, nullptr, 0, scale, loss_accum, gA, nullptr, gB,
                         (cudaStream_t)stream);
}
