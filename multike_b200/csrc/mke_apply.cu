// Phase 2: per touched row, gradient of l2_normalize + Adagrad apply + gradient/flag reset.
// Replaces the dense l2_normalize backward and tf.train.AdagradOptimizer.apply_gradients of
// MultiKE_model.py:15-31 (base/initializers.py:26 for the normalised view).  Rows whose gradient
// is identically zero are a no-op under Adagrad (acc += 0, v -= 0) and are skipped via the
// `touched` byte map written by phase 1.
#include "mke_common.cuh"

namespace mke {

constexpr int kApplyThreads = 256;
constexpr int kApplyWarps = kApplyThreads / 32;

template <int NV>
__global__ void __launch_bounds__(kApplyThreads)
    apply_adagrad_kernel(float* __restrict__ var, float* __restrict__ grad,
                         uint8_t* __restrict__ touched, float* __restrict__ acc, int rows,
                         int stride, int nchunk, int normalised, float lr) {
  const int lane = threadIdx.x & 31;
  const int gwarp = blockIdx.x * kApplyWarps + (threadIdx.x >> 5);
  const int nwarps = gridDim.x * kApplyWarps;
  for (int base = gwarp * 32; base < rows; base += nwarps * 32) {
    const int my = base + lane;
    const bool flag = (my < rows) && (touched[my] != 0);
    uint32_t m = __ballot_sync(0xffffffffu, flag);
    if (flag) touched[my] = 0;
    while (m) {
      const int row = base + (__ffs(m) - 1);
      m &= m - 1;
      float* pv = var + (size_t)row * stride;
      float* pg = grad + (size_t)row * stride;
      float* pa = acc + (size_t)row * stride;
      float4 g[NV], v[NV], a[NV];
      float ss = 0.f, vg = 0.f;
#pragma unroll
      for (int q = 0; q < NV; ++q) {
        const int c = lane + 32 * q;
        if (c < nchunk) {
          g[q] = *reinterpret_cast<const float4*>(pg + 4 * c);
          v[q] = *reinterpret_cast<const float4*>(pv + 4 * c);
          a[q] = *reinterpret_cast<const float4*>(pa + 4 * c);
        } else {
          g[q] = v[q] = f4_zero();
          a[q] = make_float4(1.f, 1.f, 1.f, 1.f);
        }
        ss += dot4(v[q], v[q]);
        vg += dot4(v[q], g[q]);
      }
      float inv = 1.f, coef = 0.f;
      if (normalised) {
        warp_sum2(ss, vg);
        // y = v * rsqrt(max(|v|^2, eps)); the max() routes no gradient to |v|^2 below eps
        inv = rsqrtf(fmaxf(ss, kNormEps));
        coef = (ss >= kNormEps) ? vg * inv * inv : 0.f;
      }
#pragma unroll
      for (int q = 0; q < NV; ++q) {
        const int c = lane + 32 * q;
        if (c < nchunk) {
          float4 gv;
          gv.x = (g[q].x - v[q].x * coef) * inv;
          gv.y = (g[q].y - v[q].y * coef) * inv;
          gv.z = (g[q].z - v[q].z * coef) * inv;
          gv.w = (g[q].w - v[q].w * coef) * inv;
          a[q].x += gv.x * gv.x;
          a[q].y += gv.y * gv.y;
          a[q].z += gv.z * gv.z;
          a[q].w += gv.w * gv.w;
          // var -= grad * lr * rsqrt(accum)   (ApplyAdagrad, no epsilon) [TF semantics]
          v[q].x -= gv.x * lr * (a[q].x > 0.f ? rsqrtf(a[q].x) : 0.f);
          v[q].y -= gv.y * lr * (a[q].y > 0.f ? rsqrtf(a[q].y) : 0.f);
          v[q].z -= gv.z * lr * (a[q].z > 0.f ? rsqrtf(a[q].z) : 0.f);
          v[q].w -= gv.w * lr * (a[q].w > 0.f ? rsqrtf(a[q].w) : 0.f);
          *reinterpret_cast<float4*>(pv + 4 * c) = v[q];
          *reinterpret_cast<float4*>(pa + 4 * c) = a[q];
          *reinterpret_cast<float4*>(pg + 4 * c) = f4_zero();
        }
      }
    }
  }
}

template <int NV>
static int launch_apply(const mke_table_t* t, float* acc, float lr, cudaStream_t stream) {
  auto kern = apply_adagrad_kernel<NV>;
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kApplyThreads, 0) != cudaSuccess ||
      per_sm < 1)
    per_sm = 1;
  const int full = sm_count() * per_sm;
  int need = ((t->rows + 31) / 32 + kApplyWarps - 1) / kApplyWarps;
  if (need > full) need = full;
  if (need < 1) need = 1;
  kern<<<need, kApplyThreads, 0, stream>>>(t->var, t->grad, t->touched, acc, t->rows, t->stride,
                                           (t->dim + 3) / 4, t->normalised, lr);
  MKE_CHECK_LAUNCH("apply_adagrad_kernel");
  return 0;
}

}  // namespace mke

using namespace mke;

extern "C" int mke_rows_apply_adagrad(const mke_table_t* table, float* acc, float lr,
                                      mke_stream_t stream) {
  MKE_CHECK_ARG(table && table->var && table->grad && table->touched && acc,
                "apply needs var/grad/touched/acc");
  MKE_CHECK_ARG(table->stride % 4 == 0 && table->dim <= table->stride && table->dim > 0,
                "bad stride/dim");
  if (table->rows <= 0) return 0;
  const int nv = ((table->dim + 3) / 4 + 31) / 32;
  MKE_CHECK_ARG(nv <= 8, "dim too large");
  cudaStream_t s = (cudaStream_t)stream;
  switch (nv) {
    case 1: return launch_apply<1>(table, acc, lr, s);
    case 2: return launch_apply<2>(table, acc, lr, s);
    case 3: case 4: return launch_apply<4>(table, acc, lr, s);
    default: return launch_apply<8>(table, acc, lr, s);
  }
}
