// Phase 2: per touched row, gradient of l2_normalize + Adagrad apply + gradient/flag reset.
// Replaces the dense l2_normalize backward and tf.train.AdagradOptimizer.apply_gradients of
// MultiKE_model.py:15-31 (base/initializers.py:26 for the normalised view).  Rows whose gradient
// is identically zero are a no-op under Adagrad (acc += 0, v -= 0) and are skipped via the
// `touched` byte map written by phase 1.
#include <cstdlib>
#include "mke_apply.cuh"

namespace mke {

constexpr int kApplyThreads = 256;
constexpr int kApplyWarps = kApplyThreads / 32;

template <int NV>
__global__ void __launch_bounds__(kApplyThreads)
    apply_adagrad_kernel(float* __restrict__ var, float* __restrict__ grad,
                         uint8_t* __restrict__ touched, float* __restrict__ acc, int rows,
                         int stride, int nchunk, int normalised, float lr, int replicas) {
  const int lane = threadIdx.x & 31;
  const int gwarp = blockIdx.x * kApplyWarps + (threadIdx.x >> 5);
  const int nwarps = gridDim.x * kApplyWarps;
  for (int base = gwarp * 32; base < rows; base += nwarps * 32) {
    const int my = base + lane;
    const bool flag = (my < rows) && (touched == nullptr || touched[my] != 0);
    uint32_t m = __ballot_sync(0xffffffffu, flag);
    if (flag && touched != nullptr) touched[my] = 0;
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
          for (int rep = 1; rep < replicas; ++rep)
            g[q] = f4_add(g[q], *reinterpret_cast<const float4*>(pg + (size_t)rep * rows * stride + 4 * c));
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
          for (int rep = 0; rep < replicas; ++rep)
            *reinterpret_cast<float4*>(pg + (size_t)rep * rows * stride + 4 * c) = f4_zero();
        }
      }
    }
  }
}

// Large tables: flag chunks of 32 rows per warp.  Small tables (no flags): one row per quarter, every row.
template <int FPL>
__device__ __forceinline__ void apply_rows(const ApplyTable& T, int gwarp, int nwarps, int lane) {
  if (T.touched == nullptr) {
    for (int r0 = gwarp * 4; r0 < T.rows; r0 += nwarps * 4) apply_row4<FPL>(T, r0, lane);
    return;
  }
  for (int base = gwarp * 32; base < T.rows; base += nwarps * 32) apply_flag_chunk<FPL>(T, base, lane);
}

// Up to two tables of equal stride in one launch (the entity and the relation table of a view):
// the first blocks_a thread blocks sweep table A, the others table B, concurrently.
template <int FPL, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
    apply_adagrad_q8_kernel(const ApplyTable A, const ApplyTable B, int blocks_a) {
  constexpr int WARPS = THREADS / 32;
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  if ((int)blockIdx.x < blocks_a)
    apply_rows<FPL>(A, blockIdx.x * WARPS + wib, blocks_a * WARPS, lane);
  else
    apply_rows<FPL>(B, (blockIdx.x - blocks_a) * WARPS + wib, (gridDim.x - blocks_a) * WARPS, lane);
}

template <int FPL, int THREADS, int MINB>
static int launch_apply_q8(const ApplyTable& A, const ApplyTable& B, cudaStream_t stream) {
  auto kern = apply_adagrad_q8_kernel<FPL, THREADS, MINB>;
  constexpr int WARPS = THREADS / 32;
  static int per_sm_cached = 0;
  if (per_sm_cached == 0) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, THREADS, 0) != cudaSuccess || per_sm < 1)
      per_sm = 1;
    per_sm_cached = per_sm;
  }
  const int full = sm_count() * per_sm_cached;
  auto blocks_for = [&](const ApplyTable& T) {
    const int rows_per_warp = T.touched ? 32 : 4;
    int need = ((T.rows + rows_per_warp - 1) / rows_per_warp + WARPS - 1) / WARPS;
    return need > full ? full : need;
  };
  const int ba = A.rows > 0 ? blocks_for(A) : 0;
  const int bb = B.rows > 0 ? blocks_for(B) : 0;
  if (ba + bb < 1) return 0;
  kern<<<ba + bb, THREADS, 0, stream>>>(A, B, ba);
  MKE_CHECK_LAUNCH("apply_adagrad_q8_kernel");
  return 0;
}

// returns 1 when the stride has no quarter-warp instantiation
static int dispatch_apply_q8(int stride, const ApplyTable& A, const ApplyTable& B, cudaStream_t s) {
  switch (stride) {
    case 32: return launch_apply_q8<4, 128, 12>(A, B, s);
    case 64: return launch_apply_q8<8, 128, 10>(A, B, s);
    case 80: return launch_apply_q8<10, 128, 6>(A, B, s);
    case 104: return launch_apply_q8<13, 128, 8>(A, B, s);
    case 128: return launch_apply_q8<16, 128, 6>(A, B, s);
    default: return 1;
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
  int need = ((table_local_rows(t) + 31) / 32 + kApplyWarps - 1) / kApplyWarps;
  if (need > full) need = full;
  if (need < 1) need = 1;
  kern<<<need, kApplyThreads, 0, stream>>>(t->var, t->grad, t->touched, acc, table_local_rows(t), t->stride,
                                           (t->dim + 3) / 4, t->normalised, lr,
                                           t->grad_replicas > 1 ? t->grad_replicas : 1);
  MKE_CHECK_LAUNCH("apply_adagrad_kernel");
  return 0;
}

}  // namespace mke

using namespace mke;

extern "C" int mke_rows_apply_adagrad(const mke_table_t* table, float* acc, float lr,
                                      mke_stream_t stream) {
  MKE_CHECK_ARG(table && table->var && table->grad && acc, "apply needs var/grad/acc");
  MKE_CHECK_ARG(table->stride % 4 == 0 && table->dim <= table->stride && table->dim > 0,
                "bad stride/dim");
  if (table->rows <= 0) return 0;
  const int nv = ((table->dim + 3) / 4 + 31) / 32;
  MKE_CHECK_ARG(nv <= 8, "dim too large");
  cudaStream_t s = (cudaStream_t)stream;
  static const int generic = getenv("MKE_APPLY_GENERIC") ? atoi(getenv("MKE_APPLY_GENERIC")) : 0;
  if (!generic) {
    const ApplyTable A = apply_table(table, acc, lr);
    const ApplyTable B{nullptr, nullptr, nullptr, nullptr, 0, 0, 0.f, 1, 0};
    const int rc = dispatch_apply_q8(table->stride, A, B, s);
    if (rc <= 0) return rc;
  }
  switch (nv) {
    case 1: return launch_apply<1>(table, acc, lr, s);
    case 2: return launch_apply<2>(table, acc, lr, s);
    case 3: case 4: return launch_apply<4>(table, acc, lr, s);
    default: return launch_apply<8>(table, acc, lr, s);
  }
}

extern "C" int mke_rows_apply_adagrad_pair(const mke_table_t* a, float* acc_a, float lr_a,
                                           const mke_table_t* b, float* acc_b, float lr_b,
                                           mke_stream_t stream) {
  MKE_CHECK_ARG(a && b, "null table");
  if (a->stride == b->stride && a->var && a->grad && acc_a && b->var && b->grad && acc_b &&
      a->rows > 0 && b->rows > 0 && !(a->grad_replicas > 1 && a->n_shards > 1) &&
      (int64_t)a->rows + (int64_t)b->rows < (1ll << 31)) {
    const ApplyTable A = apply_table(a, acc_a, lr_a);
    const ApplyTable B = apply_table(b, acc_b, lr_b);
    const int rc = dispatch_apply_q8(a->stride, A, B, (cudaStream_t)stream);
    if (rc <= 0) return rc;
  }
  if (int rc = mke_rows_apply_adagrad(a, acc_a, lr_a, stream)) return rc;
  return mke_rows_apply_adagrad(b, acc_b, lr_b, stream);
}
