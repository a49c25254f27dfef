// SSL space-mapping step (MultiKE_model.py:241-261, trainer :439-454; losses.py:53-63): for the three views
// X_k in {name, rv, av} of one batch of entity ids, with F = ent_embeds (shared) and M_k the 3 d x d mappings,
//   loss = sum_k  |F - gl2n(X_k M_k)|^2 + ow |M_k M_k^T - I|^2 + norm_w |M_k|^2
// (gl2n = tf.nn.l2_normalize without axis: ONE norm over the whole batch), and its backward:
// gradient rows of the shared table (only `shared*` variables train, MultiKE_model.py:257) and the three mapping
// gradients.  The reference builds this from 3 GEMMs, 3 global norms and autograd; here it is three launches,
// because the global norm makes every gradient depend on three batch-wide scalars.  With Y = X M, S = sum Y^2,
// c = sum F.Y, f2 = sum F^2, r = rsqrt(max(S, eps)), G1 = X^T Y, G2 = X^T F (all accumulated in ONE pass over the rows):
//   |F - r Y|^2            = f2 - 2 r c + r^2 S
//   dL/dY                   = -2 r (F - r Y) - r^3 t Y,   t = sum dZ.Y = -2 (c - r S)        (no second term below eps)
//   dL/dM (mapping term)    = X^T dL/dY = -2 r (G2 - r G1) - r^3 t G1
//   dL/dF (per row)         = 2 sum_k (F - r_k Y_k)                                          (second pass over the rows)
//   d/dM ow |M M^T - I|^2   = 4 ow (M M^T - I) M,      d/dM norm_w |M|^2 = 2 norm_w M
// pass A: rows -> Y (kept), S, c, f2, G1, G2     finalize: one block, the d x d algebra     pass B: rows -> dL/dF.
#include "mke_common.cuh"

namespace mke {

constexpr int kSmThreads = 256;
constexpr int kSmRows = 32;   // rows per tile
constexpr int kSmMaxDim = 128;

struct SmTable {
  const float* var;
  int stride, normalised;
};
struct SmParams {
  SmTable f, x[3];
  float* f_grad;
  uint8_t* f_touched;
  const int32_t* idx;
  int n, dim, ld;          // ld = dim rounded up to 4 (tile pitch and pitch of Y in the workspace)
  const float* maps;       // [3][dim][dim]
  float* maps_grad;        // [3][dim][dim] (overwritten)
  float* Y;                // [3][n][ld]
  float* G;                // [3][2][dim][dim]  (G1, G2), zeroed by the caller
  double* sc;              // [0..2] S_k, [3..5] c_k, [6] f2, [8..10] r_k (written by finalize), zeroed by the caller
  float ow, norm_w;
  double* loss;
};

// one normalised row (warp-cooperative) into shared memory; rows past n are zero
__device__ __forceinline__ void load_row_norm(const SmTable& t, int32_t e, bool valid, int dim, float* dst, int lane) {
  float ss = 0.f;
  const float* src = t.var + (size_t)(valid ? e : 0) * t.stride;
  for (int c = lane; c < dim; c += 32) {
    const float v = valid ? __ldg(src + c) : 0.f;
    dst[c] = v;
    ss = fmaf(v, v, ss);
  }
  ss = warp_sum(ss);
  const float k = t.normalised ? rsqrtf(fmaxf(ss, kNormEps)) : 1.f;
  for (int c = lane; c < dim; c += 32) dst[c] *= k;
}

__global__ void __launch_bounds__(kSmThreads) space_pass_a_kernel(const SmParams p) {
  extern __shared__ float sm[];
  const int d = p.dim, ld = p.ld;
  float* sM = sm;                    // [d][ld]
  float* sX = sM + d * ld;           // [kSmRows][ld]
  float* sF = sX + kSmRows * ld;
  float* sY = sF + kSmRows * ld;
  __shared__ double s_red[3];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tiles = (p.n + kSmRows - 1) / kSmRows;
  constexpr int kMaxOwn = (kSmMaxDim * kSmMaxDim + kSmThreads - 1) / kSmThreads;  // entries of a d x d matrix per thread
  const int own = (d * d + kSmThreads - 1) / kSmThreads;
  {
    const int k = blockIdx.y;  // one view per grid row: the three views run side by side
    for (int e = tid; e < d * d; e += kSmThreads) sM[(e / d) * ld + (e % d)] = __ldg(p.maps + (size_t)k * d * d + e);
    float g1[kMaxOwn], g2[kMaxOwn];
#pragma unroll
    for (int m = 0; m < kMaxOwn; ++m) g1[m] = g2[m] = 0.f;
    float S = 0.f, c = 0.f, f2 = 0.f;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
      __syncthreads();  // previous tile's readers are done (and sM is in place)
      for (int r = warp; r < kSmRows; r += kSmThreads / 32) {
        const int i = tile * kSmRows + r;
        const bool valid = i < p.n;
        const int32_t e = valid ? __ldg(p.idx + i) : 0;
        load_row_norm(p.x[k], e, valid, d, sX + r * ld, lane);
        load_row_norm(p.f, e, valid, d, sF + r * ld, lane);
      }
      __syncthreads();
      {  // Y = X M: 8 threads per row, columns j = (tid & 7) + 8 m
        const int r = tid >> 3, i = tile * kSmRows + r;
        const float* xr = sX + r * ld;
        for (int j = tid & 7; j < d; j += 8) {
          float acc = 0.f;
          for (int q = 0; q < d; ++q) acc = fmaf(xr[q], sM[q * ld + j], acc);
          sY[r * ld + j] = acc;
          if (i < p.n) p.Y[((size_t)k * p.n + i) * ld + j] = acc;
          S = fmaf(acc, acc, S);
          c = fmaf(acc, sF[r * ld + j], c);
          if (k == 0) f2 = fmaf(sF[r * ld + j], sF[r * ld + j], f2);
        }
      }
      __syncthreads();
      // G1 += X^T Y, G2 += X^T F: thread owns entries e = tid + 256 m of the d x d matrices
#pragma unroll
      for (int m = 0; m < kMaxOwn; ++m) {
        const int e = tid + kSmThreads * m;
        if (m < own && e < d * d) {
          const int a = e / d, b = e % d;
          float u = g1[m], v = g2[m];
          for (int r = 0; r < kSmRows; ++r) {
            const float x = sX[r * ld + a];
            u = fmaf(x, sY[r * ld + b], u);
            v = fmaf(x, sF[r * ld + b], v);
          }
          g1[m] = u;
          g2[m] = v;
        }
      }
    }
#pragma unroll
    for (int m = 0; m < kMaxOwn; ++m) {
      const int e = tid + kSmThreads * m;
      if (m < own && e < d * d) {
        atomicAdd(p.G + ((size_t)k * 2 + 0) * d * d + e, g1[m]);
        atomicAdd(p.G + ((size_t)k * 2 + 1) * d * d + e, g2[m]);
      }
    }
    // block sums of S, c, f2 -> fp64 atomics
    S = warp_sum(S);
    c = warp_sum(c);
    f2 = warp_sum(f2);
    __syncthreads();
    if (tid < 3) s_red[tid] = 0.0;
    __syncthreads();
    if (lane == 0) {
      atomicAdd(&s_red[0], (double)S);
      atomicAdd(&s_red[1], (double)c);
      atomicAdd(&s_red[2], (double)f2);
    }
    __syncthreads();
    if (tid == 0) {
      atomicAdd(p.sc + k, s_red[0]);
      atomicAdd(p.sc + 3 + k, s_red[1]);
      if (k == 0) atomicAdd(p.sc + 6, s_red[2]);
    }
  }
}

// one block per view: the scalars, the mapping gradient and the view's share of the loss
constexpr int kSmFinThreads = 1024;
__global__ void __launch_bounds__(kSmFinThreads) space_finalize_kernel(const SmParams p) {
  extern __shared__ float sm[];
  const int d = p.dim, ld = p.ld;
  float* sM = sm;            // [d][ld]
  float* sP = sM + d * ld;   // [d][ld]: M M^T - I
  __shared__ double s_loss;
  const int tid = threadIdx.x;
  if (tid == 0) s_loss = 0.0;
  const double f2 = p.sc[6];
  {
    const int k = blockIdx.x;  // one block per view
    __syncthreads();
    const float* M = p.maps + (size_t)k * d * d;
    for (int e = tid; e < d * d; e += kSmFinThreads) sM[(e / d) * ld + (e % d)] = M[e];
    __syncthreads();
    float orth = 0.f, nrm = 0.f;
    for (int e = tid; e < d * d; e += kSmFinThreads) {
      const int a = e / d, b = e % d;
      float acc = 0.f;
      for (int q = 0; q < d; ++q) acc = fmaf(sM[a * ld + q], sM[b * ld + q], acc);
      acc -= (a == b) ? 1.f : 0.f;
      sP[a * ld + b] = acc;
      orth = fmaf(acc, acc, orth);
      nrm = fmaf(sM[a * ld + b], sM[a * ld + b], nrm);
    }
    __syncthreads();
    const double S = p.sc[k], c = p.sc[3 + k];
    const bool below = S < (double)kNormEps;
    const double r = 1.0 / sqrt(S > (double)kNormEps ? S : (double)kNormEps);
    const double t = -2.0 * (c - r * S);
    const float rf = (float)r, r3t = below ? 0.f : (float)(r * r * r * t);
    const float* G1 = p.G + ((size_t)k * 2 + 0) * d * d;
    const float* G2 = p.G + ((size_t)k * 2 + 1) * d * d;
    for (int e = tid; e < d * d; e += kSmFinThreads) {
      const int a = e / d, b = e % d;
      float pm = 0.f;  // ((M M^T - I) M)[a][b]
      for (int q = 0; q < d; ++q) pm = fmaf(sP[a * ld + q], sM[q * ld + b], pm);
      const float g = -2.f * rf * (G2[e] - rf * G1[e]) - r3t * G1[e] + 4.f * p.ow * pm + 2.f * p.norm_w * sM[a * ld + b];
      p.maps_grad[(size_t)k * d * d + e] = g;
    }
    orth = warp_sum(orth);
    nrm = warp_sum(nrm);
    if ((tid & 31) == 0) atomicAdd(&s_loss, (double)p.ow * (double)orth + (double)p.norm_w * (double)nrm);
    if (tid == 0) {
      atomicAdd(&s_loss, f2 - 2.0 * r * c + r * r * S);
      p.sc[8 + k] = r;
    }
  }
  __syncthreads();
  if (tid == 0 && p.loss != nullptr) atomicAdd(p.loss, s_loss);
}

// dL/dF rows: 2 sum_k (F - r_k Y_k), one warp per row
__global__ void __launch_bounds__(kSmThreads) space_pass_b_kernel(const SmParams p) {
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const float r0 = (float)p.sc[8], r1 = (float)p.sc[9], r2 = (float)p.sc[10];
  for (int i = blockIdx.x * (kSmThreads / 32) + wib; i < p.n; i += gridDim.x * (kSmThreads / 32)) {
    const int32_t e = __ldg(p.idx + i);
    const float* src = p.f.var + (size_t)e * p.f.stride;
    float ss = 0.f;
    for (int c = lane; c < p.dim; c += 32) {
      const float v = __ldg(src + c);
      ss = fmaf(v, v, ss);
    }
    ss = warp_sum(ss);
    const float kf = p.f.normalised ? rsqrtf(fmaxf(ss, kNormEps)) : 1.f;
    const float* y0 = p.Y + ((size_t)0 * p.n + i) * p.ld;
    const float* y1 = p.Y + ((size_t)1 * p.n + i) * p.ld;
    const float* y2 = p.Y + ((size_t)2 * p.n + i) * p.ld;
    float* g = p.f_grad + (size_t)e * p.f.stride;
    for (int c = lane; c < p.dim; c += 32) {
      const float F = __ldg(src + c) * kf;
      const float gf = 2.f * ((F - r0 * y0[c]) + (F - r1 * y1[c]) + (F - r2 * y2[c]));
      atomicAdd(g + c, gf);
    }
    if (lane == 0) mark_touched(p.f_touched, e);
  }
}

}  // namespace mke

using namespace mke;

extern "C" int64_t mke_space_mapping_workspace_floats(int32_t n, int32_t dim) {
  if (n < 0 || dim < 1 || dim > kSmMaxDim) return 0;
  const int64_t ld = (dim + 3) / 4 * 4;
  return 3 * (int64_t)n * ld + 6 * (int64_t)dim * dim + 2 * 16;  // Y, G1/G2, 16 doubles
}

extern "C" int mke_space_mapping_fwd_bwd(const mke_table_t* shared, const mke_table_t* name, const mke_table_t* rv,
                                         const mke_table_t* av, const int32_t* idx, int32_t n, const float* maps,
                                         float* maps_grad, float orthogonal_weight, float norm_w, float* workspace,
                                         double* loss_accum, mke_stream_t stream) {
  MKE_CHECK_ARG(shared && name && rv && av && maps && maps_grad && workspace, "null argument");
  MKE_CHECK_ARG(shared->var && shared->grad && shared->grad_replicas <= 1 && shared->n_shards <= 1, "shared table");
  const int d = shared->dim;
  MKE_CHECK_ARG(d >= 1 && d <= kSmMaxDim, "dim %d outside [1, %d]", d, kSmMaxDim);
  MKE_CHECK_ARG(n >= 0, "negative n");
  if (n == 0) return 0;
  MKE_CHECK_ARG(idx, "idx is null");
  SmParams p{};
  const mke_table_t* views[3] = {name, rv, av};
  for (int k = 0; k < 3; ++k) {
    MKE_CHECK_ARG(views[k]->var && views[k]->dim == d && views[k]->n_shards <= 1, "view table %d", k);
    p.x[k] = SmTable{views[k]->var, views[k]->stride, views[k]->normalised};
  }
  p.f = SmTable{shared->var, shared->stride, shared->normalised};
  p.f_grad = shared->grad;
  p.f_touched = shared->touched;
  p.idx = idx;
  p.n = n;
  p.dim = d;
  p.ld = (d + 3) / 4 * 4;
  p.maps = maps;
  p.maps_grad = maps_grad;
  p.Y = workspace;
  p.G = workspace + 3 * (size_t)n * p.ld;
  p.sc = reinterpret_cast<double*>(p.G + 6 * (size_t)d * d);
  MKE_CHECK_ARG(((uintptr_t)p.sc & 7) == 0, "workspace must be 8-byte aligned");
  p.ow = orthogonal_weight;
  p.norm_w = norm_w;
  p.loss = loss_accum;
  cudaStream_t s = (cudaStream_t)stream;
  if (cudaError_t e = cudaMemsetAsync(p.G, 0, (6 * (size_t)d * d + 32) * sizeof(float), s)) return cuda_fail(e, "memset");
  const size_t smem_a = ((size_t)d * p.ld + 3 * kSmRows * p.ld) * sizeof(float);
  const size_t smem_f = 2 * (size_t)d * p.ld * sizeof(float);
  static bool configured = false;
  if (!configured) {
    const int cap = (int)((2 * (size_t)kSmMaxDim * kSmMaxDim) * sizeof(float));
    if (cudaError_t e = cudaFuncSetAttribute(space_pass_a_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, cap))
      return cuda_fail(e, "cudaFuncSetAttribute(space_pass_a_kernel)");
    if (cudaError_t e = cudaFuncSetAttribute(space_finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, cap))
      return cuda_fail(e, "cudaFuncSetAttribute(space_finalize_kernel)");
    configured = true;
  }
  const int tiles = (n + kSmRows - 1) / kSmRows;
  const int blocks_a = tiles < sm_count() ? tiles : sm_count();
  space_pass_a_kernel<<<dim3(blocks_a, 3), kSmThreads, smem_a, s>>>(p);
  MKE_CHECK_LAUNCH("space_pass_a_kernel");
  space_finalize_kernel<<<3, kSmFinThreads, smem_f, s>>>(p);
  MKE_CHECK_LAUNCH("space_finalize_kernel");
  int blocks_b = (n + kSmThreads / 32 - 1) / (kSmThreads / 32);
  if (blocks_b > sm_count() * 8) blocks_b = sm_count() * 8;
  space_pass_b_kernel<<<blocks_b, kSmThreads, 0, s>>>(p);
  MKE_CHECK_LAUNCH("space_pass_b_kernel");
  return 0;
}
