// Attribute view, LIVE score of the reference: conv() of MultiKE_model.py:34-63 with the loss of
// :144-149 / :183 / :214-218 and its complete backward, for one batch of (h, a, v[, w]) rows.
//
//   x0 = BN([a; v])            y = gamma * x / sqrt(1 + 1e-3) + beta   (inference-mode BN over width)
//   c1 = tanh(conv2x4(x0))     1 -> 2 channels, SAME (extra padding at the end: rows (0,1), cols (1,2))
//   c2 = tanh(conv2x4(c1))     2 -> 2 channels
//   z  = l2_normalize(c2, width axis)            per (row, channel)
//   u  = tanh(flat(z) Wd + bd)                   flat index (h*D + w)*2 + c,  Wd [4D, D]
//   u^ = u / |U|_F                               GLOBAL norm of the whole [B, D] batch tensor (:60)
//   loss += scale * w_i * log(1 + exp(|h^_i - u^_i|^2))
//
// The global norm couples the batch, so the step is sample-parallel passes separated by two scalar reductions
// (S = sum u^2, T = sum g_u^ . u^).  The dense layer (4D -> D, 300 x 75 at dim 75) is the step's only real arithmetic --
// 0.68 GFLOP at the shipped batch of 5 000 -- and it is three true GEMMs, which run on the tensor cores at
// fp32-equivalent precision (gemm_tf32x3_launch, mke_gemm.cu; as per-warp GEMVs they were 485 of the step's 498 us):
//   P1  conv forward (one warp per sample, activations in shared memory)  -> Z [n, 4D] (TF32 part + remainder)
//   G1  U_pre = Z Wd + bd                                                   (tcgen05)
//   PU  U = tanh(U_pre), S = sum U^2
//   P2  score / loss -> coef, T, entity gradient rows
//   PG  G_pre = d loss / d U_pre                                           -> GP [n, D] (split), gbd
//   G2  G_z = G_pre Wd^T                                                    (tcgen05)
//   P3  conv backward from G_z (conv forward recomputed) -> attribute gradient rows, conv / BN parameter gradients
//   G3  gWd += Z^T G_pre  (K = n, split over grid z, accumulating epilogue)  (tcgen05, on transposed copies)
// Parameters travel as ONE flat vector theta (layout: oracle/attr_cnn.py::layout):
//   gamma[D] beta[D] k1[2][4][1][2] b1[2] k2[2][4][2][2] b2[2] wd[4D][D] bd[D]
#include "mke_common.cuh"

namespace mke {

constexpr int kCnnThreads = 128;
constexpr int kCnnWarps = kCnnThreads / 32;
constexpr int kCnnMaxD = 128;
constexpr float kBnEps = 1e-3f;
constexpr int kSmall = 52;  // k1 16 + b1 2 + k2 32 + b2 2
constexpr int kCnnMaxBlocks = 1024;  // sample-parallel grids are capped at 4 blocks per SM

struct CnnLayout {
  int D;
  __host__ __device__ int gamma() const { return 0; }
  __host__ __device__ int beta() const { return D; }
  __host__ __device__ int k1() const { return 2 * D; }
  __host__ __device__ int b1() const { return 2 * D + 16; }
  __host__ __device__ int k2() const { return 2 * D + 18; }
  __host__ __device__ int b2() const { return 2 * D + 50; }
  __host__ __device__ int wd() const { return 2 * D + 52; }
  __host__ __device__ int bd() const { return 2 * D + 52 + 4 * D * D; }
  __host__ __device__ int total() const { return 2 * D + 52 + 4 * D * D + D; }
};

struct CnnParams {
  const float *ent_var, *attr_var, *val_var;
  float *ent_grad, *attr_grad;
  uint8_t *ent_touched, *attr_touched;
  int ent_stride, attr_stride, val_stride, ent_norm, attr_norm, val_norm;
  int attr_rep;
  size_t attr_rep_floats;
  const int32_t *ih, *ia, *iv;
  const float* w;
  int n, D;
  float scale;
  const float* theta;
  float* gtheta;
  float *U, *COEF;           // workspace [n, D] (pre-activation, then u), [n]
  float *Z_hi, *Z_lo;        // [n, 4D] normalised conv output, as the GEMMs read it
  float *GP_hi, *GP_lo;      // [n, Dp] gradient of the dense pre-activation (pad columns zero)
  float* GZ;                 // [n, 4D] gradient of z
  float* PART;               // [gridDim.x][kSmall + 3D] per-block partial sums of the small parameter gradients
  int Dp;
  double* red;               // [0] = S, [1] = T
  double* loss;
};

// per-warp activations
struct WarpAct {
  float* x0;  // [2][D]
  float* c1;  // [2][D][2]
  float* c2;  // [2][D][2]   (z after normalisation)
};
__device__ __forceinline__ int act_floats(int D) { return 2 * D + 4 * D + 4 * D; }

// row of a table as the model reads it (normalised view if flagged); lanes stride over columns
__device__ __forceinline__ float row_scale(const float* row, int D, int normalised, int lane) {
  if (!normalised) return 1.f;
  float s = 0.f;
  for (int c = lane; c < D; c += 32) s = fmaf(row[c], row[c], s);
  return rsqrtf(fmaxf(warp_sum(s), kNormEps));
}

// forward of one sample up to z (in act.c2) and the four inverse norms; small = smem copy of
// gamma..b2.  Returns nothing; all lanes must call.
__device__ __forceinline__ void cnn_forward(const CnnParams& p, const float* small, const WarpAct& act,
                                            const float* arow, float ascale, const float* vrow, float vscale,
                                            int lane, float (&inv_n)[4], float (&nsum)[4]) {
  const int D = p.D;
  const float* gamma = small;
  const float* beta = small + D;
  const float* k1 = small + 2 * D;
  const float* b1 = k1 + 16;
  const float* k2 = b1 + 2;
  const float* b2 = k2 + 32;
  const float bn = rsqrtf(1.f + kBnEps);
  for (int w = lane; w < D; w += 32) {
    const float gs = gamma[w] * bn;
    act.x0[w] = fmaf(arow[w] * ascale, gs, beta[w]);
    act.x0[D + w] = fmaf(vrow[w] * vscale, gs, beta[w]);
  }
  __syncwarp();
  for (int w = lane; w < D; w += 32) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float o0 = b1[0], o1 = b1[1];
#pragma unroll
      for (int dh = 0; dh < 2; ++dh) {
        if (h + dh >= 2) continue;
#pragma unroll
        for (int dw = 0; dw < 4; ++dw) {
          const int ww = w + dw - 1;
          if (ww < 0 || ww >= D) continue;
          const float x = act.x0[(h + dh) * D + ww];
          o0 = fmaf(k1[(dh * 4 + dw) * 2 + 0], x, o0);
          o1 = fmaf(k1[(dh * 4 + dw) * 2 + 1], x, o1);
        }
      }
      act.c1[(h * D + w) * 2 + 0] = tanhf(o0);
      act.c1[(h * D + w) * 2 + 1] = tanhf(o1);
    }
  }
  __syncwarp();
  float nn[4] = {0.f, 0.f, 0.f, 0.f};
  for (int w = lane; w < D; w += 32) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float o0 = b2[0], o1 = b2[1];
#pragma unroll
      for (int dh = 0; dh < 2; ++dh) {
        if (h + dh >= 2) continue;
#pragma unroll
        for (int dw = 0; dw < 4; ++dw) {
          const int ww = w + dw - 1;
          if (ww < 0 || ww >= D) continue;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const float x = act.c1[((h + dh) * D + ww) * 2 + c];
            o0 = fmaf(k2[((dh * 4 + dw) * 2 + c) * 2 + 0], x, o0);
            o1 = fmaf(k2[((dh * 4 + dw) * 2 + c) * 2 + 1], x, o1);
          }
        }
      }
      const float t0 = tanhf(o0), t1 = tanhf(o1);
      act.c2[(h * D + w) * 2 + 0] = t0;
      act.c2[(h * D + w) * 2 + 1] = t1;
      nn[h * 2 + 0] = fmaf(t0, t0, nn[h * 2 + 0]);
      nn[h * 2 + 1] = fmaf(t1, t1, nn[h * 2 + 1]);
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    nsum[k] = warp_sum(nn[k]);
    inv_n[k] = rsqrtf(fmaxf(nsum[k], kNormEps));
  }
  __syncwarp();
  for (int w = lane; w < D; w += 32) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      act.c2[(h * D + w) * 2 + 0] *= inv_n[h * 2 + 0];
      act.c2[(h * D + w) * 2 + 1] *= inv_n[h * 2 + 1];
    }
  }
  __syncwarp();
}

__device__ __forceinline__ WarpAct warp_act(float* base, int wib, int D, int extra) {
  float* b = base + (size_t)wib * (act_floats(D) + extra);
  return WarpAct{b, b + 2 * D, b + 6 * D};
}

// x -> TF32 part + exact remainder (the operand format of gemm_tf32x3)
__device__ __forceinline__ void cnn_split(float v, float& hi, float& lo) {
  hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
  lo = v - hi;
}

// ---- P1: conv forward to z ------------------------------------------------------------------
__global__ void __launch_bounds__(kCnnThreads) cnn_p1_kernel(const CnnParams p) {
  extern __shared__ float smem[];
  const int D = p.D, lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  float* small = smem;
  for (int k = threadIdx.x; k < 2 * D + kSmall; k += kCnnThreads) small[k] = p.theta[k];
  __syncthreads();
  const WarpAct act = warp_act(smem + 2 * D + kSmall, wib, D, 0);
  for (int i = blockIdx.x * kCnnWarps + wib; i < p.n; i += gridDim.x * kCnnWarps) {
    const float* arow = p.attr_var + (size_t)__ldg(p.ia + i) * p.attr_stride;
    const float* vrow = p.val_var + (size_t)__ldg(p.iv + i) * p.val_stride;
    const float as = row_scale(arow, D, p.attr_norm, lane), vs = row_scale(vrow, D, p.val_norm, lane);
    float inv_n[4], nsum[4];
    cnn_forward(p, small, act, arow, as, vrow, vs, lane, inv_n, nsum);
    for (int k = lane; k < 4 * D; k += 32) {
      float hi, lo;
      cnn_split(act.c2[k], hi, lo);
      p.Z_hi[(size_t)i * 4 * D + k] = hi;
      p.Z_lo[(size_t)i * 4 * D + k] = lo;
    }
    __syncwarp();
  }
}

// ---- PU: u = tanh(pre-activation), S = sum u^2 ----------------------------------------------------
__global__ void cnn_u_kernel(const CnnParams p) {
  const size_t total = (size_t)p.n * p.D;
  float ssum = 0.f;
  for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < total; k += (size_t)gridDim.x * blockDim.x) {
    const float u = tanhf(p.U[k]);
    p.U[k] = u;
    ssum = fmaf(u, u, ssum);
  }
  // one atomic per block: thousands of fp64 atomics on ONE address serialise in a single L2 slice
  __shared__ float s_w[32];
  ssum = warp_sum(ssum);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = ssum;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += (double)s_w[w];
    if (tot != 0.0) atomicAdd(p.red, tot);
  }
}

// Wd [4D, D] of theta -> Wd^T [D, 4D] (operand B of G1) and Wd padded to [4D, Dp] (operand B of G2), both split
__global__ void cnn_wd_prep_kernel(const float* __restrict__ wd, int D, int Dp, float* __restrict__ t_hi,
                                   float* __restrict__ t_lo, float* __restrict__ p_hi, float* __restrict__ p_lo) {
  const int total = 4 * D * Dp;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int k = idx / Dp, j = idx - k * Dp;
    float hi = 0.f, lo = 0.f;
    if (j < D) cnn_split(wd[(size_t)k * D + j], hi, lo);
    p_hi[idx] = hi;
    p_lo[idx] = lo;
    if (j < D) {
      t_hi[(size_t)j * 4 * D + k] = hi;
      t_lo[(size_t)j * 4 * D + k] = lo;
    }
  }
}

// src [R, ld] (two arrays) -> dst [C, Rp] (two arrays), 32 x 32 tiles through shared memory
__global__ void cnn_transpose2_kernel(const float* __restrict__ a, const float* __restrict__ b, int R, int C, int ld,
                                      float* __restrict__ at, float* __restrict__ bt, int Rp) {
  __shared__ float ta[32][33], tb[32][33];
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int y = threadIdx.y; y < 32; y += blockDim.y) {
    const int r = r0 + y, c = c0 + threadIdx.x;
    const bool ok = r < R && c < C;
    ta[y][threadIdx.x] = ok ? a[(size_t)r * ld + c] : 0.f;
    tb[y][threadIdx.x] = ok ? b[(size_t)r * ld + c] : 0.f;
  }
  __syncthreads();
  for (int y = threadIdx.y; y < 32; y += blockDim.y) {
    const int c = c0 + y, r = r0 + threadIdx.x;
    if (c < C && r < Rp) {
      at[(size_t)c * Rp + r] = ta[threadIdx.x][y];
      bt[(size_t)c * Rp + r] = tb[threadIdx.x][y];
    }
  }
}

// ---- P2: score, loss, coef, T, entity gradient rows ---------------------------------------------
__global__ void __launch_bounds__(kCnnThreads) cnn_p2_kernel(const CnnParams p) {
  const int D = p.D, lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const float rS = rsqrtf(fmaxf((float)p.red[0], kNormEps));
  float loss_local = 0.f, t_local = 0.f;
  for (int i = blockIdx.x * kCnnWarps + wib; i < p.n; i += gridDim.x * kCnnWarps) {
    const int32_t h = __ldg(p.ih + i);
    const float* hrow = p.ent_var + (size_t)h * p.ent_stride;
    const float hs = row_scale(hrow, D, p.ent_norm, lane);
    float d[kCnnMaxD / 32], uh[kCnnMaxD / 32], s = 0.f;
#pragma unroll
    for (int q = 0; q < kCnnMaxD / 32; ++q) {
      const int j = lane + 32 * q;
      uh[q] = (j < D) ? p.U[(size_t)i * D + j] * rS : 0.f;
      d[q] = (j < D) ? hrow[j] * hs - uh[q] : 0.f;
      s = fmaf(d[q], d[q], s);
    }
    s = warp_sum(s);
    const float ex = expf(s), onep = 1.f + ex;                  // log(1 + exp(-score)), score = -s
    const float wgt = (p.w ? __ldg(p.w + i) : 1.f) * p.scale;
    loss_local += wgt * logf(onep);
    const float coef = 2.f * wgt * (ex / onep);
    float tt = 0.f;
#pragma unroll
    for (int q = 0; q < kCnnMaxD / 32; ++q) {
      const int j = lane + 32 * q;
      if (j < D) {
        tt = fmaf(-coef * d[q], uh[q], tt);                      // g_u^ . u^
        if (p.ent_grad) atomicAdd(p.ent_grad + (size_t)h * p.ent_stride + j, coef * d[q]);
      }
    }
    t_local += tt;
    if (lane == 0) {
      p.COEF[i] = coef;
      if (p.ent_grad) mark_touched(p.ent_touched, h);
    }
  }
  loss_local = warp_sum(loss_local);   // every lane held the same value: undo the 32x
  t_local = warp_sum(t_local);
  __shared__ float s_l[kCnnWarps], s_t[kCnnWarps];
  if (lane == 0) {
    s_l[wib] = loss_local;
    s_t[wib] = t_local;
  }
  __syncthreads();
  if (threadIdx.x == 0) {  // one atomic pair per block
    double l = 0.0, t = 0.0;
    for (int w = 0; w < kCnnWarps; ++w) {
      l += (double)s_l[w];
      t += (double)s_t[w];
    }
    if (l != 0.0 && p.loss) atomicAdd(p.loss, l / 32.0);
    if (t != 0.0) atomicAdd(p.red + 1, t);
  }
}

// ---- PG: gradient of the dense pre-activation, gbd --------------------------------------------------
__global__ void __launch_bounds__(kCnnThreads) cnn_gpre_kernel(const CnnParams p) {
  const int D = p.D, Dp = p.Dp, lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const float S = (float)p.red[0], T = (float)p.red[1];
  const float rS = rsqrtf(fmaxf(S, kNormEps));
  float gbd[kCnnMaxD / 32];
#pragma unroll
  for (int q = 0; q < kCnnMaxD / 32; ++q) gbd[q] = 0.f;
  for (int i = blockIdx.x * kCnnWarps + wib; i < p.n; i += gridDim.x * kCnnWarps) {
    const int32_t h = __ldg(p.ih + i);
    const float* hrow = p.ent_var + (size_t)h * p.ent_stride;
    const float hs = row_scale(hrow, D, p.ent_norm, lane);
    const float coef = p.COEF[i];
#pragma unroll
    for (int q = 0; q < kCnnMaxD / 32; ++q) {
      const int j = lane + 32 * q;
      if (j < Dp) {
        float gp = 0.f;
        if (j < D) {
          const float u = p.U[(size_t)i * D + j];
          const float uh = u * rS;
          const float guh = -coef * (hrow[j] * hs - uh);
          const float gu = (S >= kNormEps) ? (guh - uh * T) * rS : guh * rS;
          gp = gu * (1.f - u * u);
          gbd[q] += gp;
        }
        float hi, lo;
        cnn_split(gp, hi, lo);
        p.GP_hi[(size_t)i * Dp + j] = hi;
        p.GP_lo[(size_t)i * Dp + j] = lo;
      }
    }
  }
  // per-block partial sums (no atomics on the 75 hot addresses): PART[block][kSmall + 2D + j]
  __shared__ float s_bd[kCnnWarps][kCnnMaxD];
#pragma unroll
  for (int q = 0; q < kCnnMaxD / 32; ++q) s_bd[wib][lane + 32 * q] = gbd[q];
  __syncthreads();
  for (int j = threadIdx.x; j < D; j += kCnnThreads) {
    float v = 0.f;
    for (int w = 0; w < kCnnWarps; ++w) v += s_bd[w][j];
    p.PART[(size_t)blockIdx.x * (kSmall + 3 * D) + kSmall + 2 * D + j] = v;
  }
}

// ---- P3: backward through the width normalisation, conv2, conv1, BN (g_z comes from the GEMM) -------------
__global__ void __launch_bounds__(kCnnThreads) cnn_p3_kernel(const CnnParams p) {
  extern __shared__ float smem[];
  const int D = p.D, lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const CnnLayout L{D};
  float* small = smem;
  for (int k = threadIdx.x; k < 2 * D + kSmall; k += kCnnThreads) small[k] = p.theta[k];
  __syncthreads();
  // per warp: activations + G2 [2][D][2] + G1 [2][D][2] + gx0 [2][D]
  const int extra = 4 * D + 4 * D + 2 * D;
  const WarpAct act = warp_act(smem + 2 * D + kSmall, wib, D, extra);
  float* G2 = act.c2 + 4 * D;
  float* G1 = G2 + 4 * D;
  float* gx0 = G1 + 4 * D;
  const float* k1 = small + 2 * D;
  const float* k2 = k1 + 18;
  const float* gamma = small;
  const float bn = rsqrtf(1.f + kBnEps);
  float* attr_grad = p.attr_grad ? p.attr_grad + (size_t)(blockIdx.x % (unsigned)p.attr_rep) * p.attr_rep_floats : nullptr;
  float g_small[kSmall];  // per-lane partial sums of k1, b1, k2, b2 gradients
#pragma unroll
  for (int k = 0; k < kSmall; ++k) g_small[k] = 0.f;
  float g_gamma[kCnnMaxD / 32], g_beta[kCnnMaxD / 32];
#pragma unroll
  for (int q = 0; q < kCnnMaxD / 32; ++q) g_gamma[q] = g_beta[q] = 0.f;

  for (int i = blockIdx.x * kCnnWarps + wib; i < p.n; i += gridDim.x * kCnnWarps) {
    const int32_t a = __ldg(p.ia + i);
    const float* arow = p.attr_var + (size_t)a * p.attr_stride;
    const float* vrow = p.val_var + (size_t)__ldg(p.iv + i) * p.val_stride;
    const float as = row_scale(arow, D, p.attr_norm, lane), vs = row_scale(vrow, D, p.val_norm, lane);
    float inv_n[4], nsum[4];
    cnn_forward(p, small, act, arow, as, vrow, vs, lane, inv_n, nsum);
    // g_z (from G2 = G_pre Wd^T), then the width-normalisation backward
    float dots[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k = lane; k < 4 * D; k += 32) {
      const float gz = p.GZ[(size_t)i * 4 * D + k];
      G2[k] = gz;
      dots[((k / 2) / D) * 2 + (k & 1)] += gz * act.c2[k];
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) dots[k] = warp_sum(dots[k]);
    __syncwarp();
    for (int k = lane; k < 4 * D; k += 32) {
      const int hc = ((k / 2) / D) * 2 + (k & 1);
      const float z = act.c2[k];
      const float n = 1.f / inv_n[hc];
      // y = x * rsqrt(max(sum x^2, eps)); no projection term below the clamp
      const float gc2 = (nsum[hc] >= kNormEps) ? (G2[k] - z * dots[hc]) * inv_n[hc] : G2[k] * inv_n[hc];
      const float c2 = z * n;
      G2[k] = gc2 * (1.f - c2 * c2);
    }
    __syncwarp();
    // conv2 backward: parameter gradients and g_c1 -> G1 (pre-activation gradient of conv1)
    for (int w = lane; w < D; w += 32) {
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const float go0 = G2[(hh * D + w) * 2 + 0], go1 = G2[(hh * D + w) * 2 + 1];
        g_small[50] += go0;
        g_small[51] += go1;
#pragma unroll
        for (int dh = 0; dh < 2; ++dh) {
          if (hh + dh >= 2) continue;
#pragma unroll
          for (int dw = 0; dw < 4; ++dw) {
            const int ww = w + dw - 1;
            if (ww < 0 || ww >= D) continue;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              const float x = act.c1[((hh + dh) * D + ww) * 2 + c];
              g_small[18 + ((dh * 4 + dw) * 2 + c) * 2 + 0] = fmaf(go0, x, g_small[18 + ((dh * 4 + dw) * 2 + c) * 2 + 0]);
              g_small[18 + ((dh * 4 + dw) * 2 + c) * 2 + 1] = fmaf(go1, x, g_small[18 + ((dh * 4 + dw) * 2 + c) * 2 + 1]);
            }
          }
        }
      }
    }
    for (int w = lane; w < D; w += 32) {
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          float g = 0.f;  // g_c1[hh][w][c] = sum K2[dh][dw][c][f] G2[hh-dh][w-dw+1][f]
#pragma unroll
          for (int dh = 0; dh < 2; ++dh) {
            if (hh - dh < 0) continue;
#pragma unroll
            for (int dw = 0; dw < 4; ++dw) {
              const int ww = w - dw + 1;
              if (ww < 0 || ww >= D) continue;
              g = fmaf(k2[((dh * 4 + dw) * 2 + c) * 2 + 0], G2[((hh - dh) * D + ww) * 2 + 0], g);
              g = fmaf(k2[((dh * 4 + dw) * 2 + c) * 2 + 1], G2[((hh - dh) * D + ww) * 2 + 1], g);
            }
          }
          const float c1v = act.c1[(hh * D + w) * 2 + c];
          G1[(hh * D + w) * 2 + c] = g * (1.f - c1v * c1v);
        }
      }
    }
    __syncwarp();
    // conv1 backward
    for (int w = lane; w < D; w += 32) {
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const float go0 = G1[(hh * D + w) * 2 + 0], go1 = G1[(hh * D + w) * 2 + 1];
        g_small[16] += go0;
        g_small[17] += go1;
#pragma unroll
        for (int dh = 0; dh < 2; ++dh) {
          if (hh + dh >= 2) continue;
#pragma unroll
          for (int dw = 0; dw < 4; ++dw) {
            const int ww = w + dw - 1;
            if (ww < 0 || ww >= D) continue;
            const float x = act.x0[(hh + dh) * D + ww];
            g_small[(dh * 4 + dw) * 2 + 0] = fmaf(go0, x, g_small[(dh * 4 + dw) * 2 + 0]);
            g_small[(dh * 4 + dw) * 2 + 1] = fmaf(go1, x, g_small[(dh * 4 + dw) * 2 + 1]);
          }
        }
      }
    }
    for (int w = lane; w < D; w += 32) {
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        float g = 0.f;
#pragma unroll
        for (int dh = 0; dh < 2; ++dh) {
          if (hh - dh < 0) continue;
#pragma unroll
          for (int dw = 0; dw < 4; ++dw) {
            const int ww = w - dw + 1;
            if (ww < 0 || ww >= D) continue;
            g = fmaf(k1[(dh * 4 + dw) * 2 + 0], G1[((hh - dh) * D + ww) * 2 + 0], g);
            g = fmaf(k1[(dh * 4 + dw) * 2 + 1], G1[((hh - dh) * D + ww) * 2 + 1], g);
          }
        }
        gx0[hh * D + w] = g;
      }
    }
    __syncwarp();
    // BN backward + attribute row gradient
#pragma unroll
    for (int q = 0; q < kCnnMaxD / 32; ++q) {
      const int w = lane + 32 * q;
      if (w < D) {
        const float ga = gx0[w], gv = gx0[D + w];
        g_gamma[q] = fmaf(ga * bn, arow[w] * as, fmaf(gv * bn, vrow[w] * vs, g_gamma[q]));
        g_beta[q] += ga + gv;
        if (attr_grad) atomicAdd(attr_grad + (size_t)a * p.attr_stride + w, ga * gamma[w] * bn);
      }
    }
    if (lane == 0 && attr_grad) mark_touched(p.attr_touched, a);
    __syncwarp();
  }
  // parameter gradients: warp sums -> block sums in shared memory -> PART[block][0 .. kSmall + 2D) (atomics on these
  // 200 addresses from 2 368 warps serialised in two L2 lines: 118 of the kernel's 133 us)
  __syncthreads();   // every warp is done with its activations: the region is reused
  float* s_part = smem + 2 * D + kSmall;   // [kCnnWarps][kSmall + 2D]
  const int PW = kSmall + 2 * D;
#pragma unroll
  for (int k = 0; k < kSmall; ++k) {
    const float v = warp_sum(g_small[k]);
    if (lane == 0) s_part[wib * PW + k] = v;
  }
#pragma unroll
  for (int q = 0; q < kCnnMaxD / 32; ++q) {
    const int w = lane + 32 * q;
    if (w < D) {
      s_part[wib * PW + kSmall + w] = g_gamma[q];
      s_part[wib * PW + kSmall + D + w] = g_beta[q];
    }
  }
  __syncthreads();
  for (int k = threadIdx.x; k < PW; k += kCnnThreads) {
    float v = 0.f;
    for (int w = 0; w < kCnnWarps; ++w) v += s_part[w * PW + k];
    p.PART[(size_t)blockIdx.x * (kSmall + 3 * D) + k] = v;
  }
}

// sum of the per-block partials -> gtheta (k1, b1, k2, b2 | gamma | beta | bd): a block of 32 warps owns 32 columns, warp w
// sums the rows w, w + 32, ... (coalesced over the columns, 19 loads in flight per thread), then shared memory
__global__ void __launch_bounds__(1024) cnn_reduce_kernel(const float* __restrict__ part, int blocks, int D,
                                                          float* __restrict__ gtheta) {
  __shared__ float s_sum[32][33];
  const CnnLayout L{D};
  const int PW = kSmall + 3 * D;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int k = blockIdx.x * 32 + lane;
  float v = 0.f;
  if (k < PW)
    for (int b = warp; b < blocks; b += 32) v += part[(size_t)b * PW + k];
  s_sum[warp][lane] = v;
  __syncthreads();
  if (warp == 0 && k < PW) {
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < 32; ++w) tot += s_sum[w][lane];
    int dst;
    if (k < kSmall) dst = L.k1() + k;
    else if (k < kSmall + D) dst = L.gamma() + (k - kSmall);
    else if (k < kSmall + 2 * D) dst = L.beta() + (k - kSmall - D);
    else dst = L.bd() + (k - kSmall - 2 * D);
    gtheta[dst] += tot;
  }
}

// dense Adagrad for the flat parameter vector (AdagradOptimizer on tf.layers variables)
__global__ void dense_adagrad_kernel(float* __restrict__ theta, float* __restrict__ g, float* __restrict__ acc,
                                     int n, float lr) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    const float gv = g[k];
    const float a = fmaf(gv, gv, acc[k]);
    acc[k] = a;
    theta[k] -= lr * gv * (a > 0.f ? rsqrtf(a) : 0.f);
    g[k] = 0.f;
  }
}

}  // namespace mke

using namespace mke;

extern "C" int64_t mke_attr_cnn_param_count(int32_t dim) {
  return dim > 0 ? (int64_t)CnnLayout{dim}.total() : 0;
}
namespace mke {
int gemm_tf32x3_launch(const float* a_hi, const float* a_lo, int64_t lda, const float* b_hi, const float* b_lo, int64_t ldb,
                       int M, int N, int K, const float* bias_or_null, float* C, int64_t ldc, int k_splits, int accumulate,
                       cudaStream_t stream);

// workspace layout (floats; every array starts on a 16-byte boundary)
struct CnnWorkspace {
  size_t U, COEF, Z_hi, Z_lo, GP_hi, GP_lo, GZ, ZT_hi, ZT_lo, GPT_hi, GPT_lo, WT_hi, WT_lo, WP_hi, WP_lo, PART, red, total;
  int Dp, np;
  CnnWorkspace(int n, int D) {
    auto up = [](size_t x) { return (x + 3) & ~(size_t)3; };
    Dp = (D + 3) & ~3;
    np = (n + 3) & ~3;
    size_t off = 0;
    auto take = [&](size_t count) {
      const size_t at = off;
      off = up(off + count);
      return at;
    };
    U = take((size_t)n * D);
    COEF = take(n);
    Z_hi = take((size_t)n * 4 * D);
    Z_lo = take((size_t)n * 4 * D);
    GP_hi = take((size_t)n * Dp);
    GP_lo = take((size_t)n * Dp);
    GZ = take((size_t)n * 4 * D);
    ZT_hi = take((size_t)4 * D * np);
    ZT_lo = take((size_t)4 * D * np);
    GPT_hi = take((size_t)Dp * np);
    GPT_lo = take((size_t)Dp * np);
    WT_hi = take((size_t)D * 4 * D);
    WT_lo = take((size_t)D * 4 * D);
    WP_hi = take((size_t)4 * D * Dp);
    WP_lo = take((size_t)4 * D * Dp);
    PART = take((size_t)kCnnMaxBlocks * (kSmall + 3 * D));
    red = take(4);  // two fp64 scalars
    total = off + 8;
  }
};
}  // namespace mke

extern "C" int64_t mke_attr_cnn_workspace_floats(int32_t n, int32_t dim) {
  if (n < 0 || dim <= 0) return -1;
  return (int64_t)CnnWorkspace(n, dim).total;
}

extern "C" int mke_attr_cnn_fwd_bwd(const mke_table_t* ent, const mke_table_t* attr, const mke_table_t* val,
                                    const int32_t* ih, const int32_t* ia, const int32_t* iv, int32_t n,
                                    const float* w_or_null, float scale, const float* theta, float* gtheta,
                                    float* workspace, double* loss_accum, mke_stream_t stream) {
  MKE_CHECK_ARG(ent && attr && val && ent->var && attr->var && val->var, "null table");
  const int D = ent->dim;
  MKE_CHECK_ARG(D > 0 && D <= kCnnMaxD && attr->dim == D && val->dim == D, "tables disagree on dim or dim > %d", kCnnMaxD);
  MKE_CHECK_ARG(ent->n_shards <= 1 && attr->n_shards <= 1 && ent->grad_replicas <= 1, "plain entity table expected");
  MKE_CHECK_ARG(n >= 0, "negative n");
  if (n == 0) return 0;
  MKE_CHECK_ARG(ih && ia && iv && theta && gtheta && workspace, "null pointer");
  MKE_CHECK_ARG(((uintptr_t)workspace & 15) == 0, "workspace must be 16-byte aligned");
  const CnnWorkspace W(n, D);
  const CnnLayout L{D};
  CnnParams p{};
  p.ent_var = ent->var; p.attr_var = attr->var; p.val_var = val->var;
  p.ent_grad = ent->grad; p.attr_grad = attr->grad;
  p.ent_touched = ent->grad ? ent->touched : nullptr;
  p.attr_touched = attr->grad ? attr->touched : nullptr;
  p.ent_stride = ent->stride; p.attr_stride = attr->stride; p.val_stride = val->stride;
  p.ent_norm = ent->normalised; p.attr_norm = attr->normalised; p.val_norm = val->normalised;
  p.attr_rep = attr->grad_replicas > 1 ? attr->grad_replicas : 1;
  p.attr_rep_floats = (size_t)attr->rows * attr->stride;
  p.ih = ih; p.ia = ia; p.iv = iv; p.w = w_or_null;
  p.n = n; p.D = D; p.Dp = W.Dp; p.scale = scale; p.theta = theta; p.gtheta = gtheta;
  p.U = workspace + W.U;
  p.COEF = workspace + W.COEF;
  p.Z_hi = workspace + W.Z_hi; p.Z_lo = workspace + W.Z_lo;
  p.GP_hi = workspace + W.GP_hi; p.GP_lo = workspace + W.GP_lo;
  p.GZ = workspace + W.GZ;
  p.PART = workspace + W.PART;
  float *zt_hi = workspace + W.ZT_hi, *zt_lo = workspace + W.ZT_lo, *gpt_hi = workspace + W.GPT_hi, *gpt_lo = workspace + W.GPT_lo;
  float *wt_hi = workspace + W.WT_hi, *wt_lo = workspace + W.WT_lo, *wp_hi = workspace + W.WP_hi, *wp_lo = workspace + W.WP_lo;
  p.red = reinterpret_cast<double*>(workspace + W.red);
  p.loss = loss_accum;
  cudaStream_t s = (cudaStream_t)stream;
  if (cudaError_t e = cudaMemsetAsync(p.red, 0, 2 * sizeof(double), s)) return cuda_fail(e, "cudaMemsetAsync");
  int blocks = (n + kCnnWarps - 1) / kCnnWarps;
  int full = sm_count() * 4;
  if (full > kCnnMaxBlocks) full = kCnnMaxBlocks;
  if (blocks > full) blocks = full;
  const size_t smem1 = (size_t)(2 * D + kSmall + kCnnWarps * (10 * D)) * sizeof(float);
  const size_t smem3 = (size_t)(2 * D + kSmall + kCnnWarps * (10 * D + 10 * D)) * sizeof(float);
  if (smem3 > 48 * 1024) {
    if (cudaError_t e = cudaFuncSetAttribute(cnn_p3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3))
      return cuda_fail(e, "cudaFuncSetAttribute");
  }
  cnn_wd_prep_kernel<<<(4 * D * W.Dp + 255) / 256, 256, 0, s>>>(theta + L.wd(), D, W.Dp, wt_hi, wt_lo, wp_hi, wp_lo);
  MKE_CHECK_LAUNCH("cnn_wd_prep_kernel");
  cnn_p1_kernel<<<blocks, kCnnThreads, smem1, s>>>(p);
  MKE_CHECK_LAUNCH("cnn_p1_kernel");
  // G1: U_pre [n, D] = Z [n, 4D] . (Wd^T [D, 4D])^T + bd
  if (int rc = gemm_tf32x3_launch(p.Z_hi, p.Z_lo, 4 * D, wt_hi, wt_lo, 4 * D, n, D, 4 * D, theta + L.bd(), p.U, D, 1, 0, s)) return rc;
  int eb = (int)(((size_t)n * D + 255) / 256);
  if (eb > sm_count() * 2) eb = sm_count() * 2;
  cnn_u_kernel<<<eb, 256, 0, s>>>(p);
  MKE_CHECK_LAUNCH("cnn_u_kernel");
  cnn_p2_kernel<<<blocks, kCnnThreads, 0, s>>>(p);
  MKE_CHECK_LAUNCH("cnn_p2_kernel");
  cnn_gpre_kernel<<<blocks, kCnnThreads, 0, s>>>(p);
  MKE_CHECK_LAUNCH("cnn_gpre_kernel");
  // G2: G_z [n, 4D] = G_pre [n, D] . (Wd [4D, D])^T
  if (int rc = gemm_tf32x3_launch(p.GP_hi, p.GP_lo, W.Dp, wp_hi, wp_lo, W.Dp, n, 4 * D, D, nullptr, p.GZ, 4 * D, 1, 0, s)) return rc;
  cnn_p3_kernel<<<blocks, kCnnThreads, smem3, s>>>(p);
  MKE_CHECK_LAUNCH("cnn_p3_kernel");
  // G3: gWd [4D, D] += (Z^T [4D, n]) . (G_pre^T [D, n])^T, K = n split over the grid
  cnn_transpose2_kernel<<<dim3((4 * D + 31) / 32, (W.np + 31) / 32), dim3(32, 8), 0, s>>>(p.Z_hi, p.Z_lo, n, 4 * D, 4 * D, zt_hi, zt_lo, W.np);
  MKE_CHECK_LAUNCH("cnn_transpose2_kernel");
  cnn_transpose2_kernel<<<dim3((D + 31) / 32, (W.np + 31) / 32), dim3(32, 8), 0, s>>>(p.GP_hi, p.GP_lo, n, D, W.Dp, gpt_hi, gpt_lo, W.np);
  MKE_CHECK_LAUNCH("cnn_transpose2_kernel");
  cnn_reduce_kernel<<<(kSmall + 3 * D + 31) / 32, 1024, 0, s>>>(p.PART, blocks, D, gtheta);
  MKE_CHECK_LAUNCH("cnn_reduce_kernel");
  const int k_splits = sm_count() / (((4 * D + 127) / 128) * ((D + 127) / 128));
  if (int rc = gemm_tf32x3_launch(zt_hi, zt_lo, W.np, gpt_hi, gpt_lo, W.np, 4 * D, D, n, nullptr, gtheta + L.wd(), D,
                                  k_splits < 1 ? 1 : k_splits, 1, s))
    return rc;
  return 0;
}

extern "C" int mke_dense_apply_adagrad(float* theta, float* grad, float* acc, int64_t n, float lr,
                                       mke_stream_t stream) {
  MKE_CHECK_ARG(theta && grad && acc && n >= 0 && n < (1ll << 31), "bad dense Adagrad arguments");
  if (n == 0) return 0;
  int blocks = (int)((n + 255) / 256);
  const int full = sm_count() * 8;
  if (blocks > full) blocks = full;
  dense_adagrad_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(theta, grad, acc, (int)n, lr);
  MKE_CHECK_LAUNCH("dense_adagrad_kernel");
  return 0;
}
