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
// The global norm couples the batch, so the step is three sample-parallel passes separated by two
// scalar reductions (S = sum u^2, T = sum g_u^ . u^), plus a small contraction for the dense
// weight gradient:  P1 forward -> U, S;  P2 score/loss -> coef, T, entity gradient rows;
// P3 backward (forward recomputed) -> attribute gradient rows, conv/BN parameter gradients, Z and
// G_pre;  P4 gWd = Z^T G_pre, gbd.  One warp per sample; activations live in shared memory.
// Parameters travel as ONE flat vector theta (layout: oracle/attr_cnn.py::layout):
//   gamma[D] beta[D] k1[2][4][1][2] b1[2] k2[2][4][2][2] b2[2] wd[4D][D] bd[D]
#include "mke_common.cuh"

namespace mke {

constexpr int kCnnThreads = 128;
constexpr int kCnnWarps = kCnnThreads / 32;
constexpr int kCnnMaxD = 128;
constexpr float kBnEps = 1e-3f;
constexpr int kSmall = 52;  // k1 16 + b1 2 + k2 32 + b2 2

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
  float *U, *COEF, *Z, *GP;  // workspace [n,D] [n] [n,4D] [n,D]
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

// u[j] = tanh(bd[j] + sum_i z[i] Wd[i][j]) for the lane's columns j = lane + 32 q
__device__ __forceinline__ void cnn_dense(const CnnParams& p, const float* z, int lane, float (&u)[kCnnMaxD / 32]) {
  const int D = p.D;
  const CnnLayout L{D};
  const float* wd = p.theta + L.wd();
  const float* bd = p.theta + L.bd();
#pragma unroll
  for (int q = 0; q < kCnnMaxD / 32; ++q) u[q] = (lane + 32 * q < D) ? __ldg(bd + lane + 32 * q) : 0.f;
  for (int i = 0; i < 4 * D; ++i) {
    const float zi = z[i];
#pragma unroll
    for (int q = 0; q < kCnnMaxD / 32; ++q)
      if (lane + 32 * q < D) u[q] = fmaf(zi, __ldg(wd + (size_t)i * D + lane + 32 * q), u[q]);
  }
#pragma unroll
  for (int q = 0; q < kCnnMaxD / 32; ++q) u[q] = tanhf(u[q]);
}

__device__ __forceinline__ WarpAct warp_act(float* base, int wib, int D, int extra) {
  float* b = base + (size_t)wib * (act_floats(D) + extra);
  return WarpAct{b, b + 2 * D, b + 6 * D};
}

// ---- P1: forward to u, S = sum u^2 ------------------------------------------------------------
__global__ void __launch_bounds__(kCnnThreads) cnn_p1_kernel(const CnnParams p) {
  extern __shared__ float smem[];
  const int D = p.D, lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  float* small = smem;
  for (int k = threadIdx.x; k < 2 * D + kSmall; k += kCnnThreads) small[k] = p.theta[k];
  __syncthreads();
  const WarpAct act = warp_act(smem + 2 * D + kSmall, wib, D, 0);
  float ssum = 0.f;
  for (int i = blockIdx.x * kCnnWarps + wib; i < p.n; i += gridDim.x * kCnnWarps) {
    const float* arow = p.attr_var + (size_t)__ldg(p.ia + i) * p.attr_stride;
    const float* vrow = p.val_var + (size_t)__ldg(p.iv + i) * p.val_stride;
    const float as = row_scale(arow, D, p.attr_norm, lane), vs = row_scale(vrow, D, p.val_norm, lane);
    float inv_n[4], nsum[4];
    cnn_forward(p, small, act, arow, as, vrow, vs, lane, inv_n, nsum);
    float u[kCnnMaxD / 32];
    cnn_dense(p, act.c2, lane, u);
#pragma unroll
    for (int q = 0; q < kCnnMaxD / 32; ++q)
      if (lane + 32 * q < D) {
        p.U[(size_t)i * D + lane + 32 * q] = u[q];
        ssum = fmaf(u[q], u[q], ssum);
      }
    __syncwarp();
  }
  ssum = warp_sum(ssum);
  if (lane == 0 && ssum != 0.f) atomicAdd(p.red, (double)ssum);
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
  if (lane == 0) {
    if (loss_local != 0.f && p.loss) atomicAdd(p.loss, (double)loss_local / 32.0);
    if (t_local != 0.f) atomicAdd(p.red + 1, (double)t_local);
  }
}

// ---- P3: backward through dense, norm, conv2, conv1, BN -------------------------------------------
__global__ void __launch_bounds__(kCnnThreads) cnn_p3_kernel(const CnnParams p) {
  extern __shared__ float smem[];
  const int D = p.D, lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const CnnLayout L{D};
  float* small = smem;
  for (int k = threadIdx.x; k < 2 * D + kSmall; k += kCnnThreads) small[k] = p.theta[k];
  __syncthreads();
  // per warp: activations + G2 [2][D][2] + G1 [2][D][2] + gpre [D] + gx0 [2][D]
  const int extra = 4 * D + 4 * D + D + 2 * D;
  const WarpAct act = warp_act(smem + 2 * D + kSmall, wib, D, extra);
  float* G2 = act.c2 + 4 * D;
  float* G1 = G2 + 4 * D;
  float* gpre = G1 + 4 * D;
  float* gx0 = gpre + D;
  const float* k1 = small + 2 * D;
  const float* k2 = k1 + 18;
  const float* gamma = small;
  const float S = (float)p.red[0], T = (float)p.red[1];
  const float rS = rsqrtf(fmaxf(S, kNormEps));
  const float bn = rsqrtf(1.f + kBnEps);
  float* attr_grad = p.attr_grad ? p.attr_grad + (size_t)(blockIdx.x % (unsigned)p.attr_rep) * p.attr_rep_floats : nullptr;
  float g_small[kSmall];  // per-lane partial sums of k1, b1, k2, b2 gradients
#pragma unroll
  for (int k = 0; k < kSmall; ++k) g_small[k] = 0.f;
  float g_gamma[kCnnMaxD / 32], g_beta[kCnnMaxD / 32];
#pragma unroll
  for (int q = 0; q < kCnnMaxD / 32; ++q) g_gamma[q] = g_beta[q] = 0.f;

  for (int i = blockIdx.x * kCnnWarps + wib; i < p.n; i += gridDim.x * kCnnWarps) {
    const int32_t a = __ldg(p.ia + i), h = __ldg(p.ih + i);
    const float* arow = p.attr_var + (size_t)a * p.attr_stride;
    const float* vrow = p.val_var + (size_t)__ldg(p.iv + i) * p.val_stride;
    const float* hrow = p.ent_var + (size_t)h * p.ent_stride;
    const float as = row_scale(arow, D, p.attr_norm, lane), vs = row_scale(vrow, D, p.val_norm, lane);
    const float hs = row_scale(hrow, D, p.ent_norm, lane);
    float inv_n[4], nsum[4];
    cnn_forward(p, small, act, arow, as, vrow, vs, lane, inv_n, nsum);
    const float coef = p.COEF[i];
    // g_pre = d loss / d (pre-activation of the dense layer)
#pragma unroll
    for (int q = 0; q < kCnnMaxD / 32; ++q) {
      const int j = lane + 32 * q;
      if (j < D) {
        const float u = p.U[(size_t)i * D + j];
        const float uh = u * rS;
        const float guh = -coef * (hrow[j] * hs - uh);
        const float gu = (S >= kNormEps) ? (guh - uh * T) * rS : guh * rS;
        const float gp = gu * (1.f - u * u);
        gpre[j] = gp;
        p.GP[(size_t)i * D + j] = gp;
      }
    }
    for (int k = lane; k < 4 * D; k += 32) p.Z[(size_t)i * 4 * D + k] = act.c2[k];
    __syncwarp();
    // g_z[k] = sum_j Wd[k][j] gpre[j]; then the width-normalisation backward
    float dots[4] = {0.f, 0.f, 0.f, 0.f};
    const float* wd = p.theta + L.wd();
    for (int k = lane; k < 4 * D; k += 32) {
      const float* wrow = wd + (size_t)k * D;
      float gz = 0.f;
      for (int j = 0; j < D; ++j) gz = fmaf(__ldg(wrow + j), gpre[j], gz);
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
  // parameter gradients of this warp -> global
#pragma unroll
  for (int k = 0; k < kSmall; ++k) {
    const float v = warp_sum(g_small[k]);
    if (lane == 0 && v != 0.f) atomicAdd(p.gtheta + L.k1() + k, v);
  }
#pragma unroll
  for (int q = 0; q < kCnnMaxD / 32; ++q) {
    const int w = lane + 32 * q;
    if (w < D) {
      if (g_gamma[q] != 0.f) atomicAdd(p.gtheta + L.gamma() + w, g_gamma[q]);
      if (g_beta[q] != 0.f) atomicAdd(p.gtheta + L.beta() + w, g_beta[q]);
    }
  }
}

// ---- P4: gWd[i][j] += sum_b Z[b][i] GP[b][j],  gbd[j] += sum_b GP[b][j] ------------------------------
__global__ void cnn_p4_kernel(const CnnParams p, int splits) {
  const int D = p.D, i = blockIdx.x, j = threadIdx.x;
  const CnnLayout L{D};
  if (j >= D) return;
  const int chunk = (p.n + splits - 1) / splits;
  const int b0 = blockIdx.y * chunk, b1 = min(p.n, b0 + chunk);
  float acc = 0.f, accb = 0.f;
  for (int b = b0; b < b1; ++b) {
    const float gp = p.GP[(size_t)b * D + j];
    acc = fmaf(p.Z[(size_t)b * 4 * D + i], gp, acc);
    accb += gp;
  }
  atomicAdd(p.gtheta + L.wd() + (size_t)i * D + j, acc);
  if (i == 0) atomicAdd(p.gtheta + L.bd() + j, accb);
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
extern "C" int64_t mke_attr_cnn_workspace_floats(int32_t n, int32_t dim) {
  return (int64_t)n * (6 * (int64_t)dim + 1) + 8;
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
  p.n = n; p.D = D; p.scale = scale; p.theta = theta; p.gtheta = gtheta;
  p.U = workspace;
  p.COEF = p.U + (size_t)n * D;
  p.Z = p.COEF + n;
  p.GP = p.Z + (size_t)n * 4 * D;
  // the two fp64 scalars live behind the float workspace, 8-byte aligned
  size_t off = (size_t)n * (6 * (size_t)D + 1);
  off = (off + 1) & ~(size_t)1;
  p.red = reinterpret_cast<double*>(workspace + off);
  p.loss = loss_accum;
  cudaStream_t s = (cudaStream_t)stream;
  if (cudaError_t e = cudaMemsetAsync(p.red, 0, 2 * sizeof(double), s)) return cuda_fail(e, "cudaMemsetAsync");
  int blocks = (n + kCnnWarps - 1) / kCnnWarps;
  const int full = sm_count() * 4;
  if (blocks > full) blocks = full;
  const size_t smem1 = (size_t)(2 * D + kSmall + kCnnWarps * (10 * D)) * sizeof(float);
  const size_t smem3 = (size_t)(2 * D + kSmall + kCnnWarps * (10 * D + 11 * D)) * sizeof(float);
  if (smem3 > 48 * 1024) {
    if (cudaError_t e = cudaFuncSetAttribute(cnn_p3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3))
      return cuda_fail(e, "cudaFuncSetAttribute");
  }
  cnn_p1_kernel<<<blocks, kCnnThreads, smem1, s>>>(p);
  MKE_CHECK_LAUNCH("cnn_p1_kernel");
  cnn_p2_kernel<<<blocks, kCnnThreads, 0, s>>>(p);
  MKE_CHECK_LAUNCH("cnn_p2_kernel");
  cnn_p3_kernel<<<blocks, kCnnThreads, smem3, s>>>(p);
  MKE_CHECK_LAUNCH("cnn_p3_kernel");
  const int splits = n >= 2048 ? 8 : 1;
  cnn_p4_kernel<<<dim3(4 * D, splits), ((D + 31) / 32) * 32, 0, s>>>(p, splits);
  MKE_CHECK_LAUNCH("cnn_p4_kernel");
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
