// Phase-2 row helpers shared by apply_adagrad_q8_kernel (mke_apply.cu) and the persistent step
// kernel (mke_rel_persist.cu): quarter-warp row loads/stores, ApplyTable, apply_one_row.
#pragma once
#include <cstdlib>
#include "mke_common.cuh"

namespace mke {

// ---- quarter-warp layout (default for strides 32/64/80/104/128) -------------------------------
// Lane `sub` of a quarter owns FPL = stride/8 floats of a row (same layout as mke_rel_q8.cu), so
// the three reads and three writes of a row are full 128-byte segments and one 3-step shuffle
// reduction of two values replaces two 5-step warp ones.  All loads of a row are issued before
// first use.
constexpr uint32_t kFullMask = 0xffffffffu;

template <int FPL>
__device__ __forceinline__ void q_load(const float* __restrict__ row, int sub, float (&x)[FPL]) {
  constexpr int NV4 = FPL / 4, REM = FPL % 4;
#pragma unroll
  for (int c = 0; c < NV4; ++c) {
    const float4 v = *reinterpret_cast<const float4*>(row + (c * 8 + sub) * 4);
    x[4 * c] = v.x; x[4 * c + 1] = v.y; x[4 * c + 2] = v.z; x[4 * c + 3] = v.w;
  }
  const float* tail = row + NV4 * 32 + REM * sub;
  if constexpr (REM == 2) {
    const float2 v = *reinterpret_cast<const float2*>(tail);
    x[4 * NV4] = v.x; x[4 * NV4 + 1] = v.y;
  } else {
#pragma unroll
    for (int k = 0; k < REM; ++k) x[4 * NV4 + k] = tail[k];
  }
}
template <int FPL>
__device__ __forceinline__ void q_store(float* __restrict__ row, int sub, const float (&x)[FPL]) {
  constexpr int NV4 = FPL / 4, REM = FPL % 4;
#pragma unroll
  for (int c = 0; c < NV4; ++c)
    *reinterpret_cast<float4*>(row + (c * 8 + sub) * 4) =
        make_float4(x[4 * c], x[4 * c + 1], x[4 * c + 2], x[4 * c + 3]);
  float* tail = row + NV4 * 32 + REM * sub;
  if constexpr (REM == 2) {
    *reinterpret_cast<float2*>(tail) = make_float2(x[4 * NV4], x[4 * NV4 + 1]);
  } else {
#pragma unroll
    for (int k = 0; k < REM; ++k) tail[k] = x[4 * NV4 + k];
  }
}

// streaming variants (ld/st.global.cs: evict-first in L1 and L2) for data that is touched once per
// step -- the Adagrad accumulator -- so that it does not push gradient / variable rows out of L2
template <int FPL>
__device__ __forceinline__ void q_load_cs(const float* __restrict__ row, int sub, float (&x)[FPL]) {
  constexpr int NV4 = FPL / 4, REM = FPL % 4;
#pragma unroll
  for (int c = 0; c < NV4; ++c) {
    const float4 v = __ldcs(reinterpret_cast<const float4*>(row + (c * 8 + sub) * 4));
    x[4 * c] = v.x; x[4 * c + 1] = v.y; x[4 * c + 2] = v.z; x[4 * c + 3] = v.w;
  }
  const float* tail = row + NV4 * 32 + REM * sub;
  if constexpr (REM == 2) {
    const float2 v = __ldcs(reinterpret_cast<const float2*>(tail));
    x[4 * NV4] = v.x; x[4 * NV4 + 1] = v.y;
  } else {
#pragma unroll
    for (int k = 0; k < REM; ++k) x[4 * NV4 + k] = __ldcs(tail + k);
  }
}
template <int FPL>
__device__ __forceinline__ void q_store_cs(float* __restrict__ row, int sub, const float (&x)[FPL]) {
  constexpr int NV4 = FPL / 4, REM = FPL % 4;
#pragma unroll
  for (int c = 0; c < NV4; ++c)
    __stcs(reinterpret_cast<float4*>(row + (c * 8 + sub) * 4),
           make_float4(x[4 * c], x[4 * c + 1], x[4 * c + 2], x[4 * c + 3]));
  float* tail = row + NV4 * 32 + REM * sub;
  if constexpr (REM == 2) {
    __stcs(reinterpret_cast<float2*>(tail), make_float2(x[4 * NV4], x[4 * NV4 + 1]));
  } else {
#pragma unroll
    for (int k = 0; k < REM; ++k) __stcs(tail + k, x[4 * NV4 + k]);
  }
}

struct ApplyTable {
  float* var;
  float* grad;
  uint8_t* touched;
  float* acc;
  int rows;
  int normalised;
  float lr;
  int replicas;  // gradient copies to sum and re-zero (mke_table_t.grad_replicas), >= 1
  int hint;      // bit 0: accumulator rows streamed (evict-first); bit 1: variable rows streamed too
};
inline int apply_hint() {
  static const int h = getenv("MKE_APPLY_HINT") ? atoi(getenv("MKE_APPLY_HINT")) : 0;
  return h;
}
inline ApplyTable apply_table(const mke_table_t* t, float* acc, float lr) {
  // a row-sharded table is swept shard by shard: this rank's rows only
  return ApplyTable{t->var, t->grad, t->touched, acc, table_local_rows(t), t->normalised, lr,
                    t->grad_replicas > 1 ? t->grad_replicas : 1, apply_hint()};
}

// Normalise-backward + Adagrad for the row at float offset `off`, executed by a quarter (lanes with
// on == false run along for the shuffles and write nothing).
template <int FPL>
__device__ __forceinline__ void apply_one_row(const ApplyTable& T, size_t off, bool on, int sub) {
  constexpr int stride = FPL * 8;
  const size_t rep_floats = (size_t)T.rows * stride;
  float g[FPL], v[FPL], a[FPL];
  q_load<FPL>(T.grad + off, sub, g);
  if (T.hint & 2) q_load_cs<FPL>(T.var + off, sub, v); else q_load<FPL>(T.var + off, sub, v);
  if (T.hint & 1) q_load_cs<FPL>(T.acc + off, sub, a); else q_load<FPL>(T.acc + off, sub, a);
  // gradient copies of small hot tables (mke_table_t.grad_replicas)
  for (int rep = 1; rep < T.replicas; ++rep) {
    float g2[FPL];
    q_load<FPL>(T.grad + rep * rep_floats + off, sub, g2);
#pragma unroll
    for (int k = 0; k < FPL; ++k) g[k] += g2[k];
  }
  float inv = 1.f, coef = 0.f;
  if (T.normalised) {
    float ss = 0.f, vg = 0.f;
#pragma unroll
    for (int k = 0; k < FPL; ++k) {
      ss = fmaf(v[k], v[k], ss);
      vg = fmaf(v[k], g[k], vg);
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      ss += __shfl_xor_sync(kFullMask, ss, o);
      vg += __shfl_xor_sync(kFullMask, vg, o);
    }
    // y = v * rsqrt(max(|v|^2, eps)); the max() routes no gradient to |v|^2 below eps
    inv = rsqrtf(fmaxf(ss, kNormEps));
    coef = (ss >= kNormEps) ? vg * inv * inv : 0.f;
  }
#pragma unroll
  for (int k = 0; k < FPL; ++k) {
    const float gv = (g[k] - v[k] * coef) * inv;
    a[k] = fmaf(gv, gv, a[k]);
    // var -= grad * lr * rsqrt(accum)   (ApplyAdagrad, no epsilon) [TF semantics]
    v[k] -= gv * T.lr * (a[k] > 0.f ? rsqrtf(a[k]) : 0.f);
    g[k] = 0.f;
  }
  if (on) {
    if (T.hint & 2) q_store_cs<FPL>(T.var + off, sub, v); else q_store<FPL>(T.var + off, sub, v);
    if (T.hint & 1) q_store_cs<FPL>(T.acc + off, sub, a); else q_store<FPL>(T.acc + off, sub, a);
    for (int rep = 0; rep < T.replicas; ++rep) q_store<FPL>(T.grad + rep * rep_floats + off, sub, g);
  }
}

// One chunk of CH (32 or 16) flagged-row candidates: the warp reads CH flag bytes, clears the set ones, then its
// four quarters take four flagged rows at a time (flagged rows are compacted by ballot, so all quarters stay
// busy whatever the touched fraction is).
template <int FPL, int CH = 32>
__device__ __forceinline__ void apply_flag_chunk(const ApplyTable& T, int base, int lane) {
  constexpr int stride = FPL * 8;
  const int sub = lane & 7;
  const int q = lane >> 3;
  const int my = base + lane;
  const bool flag = (lane < CH) && (my < T.rows) && (T.touched[my] != 0);
  uint32_t m = __ballot_sync(kFullMask, flag);
  if (flag) T.touched[my] = 0;
  while (m) {  // warp-uniform
    // quarter q takes the q-th lowest flagged row of the remaining ones
    uint32_t mm = m;
    int bit = -1;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int b = mm ? (__ffs(mm) - 1) : -1;
      if (k == q) bit = b;
      mm &= mm - 1;
    }
    m = mm;
    const bool on = bit >= 0;
    apply_one_row<FPL>(T, (size_t)(base + (on ? bit : 0)) * stride, on, sub);
  }
}
// Four consecutive rows of a table without flags (small tables: every row is visited), one per quarter.
template <int FPL>
__device__ __forceinline__ void apply_row4(const ApplyTable& T, int r0, int lane) {
  constexpr int stride = FPL * 8;
  const int row = r0 + (lane >> 3);
  const bool on = row < T.rows;
  apply_one_row<FPL>(T, (size_t)(on ? row : 0) * stride, on, lane & 7);
}

}  // namespace mke
