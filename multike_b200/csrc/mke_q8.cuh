// Quarter-warp row helpers shared by the phase-1 kernels (mke_rel_q8.cu, mke_rel_q8p.cu):
// per-lane row pieces, shuffle reductions, cp.async staging, gradient-row scatter.
#pragma once
#include "mke_rel.cuh"
#include "mke_sampler.cuh"

namespace mke {

__device__ __forceinline__ float qsum(float v) {
  v += __shfl_xor_sync(kFull, v, 4);
  v += __shfl_xor_sync(kFull, v, 2);
  v += __shfl_xor_sync(kFull, v, 1);
  return v;
}
__device__ __forceinline__ void qsum3(float& a, float& b, float& c) {
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) {
    a += __shfl_xor_sync(kFull, a, o);
    b += __shfl_xor_sync(kFull, b, o);
    c += __shfl_xor_sync(kFull, c, o);
  }
}
__device__ __forceinline__ void red_add_f2(float* p, float x, float y) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(x), "f"(y) : "memory");
}
__device__ __forceinline__ void red_add_f1(float* p, float x) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(x) : "memory");
}

// lane `sub` (0..7) of a quarter owns floats {32c + 4 sub .. +3 : c < FPL/4} and the FPL%4 floats
// at 32 (FPL/4) + (FPL%4) sub of a row
template <int FPL>
__device__ __forceinline__ void load_row(const float* __restrict__ row, int sub, float (&x)[FPL]) {
  constexpr int NV4 = FPL / 4, REM = FPL % 4;
#pragma unroll
  for (int c = 0; c < NV4; ++c) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(row) + c * 8 + sub);
    x[4 * c] = v.x;
    x[4 * c + 1] = v.y;
    x[4 * c + 2] = v.z;
    x[4 * c + 3] = v.w;
  }
  const float* tail = row + NV4 * 32 + REM * sub;
  if constexpr (REM == 2) {
    const float2 v = __ldg(reinterpret_cast<const float2*>(tail));
    x[4 * NV4] = v.x;
    x[4 * NV4 + 1] = v.y;
  } else {
#pragma unroll
    for (int k = 0; k < REM; ++k) x[4 * NV4 + k] = __ldg(tail + k);
  }
}

// grad_row += s * x
template <int FPL>
__device__ __forceinline__ void red_row(float* __restrict__ row, int sub, const float (&x)[FPL],
                                        float s) {
  constexpr int NV4 = FPL / 4, REM = FPL % 4;
#pragma unroll
  for (int c = 0; c < NV4; ++c)
    red_add_f4(row + (c * 8 + sub) * 4,
               make_float4(x[4 * c] * s, x[4 * c + 1] * s, x[4 * c + 2] * s, x[4 * c + 3] * s));
  float* tail = row + NV4 * 32 + REM * sub;
  if constexpr (REM == 2) {
    red_add_f2(tail, x[4 * NV4] * s, x[4 * NV4 + 1] * s);
  } else {
#pragma unroll
    for (int k = 0; k < REM; ++k) red_add_f1(tail + k, x[4 * NV4 + k] * s);
  }
}

template <int FPL>
__device__ __forceinline__ float sumsq(const float (&x)[FPL]) {
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < FPL; ++k) s = fmaf(x[k], x[k], s);
  return s;
}

// A negative whose side differs from the side of negative 0 of its positive (only possible when
// rounds with different coins contributed, or in caller-supplied batches).  Rare: everything
// is re-read from the tables and reduced straight into the gradient rows.  Executed by the whole
// warp (shuffles use the full mask); quarters with on == false compute and discard.
template <int FPL>
static __device__ __noinline__ float odd_negative(const float* vh, const float* vr, const float* vt,
                                                  const float* ve, float* gh, float* gr, float* gt, float* ge,
                                                  int ent_norm, int rel_norm, bool head_side, bool on,
                                                  int sub) {
  // One float at a time, three sweeps over the rows (norms, distance, reductions): this path runs
  // for about 0.2 % of the positives and must not set the register budget of the kernels.
  constexpr int NV4 = FPL / 4, REM = FPL % 4;
  auto off = [&](int k) { return k < 4 * NV4 ? 32 * (k >> 2) + 4 * sub + (k & 3) : 32 * NV4 + REM * sub + (k - 4 * NV4); };
  float sh = 0.f, sr = 0.f, st = 0.f, se = 0.f;
#pragma unroll 1
  for (int k = 0; k < FPL; ++k) {
    const int o = off(k);
    const float a = __ldg(vh + o), b = __ldg(vr + o), c = __ldg(vt + o), d = __ldg(ve + o);
    sh = fmaf(a, a, sh);
    sr = fmaf(b, b, sr);
    st = fmaf(c, c, st);
    se = fmaf(d, d, se);
  }
  qsum3(sh, sr, st);
  se = qsum(se);
  const float ih = ent_norm ? rsqrtf(fmaxf(sh, kNormEps)) : 1.f;
  const float ir = rel_norm ? rsqrtf(fmaxf(sr, kNormEps)) : 1.f;
  const float it = ent_norm ? rsqrtf(fmaxf(st, kNormEps)) : 1.f;
  const float ie = ent_norm ? rsqrtf(fmaxf(se, kNormEps)) : 1.f;
  auto dist = [&](int o) {
    const float rr = __ldg(vr + o) * ir;
    return head_side ? (fmaf(__ldg(ve + o), ie, rr) - __ldg(vt + o) * it)
                     : (fmaf(__ldg(vh + o), ih, rr) - __ldg(ve + o) * ie);
  };
  float sn = 0.f;
#pragma unroll 1
  for (int k = 0; k < FPL; ++k) {
    const float nd = dist(off(k));
    sn = fmaf(nd, nd, sn);
  }
  sn = qsum(sn);
  float lneg, sg;
  softplus_sigmoid(-sn, lneg, sg);
  if (!on) return 0.f;
  const float cn = -2.f * sg;
#pragma unroll 1
  for (int k = 0; k < FPL; ++k) {
    const int o = off(k);
    const float g = cn * dist(o);
    red_add_f1(gr + o, g);
    red_add_f1(ge + o, head_side ? g : -g);
    if (head_side)
      red_add_f1(gt + o, -g);
    else
      red_add_f1(gh + o, g);
  }
  return lneg;
}

// ---- per-lane staging of rows in shared memory (cp.async / LDGSTS) --------------------------
// A lane copies exactly the pieces of a row it will later read back, so no cross-lane
// visibility is involved: cp.async.wait_group by the lane itself is the only synchronisation,
// and no register is tied up while a row is in flight.
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16_cg(uint32_t dst, const void* src) {  // L2 only: no L1 allocation
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

constexpr int kSlots = 3;  // staging slots per lane: h, r, t rows, then a ring of negative rows

// Warp staging area: [slot][16-byte chunk][lane] then [slot][lane] tails -> every LDS/LDGSTS of a
// warp touches 32 consecutive pieces (conflict free).
template <int FPL, int SLOTS = kSlots>
struct Stage {
  static constexpr int NV4 = FPL / 4, REM = FPL % 4;
  static constexpr int kTailBase = SLOTS * NV4 * 32 * 16;
  static constexpr int kBytes = kTailBase + SLOTS * 32 * REM * 4;
  uint32_t base;  // shared-space address of this lane's first piece
  int lane;
  __device__ __forceinline__ uint32_t chunk(int slot, int c) const {
    return base + ((slot * NV4 + c) * 32 + lane) * 16;
  }
  __device__ __forceinline__ uint32_t tail(int slot) const {
    return base + kTailBase + (slot * 32 + lane) * (REM * 4);
  }
  template <bool CG = false>
  __device__ __forceinline__ void issue(int slot, const float* __restrict__ row, int sub) const {
#pragma unroll
    for (int c = 0; c < NV4; ++c) {
      if constexpr (CG)
        cp_async16_cg(chunk(slot, c), row + (c * 8 + sub) * 4);
      else
        cp_async16(chunk(slot, c), row + (c * 8 + sub) * 4);
    }
    const float* t = row + NV4 * 32 + REM * sub;
    if constexpr (REM == 2) cp_async8(tail(slot), t);
    if constexpr (REM == 1) cp_async4(tail(slot), t);
    if constexpr (REM == 3) {
      cp_async4(tail(slot), t);
      cp_async4(tail(slot) + 4, t + 1);
      cp_async4(tail(slot) + 8, t + 2);
    }
  }
  // 16-byte piece c (< NV4) of this lane's part of the row in `slot`
  __device__ __forceinline__ void read4(int slot, int c, float (&v)[4]) const {
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3])
                 : "r"(chunk(slot, c)));
  }
  // the FPL % 4 floats after the 16-byte pieces
  __device__ __forceinline__ void read_tail(int slot, float (&v)[REM > 0 ? REM : 1]) const {
    if constexpr (REM == 2)
      asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v[0]), "=f"(v[1]) : "r"(tail(slot)));
    if constexpr (REM == 1 || REM == 3) {
#pragma unroll
      for (int k = 0; k < REM; ++k) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v[k]) : "r"(tail(slot) + 4 * k));
    }
  }
  __device__ __forceinline__ void read(int slot, float (&x)[FPL]) const {
#pragma unroll
    for (int c = 0; c < NV4; ++c) {
      float4 v;
      asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                   : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                   : "r"(chunk(slot, c)));
      x[4 * c] = v.x;
      x[4 * c + 1] = v.y;
      x[4 * c + 2] = v.z;
      x[4 * c + 3] = v.w;
    }
    if constexpr (REM == 2)
      asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(x[4 * NV4]), "=f"(x[4 * NV4 + 1]) : "r"(tail(slot)));
    if constexpr (REM == 1 || REM == 3) {
#pragma unroll
      for (int k = 0; k < REM; ++k)
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x[4 * NV4 + k]) : "r"(tail(slot) + 4 * k));
    }
  }
};

// ---- gradient rows back to HBM ---------------------------------------------------------------
// BULK == false: red.global.add.v4/v2.f32 from registers (24 LSU lane-operations per 80-float row).
// BULK == true : the quarter writes the row into its 320-byte shared buffer and its leader lane
//                hands it to the TMA engine (cp.reduce.async.bulk .add.f32, SASS UBLKRED): the
//                element-wise add happens at L2 like RED, but off the LSU, which the gather side
//                (LDGSTS/LDS) keeps busy.  profiles/ has the A/B measurement.
template <int FPL, bool BULK>
struct RowScatter {
  uint32_t buf;  // shared-space address of this quarter's row buffer
  uint32_t qmask;
  int sub;
  // called quarter-uniformly (all 8 lanes of the quarter or none)
  __device__ __forceinline__ void add(float* __restrict__ grad_row, const float (&x)[FPL], float s) const {
    if constexpr (!BULK) {
      red_row<FPL>(grad_row, sub, x, s);
    } else {
      constexpr int NV4 = FPL / 4, REM = FPL % 4;
      // the engine must have finished READING the previous row out of the buffer
      if (sub == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      __syncwarp(qmask);
#pragma unroll
      for (int c = 0; c < NV4; ++c)
        asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(buf + (c * 8 + sub) * 16),
                     "f"(x[4 * c] * s), "f"(x[4 * c + 1] * s), "f"(x[4 * c + 2] * s), "f"(x[4 * c + 3] * s)
                     : "memory");
      const uint32_t tail = buf + NV4 * 128 + REM * 4 * sub;
      if constexpr (REM == 2)
        asm volatile("st.shared.v2.f32 [%0], {%1,%2};" ::"r"(tail), "f"(x[4 * NV4] * s), "f"(x[4 * NV4 + 1] * s)
                     : "memory");
      if constexpr (REM == 1 || REM == 3) {
#pragma unroll
        for (int k = 0; k < REM; ++k)
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(tail + 4 * k), "f"(x[4 * NV4 + k] * s) : "memory");
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> async proxy
      __syncwarp(qmask);
      if (sub == 0) {
        asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(grad_row),
                     "r"(buf), "n"(FPL * 32)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
  }
  __device__ __forceinline__ void drain() const {
    if constexpr (BULK) {
      if (sub == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
  }
};

}  // namespace mke
