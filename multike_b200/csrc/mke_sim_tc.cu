// Similarity search on the tensor cores: the [n1, n2] inner products of
// base/similarity.py:9-52 sim(metric='inner') behind greedy_alignment / calculate_rank (base/alignment.py:8-79,
// 141-163) and find_neighbours (base/batch.py:119-150), as tcgen05 (UMMA) tiles fed by TMA with the consumer of the
// sims -- rank of the gold column + arg-max, or the rows of sims the exact top-k select reads -- fused into the
// TMEM epilogue.  The similarity matrix of the evaluator (14 GB at test size) never exists.
//
// Precision: 3xTF32 like the auto-encoder GEMM (mke_gemm.cu).  Prepared rows are split into hi (TF32) + lo (exact
// remainder); sim = a_hi b_hi (accumulator 1) + a_hi b_lo + a_lo b_hi (accumulator 2), summed in fp32 by the epilogue;
// the dropped a_lo b_lo terms are < 2^-22 of |a||b|.  K is at most 128, so the tensor core's truncating accumulation
// stays below fp32 rounding without the GEMM's chunking.  What the rank rules need is that EQUAL ROWS GIVE BIT-EQUAL
// SIMS wherever they sit in a tile: every element of a UMMA tile is the same dot-product circuit over the same k order,
// and the gold column's score is produced by this very kernel (mode GOLD: the tile of every row block with the rows
// B[gold[i]], diagonal kept), not by a CUDA-core chain -- tests/test_gpu_sim.py holds the exact-tie cases.
//
// One CTA = one 128-row tile of A (both parts resident in shared memory for the whole launch) against a share of the
// 128-row tiles of B, streamed per 32-float k-block through a TMA ring:
//   warp 0, one lane : TMA producer (A once; per B tile nkb stages of B_hi | B_lo, 128 rows x 128 B, SWIZZLE_128B)
//   warp 1, one lane : MMA issuer, 3 tcgen05.mma.kind::tf32 (M = N = 128, K = 8) per k-step into one of two
//                      accumulator pairs in TMEM (2 x 256 columns): tile j+1 is multiplied while tile j is consumed
//   warps 2-5        : epilogue, thread = row of the tile: tcgen05.ld 32 columns at a time of both accumulators, then
//                      RANK: count of columns ranking before the gold one + running arg-max (registers; one atomic
//                      pair per row at the end), STORE: the row of sims, GOLD: the diagonal.
#include "mke_umma.cuh"

namespace mke {

constexpr int kTcThreads = 192;
constexpr int kTcTile = 128;
constexpr int kTcBoxBytes = kTcTile * 32 * 4;  // one operand box: 128 rows x one 128-byte swizzle row
constexpr int kTcSmemBudget = 224 * 1024;     // A (nkb x 2 boxes) + ring (stages x 2 boxes)

enum { kTcRank = 0, kTcStore = 1, kTcGold = 2, kTcFilter = 3 };

struct SimTcParams {
  int n1, n2;        // rows of A / of B
  int gold_n2;       // GOLD: rows of the matrix the gold indices point into (B here is the gathered copy)
  int nkb;           // k-blocks of 32 floats
  int ksteps_last;   // k-steps (of 8) of the last k-block that hold data
  int stages;        // ring depth
  int splits;        // column splits: CTA (rb, cs) owns B tiles cs, cs + splits, ...
  int row_base, rows;
  const int32_t* gold;
  float* gold_score;           // GOLD: written; RANK: read
  int32_t* rank;
  unsigned long long* best;
  float* out;
  size_t out_pitch;
  // FILTER: columns whose sim reaches the row's threshold are appended to the row's candidate list
  const float* tau;     // [rows]
  uint32_t* cand_cnt;   // [rows] (may exceed cap: the row then takes the exact path)
  uint2* cand;          // [rows][cap] (column, sim bits)
  int cap;
};

__device__ __forceinline__ uint32_t tc_ord_key(float f) {
  const uint32_t u = __float_as_uint(f);
  return u ^ ((u >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}

template <int MODE>
__global__ void __launch_bounds__(kTcThreads, 1)
    sim_tc_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                  const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                  const SimTcParams p) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  unsigned char* const a_tiles = smem;                                   // [nkb][hi | lo] boxes
  unsigned char* const ring = smem + (size_t)p.nkb * 2 * kTcBoxBytes;    // [stages][hi | lo] boxes
  uint64_t* bars = (uint64_t*)(ring + (size_t)p.stages * 2 * kTcBoxBytes);  // a_full, full[S], empty[S], accum_full[2], accum_empty[2]
  const int S = p.stages;
  uint32_t* tmem_slot = (uint32_t*)(bars + 1 + 2 * S + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rb = blockIdx.x / p.splits, cs = blockIdx.x - rb * p.splits;
  const int row0 = p.row_base + rb * kTcTile;
  const int row_end = p.row_base + p.rows < p.n1 ? p.row_base + p.rows : p.n1;
  const int ntiles = (p.n2 + kTcTile - 1) / kTcTile;
  // GOLD: only the tile that holds this row block's own rows of the gathered B
  const int jt_first = MODE == kTcGold ? row0 / kTcTile : cs;
  const int jt_step = MODE == kTcGold ? ntiles : p.splits;
  auto a_full = [&]() { return smem_u32(bars); };
  auto full = [&](int s) { return smem_u32(bars + 1 + s); };
  auto empty = [&](int s) { return smem_u32(bars + 1 + S + s); };
  auto accum_full = [&](int b) { return smem_u32(bars + 1 + 2 * S + b); };
  auto accum_empty = [&](int b) { return smem_u32(bars + 1 + 2 * S + 2 + b); };

  if (warp == 0 && lane == 0) {
    mbar_init(a_full(), 1);
    for (int s = 0; s < S; ++s) {
      mbar_init(full(s), 1);
      mbar_init(empty(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(accum_full(b), 1);
      mbar_init(accum_empty(b), 4);  // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {  // ---- TMA producer ----
      mbar_expect_tx(a_full(), (uint32_t)(p.nkb * 2 * kTcBoxBytes));
      for (int kb = 0; kb < p.nkb; ++kb) {
        tma_load_2d(&map_a_hi, a_full(), smem_u32(a_tiles + (size_t)(2 * kb) * kTcBoxBytes), kb * 32, row0);
        tma_load_2d(&map_a_lo, a_full(), smem_u32(a_tiles + (size_t)(2 * kb + 1) * kTcBoxBytes), kb * 32, row0);
      }
      int it = 0;
      for (int jt = jt_first; jt < ntiles; jt += jt_step) {
        for (int kb = 0; kb < p.nkb; ++kb, ++it) {
          const int s = it % S;
          mbar_wait(empty(s), (((uint32_t)(it / S)) & 1u) ^ 1u);
          mbar_expect_tx(full(s), 2 * kTcBoxBytes);
          const uint32_t dst = smem_u32(ring + (size_t)s * 2 * kTcBoxBytes);
          tma_load_2d(&map_b_hi, full(s), dst, kb * 32, jt * kTcTile);
          tma_load_2d(&map_b_lo, full(s), dst + kTcBoxBytes, kb * 32, jt * kTcTile);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ---- MMA issuer ----
      mbar_wait(a_full(), 0u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      int it = 0, t = 0;
      for (int jt = jt_first; jt < ntiles; jt += jt_step, ++t) {
        const int buf = t & 1;
        mbar_wait(accum_empty(buf), (((uint32_t)(t >> 1)) & 1u) ^ 1u);  // the epilogue has drained this pair
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d_main = tmem_base + (uint32_t)(buf * 256), d_cross = d_main + 128u;
        for (int kb = 0; kb < p.nkb; ++kb, ++it) {
          const int s = it % S;
          mbar_wait(full(s), ((uint32_t)(it / S)) & 1u);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t a_base = smem_u32(a_tiles + (size_t)(2 * kb) * kTcBoxBytes);
          const uint32_t b_base = smem_u32(ring + (size_t)s * 2 * kTcBoxBytes);
          const uint64_t a_hi = umma_desc_k_sw128(a_base), a_lo = umma_desc_k_sw128(a_base + kTcBoxBytes);
          const uint64_t b_hi = umma_desc_k_sw128(b_base), b_lo = umma_desc_k_sw128(b_base + kTcBoxBytes);
          const int nk = kb == p.nkb - 1 ? p.ksteps_last : 4;
          for (int k = 0; k < nk; ++k) {
            const uint64_t adv = (uint64_t)((k * 8 * 4) >> 4);  // 8 TF32 = 32 bytes along K inside the swizzle row
            const uint32_t acc = (kb > 0 || k > 0) ? 1u : 0u;
            umma_tf32(d_main, a_hi + adv, b_hi + adv, acc);
            umma_tf32(d_cross, a_hi + adv, b_lo + adv, acc);
            umma_tf32(d_cross, a_lo + adv, b_hi + adv, 1u);
          }
          umma_commit(empty(s));
        }
        umma_commit(accum_full(buf));
      }
    }
  } else {  // ---- epilogue: warps 2..5, thread = row (TMEM lane) ----
    const int lane_base = (warp & 3) * 32;
    const int l = lane_base + lane;
    const int r = row0 + l;
    const bool row_ok = r < row_end;
    float sg = __int_as_float(0x7f800000);
    int gi = -1;
    if (MODE == kTcRank && row_ok) {
      sg = __ldg(p.gold_score + r);
      gi = p.gold ? __ldg(p.gold + r) : r;
    }
    int cnt = 0, bi = 0x7fffffff;
    float bs = __int_as_float(0xff800000);
    float diag = 0.f;
    const float tau_r = (MODE == kTcFilter && row_ok) ? __ldg(p.tau + (r - p.row_base)) : 0.f;
    float* const orow = MODE == kTcStore ? p.out + (size_t)(row_ok ? r - p.row_base : 0) * p.out_pitch : nullptr;
    int t = 0;
#pragma unroll 1
    for (int jt = jt_first; jt < ntiles; jt += jt_step, ++t) {
      const int buf = t & 1;
      mbar_wait(accum_full(buf), ((uint32_t)(t >> 1)) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem_base + ((uint32_t)lane_base << 16) + (uint32_t)(buf * 256);
#pragma unroll 1
      for (int q = 0; q < 4; ++q) {
        uint32_t v[32], u[32];
        tmem_ld32(taddr + q * 32, v);         // sum of a_hi b_hi
        tmem_ld32(taddr + 128 + q * 32, u);   // sum of a_hi b_lo + a_lo b_hi
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const int col0 = jt * kTcTile + q * 32;
        if (MODE == kTcRank) {
          if (col0 + 32 <= p.n2 && (gi < col0 || gi >= col0 + 32)) {  // the common case: no masking, no gold column here
            // stable descending order (base/alignment.py:148 argsort of -sim): a column ranks before the gold one if
            // its sim is larger, or equal with a smaller column index -- all 32 columns are on one side of the gold one
            if (gi >= col0 + 32) {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float s = __uint_as_float(v[j]) + __uint_as_float(u[j]);
                cnt += s >= sg ? 1 : 0;
                if (s > bs) {  // columns come in ascending order
                  bs = s;
                  bi = col0 + j;
                }
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float s = __uint_as_float(v[j]) + __uint_as_float(u[j]);
                cnt += s > sg ? 1 : 0;
                if (s > bs) {
                  bs = s;
                  bi = col0 + j;
                }
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int col = col0 + j;
              const bool cv = col < p.n2;
              const float s = __uint_as_float(v[j]) + __uint_as_float(u[j]);
              cnt += (cv && col != gi && (s > sg || (s == sg && col < gi))) ? 1 : 0;
              if (cv && s > bs) {
                bs = s;
                bi = col;
              }
            }
          }
        } else if (MODE == kTcStore) {
          if (row_ok) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const int col = col0 + j;
              if ((size_t)col + 4 <= p.out_pitch) {  // pitch is a multiple of 4 >= n2: pad columns may be written
                float4 w;
                w.x = __uint_as_float(v[j]) + __uint_as_float(u[j]);
                w.y = __uint_as_float(v[j + 1]) + __uint_as_float(u[j + 1]);
                w.z = __uint_as_float(v[j + 2]) + __uint_as_float(u[j + 2]);
                w.w = __uint_as_float(v[j + 3]) + __uint_as_float(u[j + 3]);
                __stcs(reinterpret_cast<float4*>(orow + col), w);
              }
            }
          }
        } else if (MODE == kTcFilter) {
          if (row_ok) {
            uint32_t hit = 0u;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float s = __uint_as_float(v[j]) + __uint_as_float(u[j]);
              v[j] = __float_as_uint(s);
              hit |= (col0 + j < p.n2 && s >= tau_r) ? (1u << j) : 0u;
            }
            if (hit != 0u) {
              const int c = __popc(hit);
              const uint32_t pos = atomicAdd(p.cand_cnt + (r - p.row_base), (uint32_t)c);
              if (pos + (uint32_t)c <= (uint32_t)p.cap) {
                uint2* dst = p.cand + (size_t)(r - p.row_base) * p.cap + pos;
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if ((hit >> j) & 1u) *dst++ = make_uint2((uint32_t)(col0 + j), v[j]);
              }
            }
          }
        } else {  // GOLD: column l of the tile
          if ((l >> 5) == q) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j == (l & 31)) diag = __uint_as_float(v[j]) + __uint_as_float(u[j]);
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(accum_empty(buf)) : "memory");
    }
    if (MODE == kTcRank && row_ok) {
      if (cnt != 0) atomicAdd(p.rank + r, cnt);
      if (bi != 0x7fffffff)
        atomicMax(p.best + r, ((unsigned long long)tc_ord_key(bs) << 32) | (0xFFFFFFFFu - (uint32_t)bi));
    }
    if (MODE == kTcGold && row_ok) {
      const int g = p.gold ? __ldg(p.gold + r) : r;
      p.gold_score[r] = (g >= 0 && g < p.gold_n2) ? diag : __int_as_float(0x7f800000);  // gold outside: nothing ranks before it
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
}

// One warp per row: gather by idx, optionally x / ||x|| (sklearn.preprocessing.normalize as used by
// base/similarity.py:31-33: a zero row stays zero), zero the pad columns, split into the TF32 part and the remainder.
__global__ void sim_tc_prepare_kernel(const float* __restrict__ src, const int32_t* __restrict__ idx, int n, int stride,
                                      int dim, int normalize, float* __restrict__ hi, float* __restrict__ lo, int ws) {
  const int row = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  const float* s = src + (size_t)(idx ? __ldg(idx + row) : row) * stride;
  float ss = 0.f;
  for (int c = lane; c < dim; c += 32) {
    const float v = __ldg(s + c);
    ss = fmaf(v, v, ss);
  }
  ss = warp_sum(ss);
  float nrm = normalize ? sqrtf(ss) : 1.f;
  if (nrm == 0.f) nrm = 1.f;
  for (int c = lane; c < ws; c += 32) {
    const float v = c < dim ? __ldg(s + c) / nrm : 0.f;
    float h, l;
    split_tf32(v, h, l);
    hi[(size_t)row * ws + c] = h;
    lo[(size_t)row * ws + c] = l;
  }
}

// rows B[gold[i]] (zeros where gold is outside [0, n2)) next to each other, and the per-row outputs reset
__global__ void sim_tc_gather_gold_kernel(const float* __restrict__ b_hi, const float* __restrict__ b_lo,
                                          const int32_t* __restrict__ gold, int n1, int n2, int ws,
                                          float* __restrict__ g_hi, float* __restrict__ g_lo, int32_t* __restrict__ rank,
                                          unsigned long long* __restrict__ best) {
  const int row = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n1) return;
  const int g = gold ? __ldg(gold + row) : row;
  const bool ok = g >= 0 && g < n2;
  for (int c = lane; c < ws; c += 32) {
    g_hi[(size_t)row * ws + c] = ok ? __ldg(b_hi + (size_t)g * ws + c) : 0.f;
    g_lo[(size_t)row * ws + c] = ok ? __ldg(b_lo + (size_t)g * ws + c) : 0.f;
  }
  if (lane == 0) {
    rank[row] = 0;
    best[row] = 0ull;
  }
}

// column splits so that the grid fills whole waves of SMs (one CTA per SM: 224 KB of shared memory, all of TMEM)
static int sim_tc_splits(int row_blocks, int ntiles) {
  const int sms = sm_count();
  int best_s = 1;
  double best_eff = 0.0;
  for (int w = 1; w <= 16; ++w) {
    int s = (int)((long long)sms * w / row_blocks);
    if (s < 1) continue;
    if (s > ntiles) s = ntiles;
    if (ntiles / s < 4 && s > 1) break;  // keep the A tile load amortised over at least 4 tiles of B
    const long long ctas = (long long)row_blocks * s;
    const double eff = (double)ctas / (double)(((ctas + sms - 1) / sms) * sms);
    if (eff > best_eff + 0.02) {
      best_eff = eff;
      best_s = s;
    }
  }
  return best_s;
}

template <int MODE>
static int launch_sim_tc(const float* a_hi, const float* a_lo, const float* b_hi, const float* b_lo, int ws, SimTcParams p,
                         cudaStream_t stream) {
  CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo;
  if (int rc = make_map(&ma_hi, a_hi, p.n1, ws, ws)) return rc;
  if (int rc = make_map(&ma_lo, a_lo, p.n1, ws, ws)) return rc;
  if (int rc = make_map(&mb_hi, b_hi, p.n2, ws, ws)) return rc;
  if (int rc = make_map(&mb_lo, b_lo, p.n2, ws, ws)) return rc;
  p.nkb = (ws + 31) / 32;
  p.ksteps_last = (ws - 32 * (p.nkb - 1) + 7) / 8;
  p.stages = (kTcSmemBudget - p.nkb * 2 * kTcBoxBytes) / (2 * kTcBoxBytes);
  if (p.stages > 6) p.stages = 6;
  const int smem = (p.nkb + p.stages) * 2 * kTcBoxBytes + 1024 + 256;
  auto kern = sim_tc_kernel<MODE>;
  static int configured[4] = {0, 0, 0, 0};
  if (configured[MODE] < smem) {
    if (cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem))
      return cuda_fail(e, "cudaFuncSetAttribute(sim_tc_kernel)");
    configured[MODE] = smem;
  }
  const int row_blocks = (p.rows + kTcTile - 1) / kTcTile;
  const int ntiles = (p.n2 + kTcTile - 1) / kTcTile;
  p.splits = MODE == kTcGold ? 1 : sim_tc_splits(row_blocks, ntiles);
  kern<<<row_blocks * p.splits, kTcThreads, smem, stream>>>(ma_hi, ma_lo, mb_hi, mb_lo, p);
  MKE_CHECK_LAUNCH("sim_tc_kernel");
  return 0;
}

// ---- truncated-epsilon neighbour lists without the similarity matrix ------------------------------------------------
// find_neighbours (base/batch.py:141-150) keeps the k = 2 % most similar columns of each of 100 000 rows.  Instead of
// writing 40 GB of sims and selecting from them: (1) sims of every row with a SAMPLE of 4 096 columns (STORE tiles, 4 % of
// the work) give a per-row threshold that, with 3.4 sigma of margin, at least k columns of the full row reach; (2) the
// FILTER tiles append the (column, sim) pairs that reach it to a candidate list per row (about 1.3 k of them); (3) one
// block per row selects the exact top k of its candidates (radix select in shared memory) and emits them in ascending
// column order through a bitmap.  A row whose list came out short or overflowed is redone on the exact path.

__device__ __forceinline__ float tc_key_to_float(uint32_t key) {
  return __uint_as_float((key & 0x80000000u) ? (key ^ 0x80000000u) : ~key);
}

// key of the kth largest (kth >= 1) of n_keys keys in shared memory, and how many keys equal to it belong to the top kth
// (block of 256 threads; hist: 2048 words of shared memory)
__device__ void block_kth_key(const uint32_t* keys, int n_keys, unsigned kth, unsigned* hist, unsigned* s_pair, uint32_t& T,
                              unsigned& need_eq) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t prefix = 0u, mask = 0u;
  unsigned remaining = kth;
  for (int pass = 0; pass < 3; ++pass) {
    const int shift = pass == 0 ? 21 : (pass == 1 ? 10 : 0);
    const int nb = pass < 2 ? 2048 : 1024;
    for (int b = tid; b < nb; b += blockDim.x) hist[b] = 0u;
    __syncthreads();
    for (int i = tid; i < n_keys; i += blockDim.x) {
      const uint32_t key = keys[i];
      if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & (unsigned)(nb - 1)], 1u);
    }
    __syncthreads();
    if (warp == 0) {  // lane l owns bins [l per, (l + 1) per); suffix sums from the top bin down
      const int per = nb / 32;
      unsigned mine = 0u;
      for (int b = 0; b < per; ++b) mine += hist[lane * per + b];
      unsigned above = 0u;
      for (int l = 31; l > 0; --l) {
        const unsigned v = __shfl_sync(0xffffffffu, mine, l);
        if (l > lane) above += v;
      }
      if (above < remaining && remaining <= above + mine) {
        unsigned acc = above;
        for (int b = per - 1; b >= 0; --b) {
          const unsigned h = hist[lane * per + b];
          if (acc + h >= remaining) {
            s_pair[0] = (unsigned)(lane * per + b);
            s_pair[1] = acc;
            break;
          }
          acc += h;
        }
      }
    }
    __syncthreads();
    prefix |= s_pair[0] << shift;
    mask |= (unsigned)(nb - 1) << shift;
    remaining -= s_pair[1];
    __syncthreads();
  }
  T = prefix;
  need_eq = remaining;
}

constexpr int kTcSample = 4096;     // sampled columns
constexpr int kSelThreads = 256;

// tau[row] = the m-th largest of the row's sims with the sampled columns
__global__ void __launch_bounds__(kSelThreads) sim_tc_threshold_kernel(const float* __restrict__ ssims, int S, unsigned m,
                                                                        float* __restrict__ tau, uint32_t* __restrict__ cnt) {
  __shared__ uint32_t keys[kTcSample];
  __shared__ unsigned hist[2048];
  __shared__ unsigned s_pair[2];
  const float* row = ssims + (size_t)blockIdx.x * S;
  for (int i = threadIdx.x; i < S; i += blockDim.x) keys[i] = tc_ord_key(row[i]);
  __syncthreads();
  uint32_t T;
  unsigned need;
  block_kth_key(keys, S, m, hist, s_pair, T, need);
  if (threadIdx.x == 0) {
    tau[blockIdx.x] = tc_key_to_float(T);
    cnt[blockIdx.x] = 0u;
  }
}

__global__ void sim_tc_sample_rows_kernel(const float* __restrict__ hi, const float* __restrict__ lo, int n, int ws, int S,
                                          float* __restrict__ s_hi, float* __restrict__ s_lo) {
  const int j = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (j >= S) return;
  const size_t src = (size_t)(((long long)j * n) / S) * ws;
  for (int c = lane; c < ws; c += 32) {
    s_hi[(size_t)j * ws + c] = hi[src + c];
    s_lo[(size_t)j * ws + c] = lo[src + c];
  }
}

__global__ void sim_tc_gather_rows_kernel(const float* __restrict__ hi, const float* __restrict__ lo, const int32_t* __restrict__ rows,
                                          int row_base, int nf, int ws, float* __restrict__ g_hi, float* __restrict__ g_lo,
                                          const int32_t* __restrict__ out_rows, int32_t* __restrict__ mapped) {
  const int j = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (j >= nf) return;
  const int r = row_base + rows[j];
  for (int c = lane; c < ws; c += 32) {
    g_hi[(size_t)j * ws + c] = hi[(size_t)r * ws + c];
    g_lo[(size_t)j * ws + c] = lo[(size_t)r * ws + c];
  }
  if (lane == 0) mapped[j] = out_rows ? out_rows[r] : r;
}

struct SelectParams {
  const uint2* cand;
  const uint32_t* cnt;
  int cap, k, n, row_base;
  const int32_t* id_list;
  int id_base;
  int32_t* out;
  const int32_t* out_rows;
  int32_t* fb_list;     // rows (relative to row_base) that take the exact path
  uint32_t* fb_count;
};

// exclusive prefix sum over the block's threads
__device__ __forceinline__ unsigned block_excl_scan(unsigned v, unsigned* s_warp, unsigned& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  unsigned base = 0u, tot = 0u;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
    if (w < warp) base += s_warp[w];
    tot += s_warp[w];
  }
  total = tot;
  return base + inc - v;
}

__global__ void __launch_bounds__(kSelThreads) sim_tc_select_kernel(const SelectParams p) {
  extern __shared__ uint32_t s_sel[];
  uint32_t* keys = s_sel;                 // [cap]
  uint32_t* cols = keys + p.cap;          // [cap]
  const int words = (p.n + 31) >> 5;
  uint32_t* gt = cols + p.cap;            // [words]
  uint32_t* eq = gt + words;              // [words]
  __shared__ unsigned hist[2048];
  __shared__ unsigned s_pair[2];
  __shared__ unsigned s_warp[kSelThreads / 32];
  const int row = blockIdx.x, tid = threadIdx.x;
  const uint32_t c = p.cnt[row];
  if (c < (uint32_t)p.k || c > (uint32_t)p.cap) {  // short (threshold too high) or overflowed: the exact path redoes it
    if (tid == 0) p.fb_list[atomicAdd(p.fb_count, 1u)] = row;
    return;
  }
  const uint2* cand = p.cand + (size_t)row * p.cap;
  for (int i = tid; i < (int)c; i += blockDim.x) {
    const uint2 e = cand[i];
    cols[i] = e.x;
    keys[i] = tc_ord_key(__uint_as_float(e.y));
  }
  for (int w = tid; w < words; w += blockDim.x) {
    gt[w] = 0u;
    eq[w] = 0u;
  }
  __syncthreads();
  uint32_t T;
  unsigned need_eq;
  block_kth_key(keys, (int)c, (unsigned)p.k, hist, s_pair, T, need_eq);
  for (int i = tid; i < (int)c; i += blockDim.x) {
    const uint32_t key = keys[i], col = cols[i];
    if (key > T)
      atomicOr(&gt[col >> 5], 1u << (col & 31));
    else if (key == T)
      atomicOr(&eq[col >> 5], 1u << (col & 31));
  }
  __syncthreads();
  // ties at the k-th value: the first need_eq of them in ascending column order (the stable rule of the exact path)
  const int per = (words + blockDim.x - 1) / blockDim.x;
  const int w0 = tid * per < words ? tid * per : words, w1 = (tid + 1) * per < words ? (tid + 1) * per : words;
  unsigned mine = 0u, total;
  for (int w = w0; w < w1; ++w) mine += __popc(eq[w]);
  unsigned base = block_excl_scan(mine, s_warp, total);
  for (int w = w0; w < w1; ++w) {
    uint32_t word = eq[w];
    const unsigned pc = __popc(word);
    if (base >= need_eq) {
      word = 0u;
    } else if (base + pc > need_eq) {
      unsigned keep = need_eq - base;
      uint32_t kept = 0u;
      while (keep-- > 0u) {
        const uint32_t low = word & (0u - word);
        kept |= low;
        word ^= low;
      }
      word = kept;
    }
    base += pc;
    gt[w] |= word;
  }
  __syncthreads();
  mine = 0u;
  for (int w = w0; w < w1; ++w) mine += __popc(gt[w]);
  unsigned pos = block_excl_scan(mine, s_warp, total);
  const int orow = p.out_rows ? __ldg(p.out_rows + p.row_base + row) : p.row_base + row;
  int32_t* __restrict__ o = p.out + (size_t)orow * p.k;
  for (int w = w0; w < w1; ++w) {
    uint32_t word = gt[w];
    while (word != 0u) {
      const int col = (w << 5) + __ffs(word) - 1;
      o[pos++] = p.id_list ? __ldg(p.id_list + col) : p.id_base + col;
      word &= word - 1u;
    }
  }
}

// exact-path pieces of mke_sim.cu the fall-back uses
int sim_topk_rows_exact(const float* sims, size_t pitch, int n, int k, const int32_t* id_list, int id_base, int32_t* out,
                        const int32_t* out_rows, int rows, cudaStream_t stream);

// floats of workspace the fused search needs beyond the prepared rows: fixed part and per fused row
static void topk_fused_layout(int n, int ws, int k, size_t pitch, size_t& fixed, size_t& per_row, int& cap, unsigned& m) {
  cap = (int)(((long long)(2.2 * k) + 63) & ~63ll);
  m = (unsigned)((1.3 * (double)k * kTcSample) / (double)n) + 8u;
  fixed = (size_t)128 * pitch + 2 * (size_t)128 * ws + 2 * (size_t)kTcSample * ws + 256;
  per_row = (size_t)kTcSample + 2 * (size_t)cap + 4;
}

bool sim_topk_fused_applies(int n, int k) {
  if (!(n >= 4 * kTcSample && (long long)k * 8 <= n && n <= 262144 && k >= 64)) return false;
  // the select kernel keeps a row's candidates (2 x cap words) and two bitmaps over the columns in shared memory
  const long long cap = ((long long)(2.2 * k) + 63) & ~63ll;
  const long long words = (n + 31) >> 5;
  return (2 * cap + 2 * words) * 4 <= 200 * 1024;
}

// hi / lo: prepared rows [n, ws]; work: workspace after them (work_floats floats).  Returns 1 when the workspace is too
// small for the fused search (the caller runs the exact one).
int sim_topk_fused_tc(const float* hi, const float* lo, int n, int ws, int k, const int32_t* id_list, int id_base,
                      const int32_t* out_rows, float* work, int64_t work_floats, int32_t* out, cudaStream_t stream) {
  const size_t pitch = ((size_t)n + 3) & ~(size_t)3;
  size_t fixed, per_row;
  int cap;
  unsigned m;
  topk_fused_layout(n, ws, k, pitch, fixed, per_row, cap, m);
  if (work_floats < (int64_t)(fixed + 128 * per_row)) return 1;
  long long rf = (long long)((size_t)work_floats - fixed) / (long long)per_row;
  if (rf > n) rf = n;
  if (rf > kTcTile) rf -= rf % kTcTile;
  float* ex_sims = work;                                  // [128][pitch]   exact path of the fall-back rows
  float* g_hi = ex_sims + (size_t)128 * pitch;            // [128][ws] x 2  their prepared rows
  float* g_lo = g_hi + (size_t)128 * ws;
  float* s_hi = g_lo + (size_t)128 * ws;                  // [S][ws] x 2    the sampled columns
  float* s_lo = s_hi + (size_t)kTcSample * ws;
  float* chunk = s_lo + (size_t)kTcSample * ws + 64;
  float* ssims = chunk;                                   // [rf][S]
  float* tau = ssims + (size_t)rf * kTcSample;            // [rf]
  uint32_t* cnt = (uint32_t*)(tau + rf);                  // [rf]
  int32_t* fb_list = (int32_t*)(cnt + rf);                // [rf] + count + mapped rows
  uint32_t* fb_count = (uint32_t*)(fb_list + rf);
  uint2* cand = (uint2*)(((uintptr_t)(fb_count + 2) + 15) & ~(uintptr_t)15);  // [rf][cap]
  sim_tc_sample_rows_kernel<<<(kTcSample + 7) / 8, 256, 0, stream>>>(hi, lo, n, ws, kTcSample, s_hi, s_lo);
  MKE_CHECK_LAUNCH("sim_tc_sample_rows_kernel");
  const int words = (n + 31) >> 5;
  const size_t sel_smem = ((size_t)2 * cap + 2 * (size_t)words) * sizeof(uint32_t);
  static size_t sel_configured = 0;
  if (sel_configured < sel_smem) {
    if (cudaError_t e = cudaFuncSetAttribute(sim_tc_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sel_smem))
      return cuda_fail(e, "cudaFuncSetAttribute(sim_tc_select_kernel)");
    sel_configured = sel_smem;
  }
  for (int r0 = 0; r0 < n; r0 += (int)rf) {
    const int rows = r0 + rf < n ? (int)rf : n - r0;
    SimTcParams ps{};  // (1) sims with the sampled columns
    ps.n1 = n;
    ps.n2 = kTcSample;
    ps.row_base = r0;
    ps.rows = rows;
    ps.out = ssims;
    ps.out_pitch = kTcSample;
    if (int rc = launch_sim_tc<kTcStore>(hi, lo, s_hi, s_lo, ws, ps, stream)) return rc;
    sim_tc_threshold_kernel<<<rows, kSelThreads, 0, stream>>>(ssims, kTcSample, m, tau, cnt);
    MKE_CHECK_LAUNCH("sim_tc_threshold_kernel");
    if (cudaError_t e = cudaMemsetAsync(fb_count, 0, 2 * sizeof(uint32_t), stream)) return cuda_fail(e, "memset");
    SimTcParams pf{};  // (2) all columns, candidates only
    pf.n1 = n;
    pf.n2 = n;
    pf.row_base = r0;
    pf.rows = rows;
    pf.tau = tau;
    pf.cand_cnt = cnt;
    pf.cand = cand;
    pf.cap = cap;
    if (int rc = launch_sim_tc<kTcFilter>(hi, lo, hi, lo, ws, pf, stream)) return rc;
    SelectParams sp{cand, cnt, cap, k, n, r0, id_list, id_base, out, out_rows, fb_list, fb_count};  // (3)
    sim_tc_select_kernel<<<rows, kSelThreads, sel_smem, stream>>>(sp);
    MKE_CHECK_LAUNCH("sim_tc_select_kernel");
    uint32_t nf = 0;  // (4) rows for the exact path (expected: a few in 10 000)
    if (cudaError_t e = cudaMemcpyAsync(&nf, fb_count, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream)) return cuda_fail(e, "D2H");
    if (cudaError_t e = cudaStreamSynchronize(stream)) return cuda_fail(e, "sync");
    for (uint32_t f0 = 0; f0 < nf; f0 += 128) {
      const int cnt_f = nf - f0 < 128 ? (int)(nf - f0) : 128;
      // the mapped output rows of this batch live in the first 128 words of the (now dead) sample-sims area
      int32_t* mapped = (int32_t*)ssims;
      sim_tc_gather_rows_kernel<<<(cnt_f + 7) / 8, 256, 0, stream>>>(hi, lo, fb_list + f0, r0, cnt_f, ws, g_hi, g_lo, out_rows, mapped);
      MKE_CHECK_LAUNCH("sim_tc_gather_rows_kernel");
      SimTcParams pe{};
      pe.n1 = cnt_f;
      pe.n2 = n;
      pe.row_base = 0;
      pe.rows = cnt_f;
      pe.out = ex_sims;
      pe.out_pitch = pitch;
      if (int rc = launch_sim_tc<kTcStore>(g_hi, g_lo, hi, lo, ws, pe, stream)) return rc;
      if (int rc = sim_topk_rows_exact(ex_sims, pitch, n, k, id_list, id_base, out, mapped, cnt_f, stream)) return rc;
    }
  }
  return 0;
}

int sim_tc_ws(int dim) { return (dim + 7) & ~7; }

// the evaluator on the tensor cores; workspace layout: a_hi, a_lo [n1, ws], b_hi, b_lo [n2, ws], g_hi, g_lo [n1, ws],
// gold scores [n1 (+1)], arg-max keys [n1] (8 bytes each)
int sim_rank_tc(const float* emb1, const int32_t* idx1, int n1, const float* emb2, const int32_t* idx2, int n2, int stride,
                int dim, int normalize, const int32_t* gold, float* workspace, int32_t* rank_out, int32_t* top1_out,
                unsigned long long** best_out, cudaStream_t stream) {
  const int ws = sim_tc_ws(dim);
  float* a_hi = workspace;
  float* a_lo = a_hi + (size_t)n1 * ws;
  float* b_hi = a_lo + (size_t)n1 * ws;
  float* b_lo = b_hi + (size_t)n2 * ws;
  float* g_hi = b_lo + (size_t)n2 * ws;
  float* g_lo = g_hi + (size_t)n1 * ws;
  float* gold_score = g_lo + (size_t)n1 * ws;
  unsigned long long* best = reinterpret_cast<unsigned long long*>(gold_score + ((n1 + 1) & ~1));
  const int wpb = 8;
  sim_tc_prepare_kernel<<<(n1 + wpb - 1) / wpb, wpb * 32, 0, stream>>>(emb1, idx1, n1, stride, dim, normalize, a_hi, a_lo, ws);
  MKE_CHECK_LAUNCH("sim_tc_prepare_kernel");
  sim_tc_prepare_kernel<<<(n2 + wpb - 1) / wpb, wpb * 32, 0, stream>>>(emb2, idx2, n2, stride, dim, normalize, b_hi, b_lo, ws);
  MKE_CHECK_LAUNCH("sim_tc_prepare_kernel");
  sim_tc_gather_gold_kernel<<<(n1 + wpb - 1) / wpb, wpb * 32, 0, stream>>>(b_hi, b_lo, gold, n1, n2, ws, g_hi, g_lo, rank_out, best);
  MKE_CHECK_LAUNCH("sim_tc_gather_gold_kernel");
  SimTcParams p{};
  p.n1 = n1;
  p.row_base = 0;
  p.rows = n1;
  p.gold = gold;
  p.gold_score = gold_score;
  p.rank = rank_out;
  p.best = best;
  SimTcParams pg = p;
  pg.n2 = n1;  // the gathered rows: one per row of A
  pg.gold_n2 = n2;
  if (int rc = launch_sim_tc<kTcGold>(a_hi, a_lo, g_hi, g_lo, ws, pg, stream)) return rc;
  p.n2 = n2;
  if (int rc = launch_sim_tc<kTcRank>(a_hi, a_lo, b_hi, b_lo, ws, p, stream)) return rc;
  *best_out = best;
  return 0;
}

// rows [row_base, row_base + rows) of sims of the prepared matrix with itself -> out [rows, pitch]
int sim_store_tc(const float* hi, const float* lo, int n, int ws, int row_base, int rows, float* out, size_t pitch,
                 cudaStream_t stream) {
  SimTcParams p{};
  p.n1 = n;
  p.n2 = n;
  p.row_base = row_base;
  p.rows = rows;
  p.out = out;
  p.out_pitch = pitch;
  return launch_sim_tc<kTcStore>(hi, lo, hi, lo, ws, p, stream);
}

int sim_prepare_tc(const float* src, const int32_t* idx, int n, int stride, int dim, int normalize, float* hi, float* lo,
                   cudaStream_t stream) {
  const int wpb = 8;
  sim_tc_prepare_kernel<<<(n + wpb - 1) / wpb, wpb * 32, 0, stream>>>(src, idx, n, stride, dim, normalize, hi, lo, sim_tc_ws(dim));
  MKE_CHECK_LAUNCH("sim_tc_prepare_kernel");
  return 0;
}

}  // namespace mke
