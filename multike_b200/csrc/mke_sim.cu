// Inner-product similarity search on dense fp32 rows:
//   mke_sim_rank  -- the Hits@k / MR / MRR evaluator: rank of the gold counterpart of every row
//                    and its arg-max column (base/similarity.py:9-52 sim(metric='inner',
//                    normalize=True), base/alignment.py:8-79 greedy_alignment, :141-163
//                    calculate_rank, called from MultiKE_Late.py:14-61 through base/evaluation.py);
//                    the [n1, n2] similarity matrix (14 GB at test size) is never materialised.
//   mke_sim_topk  -- truncated-epsilon neighbour lists (base/batch.py:119-150
//                    generate_neighbours / find_neighbours, called from MultiKE_CSL.py:89-99):
//                    the k most similar rows of every row, straight into the table the on-device
//                    sampler reads (mke_kg_sampler_t.neighbours).
//
// Two implementations of the contraction.  DEFAULT: tcgen05 tiles at fp32-equivalent precision (3xTF32) with the
// consumers fused into the TMEM epilogue -- mke_sim_tc.cu.  BASELINE (mke_sim_use_tensor_cores(0) / MKE_SIM_TC=0),
// the kernels of this file: fp32 FMA pipe, one thread
// block owns a 128-row tile of A, streams 128-row tiles of B through a double-buffered cp.async
// stage and keeps an 8 x 4 block of sims per thread in registers.  Every sim is the sum of two fmaf
// chains in ascending k (even and odd columns, packed FFMA2) -- the same two chains
// sim_gold_kernel uses -- so equal rows give bit-equal sims and the tie rules below are exact.
#include <cstdlib>
#include "mke_common.cuh"

namespace mke {

constexpr int kSimTile = 128;     // rows of A / of B per tile
constexpr int kSimThreads = 512;  // 16 warps x 32 lanes, 8 x 4 sims each: 4 warps per scheduler hide the LDS latency
constexpr int kSimMaxWs = 128;    // workspace row stride supported by the shared-memory stage

// order-preserving map float -> uint32 (larger float => larger key; -0 < +0)
__device__ __forceinline__ uint32_t ord_key(float f) {
  const uint32_t u = __float_as_uint(f);
  return u ^ ((u >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}

// One warp per row: gather by idx, optionally x / ||x|| (sklearn.preprocessing.normalize as used by
// base/similarity.py:31-33: a zero row stays zero), zero the pad columns.
__global__ void sim_prepare_kernel(const float* __restrict__ src, const int32_t* __restrict__ idx, int n,
                                   int stride, int dim, int normalize, float* __restrict__ dst, int ws) {
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
  float* d = dst + (size_t)row * ws;
  for (int c = lane; c < ws; c += 32) d[c] = c < dim ? __ldg(s + c) / nrm : 0.f;
}

// sim of every row with its gold column, by the same fmaf chain as the tile kernel; resets the
// per-row outputs the tile kernel accumulates into.
__global__ void sim_gold_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                const int32_t* __restrict__ gold, int n1, int n2, int ws,
                                float* __restrict__ gold_score, int32_t* __restrict__ rank,
                                unsigned long long* __restrict__ best) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n1) return;
  const int g = gold ? __ldg(gold + i) : i;
  float acc = __int_as_float(0x7f800000);  // gold outside [0, n2): nothing ranks before it
  if (g >= 0 && g < n2) {
    const float* x = a + (size_t)i * ws;
    const float* y = b + (size_t)g * ws;
    float even = 0.f, odd = 0.f;  // the two chains of the tile kernel's packed accumulators
    for (int k = 0; k < ws; k += 2) {
      even = fmaf(__ldg(x + k), __ldg(y + k), even);
      odd = fmaf(__ldg(x + k + 1), __ldg(y + k + 1), odd);
    }
    acc = even + odd;
  }
  gold_score[i] = acc;
  rank[i] = 0;
  best[i] = 0ull;
}

__global__ void sim_finish_kernel(const unsigned long long* __restrict__ best, int n1, int32_t* __restrict__ top1) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n1) top1[i] = (int32_t)(0xFFFFFFFFu - (uint32_t)(best[i] & 0xFFFFFFFFull));
}

struct SimParams {
  const float* a;  // [n1, ws] prepared rows
  int n1;
  const float* b;  // [n2, ws]
  int n2;
  int ws;          // floats per prepared row (multiple of 8, <= kSimMaxWs)
  int splits;      // column splits: block (rb, cs) owns B tiles cs, cs + splits, ...
  int row_base;    // first row of A this launch covers
  int rows;        // rows of A this launch covers
  // rank mode
  const int32_t* gold;
  const float* gold_score;
  int32_t* rank;
  unsigned long long* best;
  // store mode
  float* out;  // [rows, out_pitch]
  size_t out_pitch;
};

__device__ __forceinline__ void sim_cp16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;  // src-size 0 => the 16 bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}

// rows [row0, row0 + 128) of m (n rows in all) -> tile [128][pitch]
__device__ __forceinline__ void sim_load_tile(uint32_t tile, const float* __restrict__ m, int n, int row0, int ws,
                                              int pitch) {
  const int pieces = ws >> 2;
  for (int q = threadIdx.x; q < kSimTile * pieces; q += kSimThreads) {
    const int r = q / pieces, c = q - r * pieces;
    const bool valid = row0 + r < n;
    sim_cp16(tile + (uint32_t)(r * pitch + 4 * c) * 4u, m + (size_t)(valid ? row0 + r : 0) * ws + 4 * c, valid);
  }
}

template <bool RANK>
__global__ void __launch_bounds__(kSimThreads, 1) sim_tile_kernel(const SimParams p) {
  extern __shared__ __align__(16) float s_sim[];
  __shared__ float s_sg[kSimTile];
  __shared__ int s_gi[kSimTile];
  const int pitch = p.ws + 4;  // (pitch / 4) odd: the 8 lanes of a 128-bit phase hit 8 bank groups
  float* const As = s_sim;
  float* const Bs0 = s_sim + kSimTile * pitch;
  const uint32_t a_addr = (uint32_t)__cvta_generic_to_shared(As);
  const uint32_t b_addr = (uint32_t)__cvta_generic_to_shared(Bs0);
  const uint32_t tile_bytes = (uint32_t)(kSimTile * pitch * 4);
  // Thread layout: warp w owns the 8 rows [8 w, 8 w + 8) of the A tile, lane l the 4 columns
  // {l + 32 j} of the B tile.  Every A read is one address per warp (a broadcast: one shared-memory
  // wavefront instead of four), every B read is conflict free, and a row's 32 sims of a store are
  // one 128-byte segment.  Per 4 embedding columns a thread issues 12 LDS.128 for 64 FFMA2.
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int RI = kSimTile / (kSimThreads / 32), CJ = 4;
  const int rb = blockIdx.x / p.splits, cs = blockIdx.x - rb * p.splits;
  const int row0 = p.row_base + rb * kSimTile;
  const int row_end = p.row_base + p.rows < p.n1 ? p.row_base + p.rows : p.n1;
  const int ntiles = (p.n2 + kSimTile - 1) / kSimTile;

  sim_load_tile(a_addr, p.a, row_end, row0, p.ws, pitch);
  if (cs < ntiles) sim_load_tile(b_addr, p.b, p.n2, cs * kSimTile, p.ws, pitch);
  asm volatile("cp.async.commit_group;" ::: "memory");

  // rank mode: gold score / column of the tile's rows (shared), per thread the count and the running
  // arg-max of its 16 rows over its columns
  if (RANK && threadIdx.x < kSimTile) {
    const int r = row0 + threadIdx.x;
    const bool ok = r < row_end;
    s_sg[threadIdx.x] = ok ? __ldg(p.gold_score + r) : __int_as_float(0x7f800000);
    s_gi[threadIdx.x] = ok ? (p.gold ? __ldg(p.gold + r) : r) : -1;
  }
  int cnt[RI], bi[RI];
  float bs[RI];
#pragma unroll
  for (int i = 0; i < RI; ++i) {
    cnt[i] = 0;
    bi[i] = 0x7fffffff;
    bs[i] = __int_as_float(0xff800000);
  }

  int buf = 0;
  for (int jt = cs; jt < ntiles; jt += p.splits) {
    const int nxt = jt + p.splits;
    if (nxt < ntiles) sim_load_tile(b_addr + (uint32_t)(buf ^ 1) * tile_bytes, p.b, p.n2, nxt * kSimTile, p.ws, pitch);
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    __syncthreads();
    const float* At = As + (warp * RI) * pitch;
    const float* Bt = Bs0 + (size_t)buf * kSimTile * pitch + lane * pitch;
    // Packed fp32 (fma.rn.f32x2, SASS FFMA2): every sim is accumulated as two chains -- the even
    // and the odd columns of the embedding -- that ride in one 64-bit register pair and are added
    // at the end; one issue slot feeds two FMAs.
    float2 acc2[RI][CJ];
#pragma unroll
    for (int i = 0; i < RI; ++i)
#pragma unroll
      for (int j = 0; j < CJ; ++j) acc2[i][j] = make_float2(0.f, 0.f);
#pragma unroll 1
    for (int k4 = 0; k4 < (p.ws >> 2); ++k4) {
      float2 b01[CJ], b23[CJ];
#pragma unroll
      for (int j = 0; j < CJ; ++j) {
        const float4 b = *reinterpret_cast<const float4*>(Bt + 32 * j * pitch + 4 * k4);
        b01[j] = make_float2(b.x, b.y);
        b23[j] = make_float2(b.z, b.w);
      }
#pragma unroll
      for (int i = 0; i < RI; ++i) {
        const float4 a = *reinterpret_cast<const float4*>(At + i * pitch + 4 * k4);
        const float2 a01 = make_float2(a.x, a.y), a23 = make_float2(a.z, a.w);
#pragma unroll
        for (int j = 0; j < CJ; ++j) {
          float2 v = acc2[i][j];
          v = __ffma2_rn(a01, b01[j], v);
          v = __ffma2_rn(a23, b23[j], v);
          acc2[i][j] = v;
        }
      }
    }
    // ---- epilogue of this tile ----------------------------------------------------------------
    const int col0 = jt * kSimTile + lane;
    if constexpr (RANK) {
#pragma unroll
      for (int i = 0; i < RI; ++i) {
        const float sg = s_sg[warp * RI + i];
        const int gi = s_gi[warp * RI + i];
#pragma unroll
        for (int j = 0; j < CJ; ++j) {
          const int col = col0 + 32 * j;
          const bool cv = col < p.n2;
          const float s = acc2[i][j].x + acc2[i][j].y;
          // stable descending order (base/alignment.py:148 argsort of -sim): a column ranks before
          // the gold one if its sim is larger, or equal with a smaller column index
          const bool before = cv && col != gi && (s > sg || (s == sg && col < gi));
          cnt[i] += before ? 1 : 0;
          if (cv && s > bs[i]) {  // this thread sees its columns in ascending order
            bs[i] = s;
            bi[i] = col;
          }
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < RI; ++i) {
        const int r = row0 + warp * RI + i;
        if (r < row_end) {
          float* o = p.out + (size_t)(r - p.row_base) * p.out_pitch;
#pragma unroll
          for (int j = 0; j < CJ; ++j) {
            const int col = col0 + 32 * j;
            if (col < p.n2) o[col] = acc2[i][j].x + acc2[i][j].y;
          }
        }
      }
    }
    __syncthreads();  // everyone is done with `buf` before the next prefetch overwrites it
    buf ^= 1;
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  if constexpr (RANK) {
    // the 32 lanes of a warp combine their columns, then one atomic per row and block
#pragma unroll
    for (int i = 0; i < RI; ++i) {
      int c = cnt[i];
      unsigned long long key =
          bi[i] == 0x7fffffff ? 0ull : (((unsigned long long)ord_key(bs[i]) << 32) | (0xFFFFFFFFu - (uint32_t)bi[i]));
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        c += __shfl_xor_sync(0xffffffffu, c, o);
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
        key = other > key ? other : key;
      }
      const int r = row0 + warp * RI + i;
      if (lane == 0 && r < row_end) {
        if (c != 0) atomicAdd(p.rank + r, c);
        atomicMax(p.best + r, key);
      }
    }
  }
}

// ---- exact top-k of one row of sims per thread block -------------------------------------------
// Radix select on the order-preserving keys (11 + 11 + 10 bits, warp-aggregated shared-memory
// histogram) gives the key T of the k-th largest sim and how many columns equal to T belong to
// the answer; the columns are then emitted in ASCENDING column order (all with key > T plus the
// first `need_eq` with key == T), i.e. the list is deterministic where np.argpartition's is not.
constexpr int kTopkThreads = 256;
constexpr int kTopkWarps = kTopkThreads / 32;

__global__ void __launch_bounds__(kTopkThreads) row_topk_kernel(const float* __restrict__ sims, size_t pitch, int n2,
                                                                int k, const int32_t* __restrict__ id_list,
                                                                int id_base, int32_t* __restrict__ out,
                                                                const int32_t* __restrict__ out_rows, int row_base) {
  __shared__ unsigned hist[2048];
  __shared__ unsigned s_bin, s_above;
  __shared__ int s_wg[kTopkWarps], s_we[kTopkWarps];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* __restrict__ s = sims + (size_t)blockIdx.x * pitch;
  uint32_t prefix = 0u, mask = 0u;
  unsigned remaining = (unsigned)k;
  for (int pass = 0; pass < 3; ++pass) {
    const int shift = pass == 0 ? 21 : (pass == 1 ? 10 : 0);
    const int nb = pass < 2 ? 2048 : 1024;
    for (int b = tid; b < nb; b += kTopkThreads) hist[b] = 0u;
    __syncthreads();
    // four consecutive columns per thread and trip (rows are 16-byte aligned and padded to a
    // multiple of four floats; the pad is masked).  Pass 0 sees every column and only a handful
    // of distinct bins (the top bits of sims in [-1, 1]): one shared-memory atomic per warp and
    // distinct bin (match.any).  Later passes see only the few columns of the threshold bin.
    for (int c0 = 0; c0 < n2; c0 += 4 * kTopkThreads) {
      const int c = c0 + 4 * tid;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < n2) v = *reinterpret_cast<const float4*>(s + c);
      const float ve[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const uint32_t key = ord_key(ve[e]);
        const bool ok = c + e < n2 && (key & mask) == prefix;
        const unsigned bin = (key >> shift) & (unsigned)(nb - 1);
        if (pass == 0) {
          const unsigned peers = __match_any_sync(0xffffffffu, ok ? bin : 0xffffffffu);
          if (ok && lane == __ffs(peers) - 1) atomicAdd(&hist[bin], (unsigned)__popc(peers));
        } else if (ok) {
          atomicAdd(&hist[bin], 1u);
        }
      }
    }
    __syncthreads();
    if (warp == 0) {
      // lane l owns bins [l * per, (l + 1) * per); suffix sums from the top bin down
      const int per = nb / 32;
      unsigned mine = 0u;
      for (int b = 0; b < per; ++b) mine += hist[lane * per + b];
      unsigned above = 0u;  // count in the bins of the lanes above this one
      for (int l = 31; l > 0; --l) {
        const unsigned v = __shfl_sync(0xffffffffu, mine, l);
        if (l > lane) above += v;
      }
      if (above < remaining && remaining <= above + mine) {
        unsigned acc = above;
        for (int b = per - 1; b >= 0; --b) {
          const unsigned h = hist[lane * per + b];
          if (acc + h >= remaining) {
            s_bin = (unsigned)(lane * per + b);
            s_above = acc;
            break;
          }
          acc += h;
        }
      }
    }
    __syncthreads();
    prefix |= s_bin << shift;
    mask |= (unsigned)(nb - 1) << shift;
    remaining -= s_above;
    __syncthreads();
  }
  const uint32_t T = prefix;
  const int need_eq = (int)remaining;
  // ---- emit in column order: warp w owns the columns [w * span, (w + 1) * span) ----------------
  const int span = (((n2 + kTopkWarps - 1) / kTopkWarps) + 31) & ~31;
  const int c_begin = warp * span, c_end = (c_begin + span < n2) ? c_begin + span : n2;
  int wg = 0, we = 0;
  for (int c0 = c_begin; c0 < c_end; c0 += 32) {
    const int c = c0 + lane;
    const uint32_t key = c < c_end ? ord_key(s[c]) : 0u;
    wg += __popc(__ballot_sync(0xffffffffu, c < c_end && key > T));
    we += __popc(__ballot_sync(0xffffffffu, c < c_end && key == T));
  }
  if (lane == 0) {
    s_wg[warp] = wg;
    s_we[warp] = we;
  }
  __syncthreads();
  int eq_base = 0, pos = 0;
  for (int v = 0; v < warp; ++v) {
    int take = need_eq - eq_base;
    take = take < 0 ? 0 : (take > s_we[v] ? s_we[v] : take);
    pos += s_wg[v] + take;
    eq_base += s_we[v];
  }
  const int orow = out_rows ? __ldg(out_rows + row_base + blockIdx.x) : row_base + (int)blockIdx.x;
  int32_t* __restrict__ o = out + (size_t)orow * k;
  const unsigned lt = (1u << lane) - 1u;
  for (int c0 = c_begin; c0 < c_end; c0 += 32) {
    const int c = c0 + lane;
    const uint32_t key = c < c_end ? ord_key(s[c]) : 0u;
    const bool gt = c < c_end && key > T, eq = c < c_end && key == T;
    const unsigned be = __ballot_sync(0xffffffffu, eq);
    const bool take = gt || (eq && eq_base + __popc(be & lt) < need_eq);
    const unsigned bt = __ballot_sync(0xffffffffu, take);
    if (take) o[pos + __popc(bt & lt)] = id_list ? __ldg(id_list + c) : id_base + c;
    pos += __popc(bt);
    eq_base += __popc(be);
  }
}

static int sim_splits(int row_blocks, int ntiles) {
  // enough blocks to fill the SMs several times over (the tail of the last wave shrinks with the
  // block size), but at least 4 B tiles per block so that the A tile load stays amortised
  int s = (8 * sm_count() + row_blocks - 1) / row_blocks;
  const int cap = ntiles / 4 > 1 ? ntiles / 4 : 1;
  s = s > cap ? cap : s;
  return s < 1 ? 1 : s;
}

static int sim_ws(int dim) { return (dim + 7) & ~7; }

template <bool RANK>
static int launch_sim_tiles(const SimParams& p, cudaStream_t stream) {
  auto kern = sim_tile_kernel<RANK>;
  const size_t smem = (size_t)3 * kSimTile * (p.ws + 4) * sizeof(float);
  static size_t configured[2] = {0, 0};
  if (configured[RANK] < smem) {
    if (cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem))
      return cuda_fail(e, "cudaFuncSetAttribute(sim_tile_kernel)");
    configured[RANK] = smem;
  }
  const int row_blocks = (p.rows + kSimTile - 1) / kSimTile;
  kern<<<row_blocks * p.splits, kSimThreads, smem, stream>>>(p);
  MKE_CHECK_LAUNCH("sim_tile_kernel");
  return 0;
}

static int sim_prepare(const float* src, const int32_t* idx, int n, int stride, int dim, int normalize, float* dst,
                       int ws, cudaStream_t stream) {
  const int warps_per_block = 8;
  sim_prepare_kernel<<<(n + warps_per_block - 1) / warps_per_block, warps_per_block * 32, 0, stream>>>(
      src, idx, n, stride, dim, normalize, dst, ws);
  MKE_CHECK_LAUNCH("sim_prepare_kernel");
  return 0;
}

// mke_sim_tc.cu: the same searches on the tensor cores (default; MKE_SIM_TC=0 selects the fp32 FMA tiles above, kept
// as the measured baseline)
int sim_tc_ws(int dim);
int sim_rank_tc(const float* emb1, const int32_t* idx1, int n1, const float* emb2, const int32_t* idx2, int n2, int stride,
                int dim, int normalize, const int32_t* gold, float* workspace, int32_t* rank_out, int32_t* top1_out,
                unsigned long long** best_out, cudaStream_t stream);
int sim_store_tc(const float* hi, const float* lo, int n, int ws, int row_base, int rows, float* out, size_t pitch,
                 cudaStream_t stream);
int sim_prepare_tc(const float* src, const int32_t* idx, int n, int stride, int dim, int normalize, float* hi, float* lo,
                   cudaStream_t stream);
static int g_sim_tc = -1;  // -1: not decided yet (MKE_SIM_TC, default on)
int sim_topk_fused_tc(const float* hi, const float* lo, int n, int ws, int k, const int32_t* id_list, int id_base,
                      const int32_t* out_rows, float* work, int64_t work_floats, int32_t* out, cudaStream_t stream);
bool sim_topk_fused_applies(int n, int k);
// exact top-k of `rows` rows of sims (row r -> output row out_rows[r]); used by the fall-back of the fused search
int sim_topk_rows_exact(const float* sims, size_t pitch, int n, int k, const int32_t* id_list, int id_base, int32_t* out,
                        const int32_t* out_rows, int rows, cudaStream_t stream) {
  row_topk_kernel<<<rows, kTopkThreads, 0, stream>>>(sims, pitch, n, k, id_list, id_base, out, out_rows, 0);
  MKE_CHECK_LAUNCH("row_topk_kernel");
  return 0;
}
static int g_sim_fused = -1;  // MKE_SIM_FUSED=0: always materialise the rows of sims (the exact path)
static bool sim_use_tc() {
  if (g_sim_tc < 0) g_sim_tc = getenv("MKE_SIM_TC") ? (atoi(getenv("MKE_SIM_TC")) != 0) : 1;
  return g_sim_tc != 0;
}

}  // namespace mke

using namespace mke;

extern "C" int mke_sim_use_tensor_cores(int32_t on) {
  if (g_sim_fused < 0) g_sim_fused = getenv("MKE_SIM_FUSED") ? (atoi(getenv("MKE_SIM_FUSED")) != 0) : 1;
  const int prev = sim_use_tc() ? (g_sim_fused ? 1 : 2) : 0;
  if (on >= 0) {
    g_sim_tc = on != 0;
    g_sim_fused = on != 2;
  }
  return prev;
}

extern "C" int64_t mke_sim_rank_workspace_floats(int32_t n1, int32_t n2, int32_t dim) {
  if (n1 < 0 || n2 < 0 || dim <= 0) return -1;
  const int64_t ws = sim_ws(dim);
  // prepared A and B as TF32 part + remainder, the gathered gold rows of B likewise, gold scores, packed (score,
  // column) arg-max keys (8 bytes each)
  return 4 * (int64_t)n1 * ws + 2 * (int64_t)n2 * ws + ((n1 + 1) & ~1) + 2 * (int64_t)n1 + 8;
}

extern "C" int mke_sim_rank(const float* emb1, const int32_t* idx1_or_null, int32_t n1, const float* emb2,
                            const int32_t* idx2_or_null, int32_t n2, int32_t stride, int32_t dim,
                            int32_t normalize, const int32_t* gold_or_null, float* workspace,
                            int32_t* rank_out, int32_t* top1_out, mke_stream_t stream_) {
  MKE_CHECK_ARG(n1 >= 0 && n2 >= 0, "negative row count");
  if (n1 == 0) return 0;
  MKE_CHECK_ARG(emb1 && emb2 && workspace && rank_out && top1_out, "null pointer");
  MKE_CHECK_ARG(n2 > 0, "no candidate rows");
  MKE_CHECK_ARG(dim > 0 && dim <= stride && sim_ws(dim) <= kSimMaxWs, "dim=%d outside (0,%d] or > stride=%d", dim,
                kSimMaxWs, stride);
  MKE_CHECK_ARG(((uintptr_t)workspace & 15) == 0, "workspace must be 16-byte aligned");
  cudaStream_t stream = (cudaStream_t)stream_;
  const int ws = sim_ws(dim);
  if (sim_use_tc()) {
    unsigned long long* best_tc = nullptr;
    if (int rc = sim_rank_tc(emb1, idx1_or_null, n1, emb2, idx2_or_null, n2, stride, dim, normalize, gold_or_null, workspace,
                             rank_out, top1_out, &best_tc, stream))
      return rc;
    sim_finish_kernel<<<(n1 + 255) / 256, 256, 0, stream>>>(best_tc, n1, top1_out);
    MKE_CHECK_LAUNCH("sim_finish_kernel");
    return 0;
  }
  float* a = workspace;
  float* b = a + (size_t)n1 * ws;
  float* gold_score = b + (size_t)n2 * ws;
  unsigned long long* best = reinterpret_cast<unsigned long long*>(gold_score + ((n1 + 1) & ~1));
  if (int rc = sim_prepare(emb1, idx1_or_null, n1, stride, dim, normalize, a, ws, stream)) return rc;
  if (int rc = sim_prepare(emb2, idx2_or_null, n2, stride, dim, normalize, b, ws, stream)) return rc;
  sim_gold_kernel<<<(n1 + 127) / 128, 128, 0, stream>>>(a, b, gold_or_null, n1, n2, ws, gold_score, rank_out, best);
  MKE_CHECK_LAUNCH("sim_gold_kernel");
  SimParams p{};
  p.a = a;
  p.n1 = n1;
  p.b = b;
  p.n2 = n2;
  p.ws = ws;
  p.row_base = 0;
  p.rows = n1;
  p.splits = sim_splits((n1 + kSimTile - 1) / kSimTile, (n2 + kSimTile - 1) / kSimTile);
  p.gold = gold_or_null;
  p.gold_score = gold_score;
  p.rank = rank_out;
  p.best = best;
  if (int rc = launch_sim_tiles<true>(p, stream)) return rc;
  sim_finish_kernel<<<(n1 + 255) / 256, 256, 0, stream>>>(best, n1, top1_out);
  MKE_CHECK_LAUNCH("sim_finish_kernel");
  return 0;
}

extern "C" int64_t mke_sim_topk_workspace_floats(int32_t n, int32_t dim, int32_t chunk_rows) {
  if (n < 0 || dim <= 0 || chunk_rows <= 0) return -1;
  const int64_t pitch = ((int64_t)n + 3) & ~3ll;
  const int64_t rows = chunk_rows < n ? chunk_rows : n;
  return 2 * (int64_t)n * sim_ws(dim) + rows * pitch + 8;  // prepared rows (TF32 part + remainder), `rows` rows of sims
}

extern "C" int mke_sim_topk(const float* emb, const int32_t* idx_or_null, int32_t n, int32_t stride, int32_t dim,
                            int32_t normalize, int32_t k, const int32_t* id_list_or_null, int32_t id_base,
                            const int32_t* out_rows_or_null, float* workspace, int64_t workspace_floats,
                            int32_t* neighbours_out, mke_stream_t stream_) {
  MKE_CHECK_ARG(n >= 0, "negative row count");
  if (n == 0) return 0;
  MKE_CHECK_ARG(emb && workspace && neighbours_out, "null pointer");
  MKE_CHECK_ARG(k >= 1 && k <= n, "k=%d outside [1, n=%d]", k, n);
  MKE_CHECK_ARG(dim > 0 && dim <= stride && sim_ws(dim) <= kSimMaxWs, "dim=%d outside (0,%d] or > stride=%d", dim,
                kSimMaxWs, stride);
  MKE_CHECK_ARG(((uintptr_t)workspace & 15) == 0, "workspace must be 16-byte aligned");
  cudaStream_t stream = (cudaStream_t)stream_;
  const int ws = sim_ws(dim);
  const size_t pitch = ((size_t)n + 3) & ~(size_t)3;
  const bool tc = sim_use_tc();
  const int64_t left = workspace_floats - (tc ? 2 : 1) * (int64_t)n * ws;
  MKE_CHECK_ARG(left >= (int64_t)pitch, "workspace of %lld floats holds no row of sims (see mke_sim_topk_workspace_floats)",
                (long long)workspace_floats);
  int chunk = (int)((left / (int64_t)pitch) < n ? (left / (int64_t)pitch) : n);
  if (chunk > kSimTile) chunk -= chunk % kSimTile;  // whole tiles of A per launch
  float* a = workspace;
  float* sims = a + (tc ? 2 : 1) * (size_t)n * ws;
  if (tc) {
    if (int rc = sim_prepare_tc(emb, idx_or_null, n, stride, dim, normalize, a, a + (size_t)n * ws, stream)) return rc;
    if (g_sim_fused < 0) g_sim_fused = getenv("MKE_SIM_FUSED") ? (atoi(getenv("MKE_SIM_FUSED")) != 0) : 1;
    if (g_sim_fused && sim_topk_fused_applies(n, k)) {
      const int rc = sim_topk_fused_tc(a, a + (size_t)n * ws, n, ws, k, id_list_or_null, id_base, out_rows_or_null, sims,
                                       workspace_floats - 2 * (int64_t)n * ws, neighbours_out, stream);
      if (rc <= 0) return rc;  // 1: the workspace is too small for it
    }
  } else {
    if (int rc = sim_prepare(emb, idx_or_null, n, stride, dim, normalize, a, ws, stream)) return rc;
  }
  for (int r0 = 0; r0 < n; r0 += chunk) {
    const int rows = r0 + chunk < n ? chunk : n - r0;
    if (tc) {
      if (int rc = sim_store_tc(a, a + (size_t)n * ws, n, ws, r0, rows, sims, pitch, stream)) return rc;
      row_topk_kernel<<<rows, kTopkThreads, 0, stream>>>(sims, pitch, n, k, id_list_or_null, id_base, neighbours_out,
                                                          out_rows_or_null, r0);
      MKE_CHECK_LAUNCH("row_topk_kernel");
      continue;
    }
    SimParams p{};
    p.a = a;
    p.n1 = n;
    p.b = a;
    p.n2 = n;
    p.ws = ws;
    p.row_base = r0;
    p.rows = rows;
    p.splits = sim_splits((rows + kSimTile - 1) / kSimTile, (n + kSimTile - 1) / kSimTile);
    p.out = sims;
    p.out_pitch = pitch;
    if (int rc = launch_sim_tiles<false>(p, stream)) return rc;
    row_topk_kernel<<<rows, kTopkThreads, 0, stream>>>(sims, pitch, n, k, id_list_or_null, id_base, neighbours_out,
                                                        out_rows_or_null, r0);
    MKE_CHECK_LAUNCH("row_topk_kernel");
  }
  return 0;
}
