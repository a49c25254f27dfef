// Persistent step kernel of the relation view: parameters shared by the kernel (mke_rel_persist.cu)
// and the step driver (mke_epoch.cu).
#pragma once
#include "mke_apply.cuh"
#include "mke_rel_q8p.cuh"

namespace mke {

// words of PersistParams::sync (uint32, zeroed before every launch)
constexpr int kPersistMaxChunk = 128;  // steps per launch at most
constexpr int kSyncBarrier = 0;   // arrivals of the grid barrier (monotonic, one per block and barrier)
constexpr int kSyncError = 1;     // != 0: a bounded wait ran out (the kernel traps right after)
constexpr int kSyncQueues = 64;   // (own cache lines, away from the polled barrier word) + 2 k + {0: apply queue of step k, 1: sampling queue of step k} of the launch
constexpr int kSyncWords = 512;   // >= kSyncQueues + 2 (kPersistMaxChunk + 1)
constexpr int kSyncBytes = kSyncWords * 4;

struct PersistParams {
  // positives: device-resident lists (t1/t2, indexed by list position), or -- host fed -- staging
  // buffers that hold step k of this launch at k * b1 * 3 (kg1) / k * b2 * 3 (kg2), valid once
  // flags[k] == flag_value
  const int32_t* t1;
  const int32_t* t2;
  const int32_t* st1;
  const int32_t* st2;
  const uint32_t* flags;
  uint32_t flag_value;
  int n1, n2, b1, b2;
  int steps_per_epoch, first_step, n_steps;
  uint64_t seed, first_global_step;
  int32_t* neg_ent[2];
  uint32_t* neg_side[2];
  ApplyTable A, B;     // entity table, relation table (phase 2)
  uint32_t* sync;      // kSyncWords words
  double* step_loss;   // device, step s of this launch adds to step_loss[s]
  double* host_loss;   // device-visible pinned host memory or NULL: step_loss[s] is stored there after phase 1
  unsigned long long* trace;  // NULL or [2 * n_steps + 2] globaltimer stamps: start, then after each barrier
  int samp_phase;      // sampler warps (experiment, MKE_PERSIST_SPLIT): 1 = sample under phase 2 only, 0 = all through the step
  int fence_mode;      // barrier fences: 0 = fence.sc (__threadfence), 1 = fence.acq_rel (measured: no difference)
  int apply_mode;      // 1: phase 2 of the flagged table as a cp.async row stream, 0: load-compute-store per row
  int apply_chunk;     // flag bytes per apply ticket: 16 or 32
  unsigned long long* block_trace;  // debug (built with -DMKE_PERSIST_TRACE, MKE_PERSIST_BLOCKTRACE=<file>): [barrier][block][4] stamps of thread 0: arrived, fenced, (barrier 2: own apply share done), released, else NULL
};

// returns 1 when the launch shape has no instantiation (caller falls back to one launch per phase)
int launch_rel_persist(const RelStepParams& p, const PersistParams& q, cudaStream_t stream);

void fill_tables(RelStepParams& p, const mke_table_t* ent, const mke_table_t* rel);

}  // namespace mke
