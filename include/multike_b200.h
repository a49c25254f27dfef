/*
 * multike_b200.h -- C-ABI of the B200-native MultiKE training hot path.
 *
 * The reference (nju-websoft/MultiKE) has no FFI: its "operators" are TensorFlow-1.x graph
 * fragments built in code/MultiKE_model.py and code/losses.py and executed by session.run.
 * Each entry point below names the reference fragment it replaces (file:line under
 * /root/reference/code).  The Python host side (multike_b200/) binds these with ctypes; the
 * binding a reference maintainer would add is shown in INTEGRATION.md.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the parameter says "host"; the caller (torch)
 *     owns all memory and keeps it alive until the stream is synchronised;
 *   - no hidden allocation, no host synchronisation, re-entrant per stream;
 *   - return 0 on success, a negative code on failure (-(int)cudaError_t for CUDA failures,
 *     MKE_EINVAL for bad arguments); never throws; mke_last_error() gives the text;
 *   - tables are row-major fp32 with a row stride of `stride` floats, stride % 4 == 0,
 *     base address 16-byte aligned, columns [dim, stride) are zero and stay zero.
 */
#ifndef MULTIKE_B200_H_
#define MULTIKE_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MKE_ABI_VERSION 5
#define MKE_EINVAL (-100000)
#define MKE_MAX_NEG 32        /* K (negatives per positive) supported by the fused kernel */
#define MKE_MAX_TRY 10        /* base/batch.py:86 max_try=10 */
#define MKE_MAX_SHARDS 8      /* GPUs of one NVLink/NVSwitch box */

typedef struct CUstream_st* mke_stream_t; /* == cudaStream_t */

/*
 * One trainable embedding table as the kernels see it.
 * Replaces one tf.get_variable + (optional) tf.nn.l2_normalize(var, 1) view
 * (base/initializers.py:22-26, MultiKE_model.py:92-99) together with the gradient that
 * optimizer.compute_gradients would build for it (MultiKE_model.py:30).
 */
typedef struct mke_table {
  float*   var;        /* [rows, stride] raw variable V (NOT normalised)                       */
  float*   grad;       /* [rows, stride] dLoss/dE accumulator, E = l2_normalize(V,1) if
                          normalised else V; all-zero between steps; NULL => constant table     */
  uint8_t* touched;    /* [rows] 1 if grad row may be non-zero; all-zero between steps.  NULL =>
                          no flags: phase 2 visits every row (right for small tables such as
                          rel_embeds / attr_embeds, whose hot rows would make the flag byte a
                          contended store target)                                                */
  int32_t  rows;
  int32_t  stride;     /* floats per row, multiple of 4                                         */
  int32_t  dim;        /* logical embedding dimension (<= stride)                               */
  int32_t  normalised; /* 1: the model reads l2_normalize(var,1) (is_l2_norm=True)              */
  int32_t  grad_replicas; /* 0/1: grad is [rows, stride].  R > 1: grad is [R, rows, stride]; phase 1
                          spreads its reductions over the R copies (a thread block adds into
                          copy blockIdx % R) and phase 2 sums and re-zeroes them.  For small hot
                          tables (rel_embeds: the top relation carries 10-16 % of a batch), where
                          thousands of reductions per step would otherwise serialise on one row  */
  /* Row sharding over the GPUs of one box.  n_shards in {0,1}: not sharded.
     n_shards = G in {2,4,8}: `rows` is the GLOBAL row count, row id lives on rank id % G at local
     row id / G; var/grad/touched above are THIS rank's shard (ceil((rows - rank)/G) rows) and
     peer_*[k] are the shards of all ranks as mapped into this process (CUDA IPC; peer_*[shard_rank]
     == var/grad/touched).  Phase 1 gathers rows and reduces gradient rows through these pointers
     over NVLink; phase 2 runs on the local shard only. */
  int32_t  n_shards;
  int32_t  shard_rank;
  int32_t  shard_split; /* 0: plain id % G placement (above).  S > 0: KG-block placement for the two
                          KGs of an alignment task (ids [0,S) = KG1, [S,rows) = KG2; relation triples
                          never cross KGs, base/batch.py:40-41): KG1 rows live on ranks [0,G/2) at
                          rank id % (G/2), local row id / (G/2); KG2 rows on ranks [G/2,G) likewise
                          with id - S.  With G = 2 every rank owns one KG and the relation view
                          needs no peer traffic at all; with G = 4/8 only (G/2-1)/(G/2) of the rows
                          of a positive are remote instead of (G-1)/G.                           */
  int32_t  shard_pad;
  float*   peer_var[MKE_MAX_SHARDS];
  float*   peer_grad[MKE_MAX_SHARDS];
  uint8_t* peer_touched[MKE_MAX_SHARDS];
} mke_table_t;

/*
 * Device-resident set of (h, r, t) triples used to filter negatives
 * (all_triples_set in base/batch.py:86-107).  Open addressing over buckets of four 64-bit keys
 * (one 32-byte sector, filled front to back), linear probing over buckets;
 * key = h<<40 | r<<24 | t, empty slot = UINT64_MAX, capacity (in keys) a power of two >= 8.
 */
typedef struct mke_tripleset {
  uint64_t* slots;
  uint64_t  capacity;  /* power of two, >= 2 * number of triples                                */
} mke_tripleset_t;

/*
 * Candidate pool of one KG for negative sampling (base/batch.py:93-94:
 * neighbor.get(entity, entities_list)).
 */
typedef struct mke_kg_sampler {
  const int32_t* entity_list;   /* [n_entities] ids of this KG, or NULL => ids are base..base+n-1 */
  int32_t        entity_base;
  int32_t        n_entities;
  const int32_t* neighbours;    /* [rows_of_entity_table, n_neighbours] truncated-eps candidates
                                   (base/batch.py:119-150) or NULL => uniform over entity_list;
                                   a row whose first entry is -1 falls back to entity_list      */
  int32_t        n_neighbours;
  mke_tripleset_t set;          /* triples of this KG (incl. the swapped "sup" triples, kg.py:59,134) */
} mke_kg_sampler_t;

int         mke_abi_version(void);
const char* mke_last_error(void);
/* number of kernel launches issued by this library since load (bench.py's gpu_launches) */
uint64_t    mke_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * Relation view, phase 1.
 * ------------------------------------------------------------------------------------------ */

/*
 * Generic TransE-style scored batch = one call of a losses.py function plus its backward:
 *   negative == 0:  loss += scale * sum_i w_i * log(1 + exp(+||H[ih_i] + M[im_i] - T[it_i]||^2))
 *   negative == 1:  loss += scale * sum_i w_i * log(1 + exp(-||H[ih_i] + M[im_i] - T[it_i]||^2))
 * Replaces, by choice of tables/flags:
 *   relation_logistic_loss           losses.py:4-12   (two calls: positives, negatives)
 *   attribute_logistic_loss          losses.py:15-27  (tail table = constant literal table)
 *   relation_logistic_loss_wo_negs   losses.py:30-34
 *   attribute_logistic_loss_wo_negs  losses.py:37-41
 *   logistic_loss_wo_negs            losses.py:44-50
 * and the six tf.nn.embedding_lookup gathers of MultiKE_model.py:123-128 / 164-166 / 194-196.
 * Gradients are accumulated (+=) into each table's grad (skipped when grad == NULL) and the
 * rows are flagged in `touched`.  score_out (optional) receives -||.||^2 per triple
 * (the pos_score / neg_score tensors of losses.py:7-8).
 */
int mke_triple_fwd_bwd(const mke_table_t* head, const mke_table_t* mid, const mke_table_t* tail,
                       const int32_t* ih, const int32_t* im, const int32_t* it, int32_t n,
                       const float* w_or_null, int32_t negative, float scale,
                       double* loss_accum, float* score_out_or_null, mke_stream_t stream);

/*
 * Fused relation-view step, phase 1, with the negative sampler on device:
 * gather -> score -> logistic loss -> gradient -> scatter-add for one batch made of a kg1
 * slice and a kg2 slice of positives (base/batch.py:33-42) and K negatives per positive drawn
 * on the fly with the semantics of generate_neg_triples_fast (base/batch.py:86-116).
 * Replaces: MultiKE_model.py:123-131 (graph), losses.py:4-12, base/batch.py:33-116 and the
 * queue/feed_dict plumbing of MultiKE_model.py:295-310.
 *   pos1/pos2   [len,3] int32 (h,r,t) rows, device
 *   K           negatives per positive, 0..MKE_MAX_NEG (K == 0 => positives only)
 *   seed, step  counter-based RNG coordinates; draws are a pure function of
 *               (seed, step, index of the positive in the batch, try, draw)
 *   w_or_null   per-positive weights [len1+len2] (applied to the positive term only) or NULL
 *   pos_scale   multiplier on the positive term (2 for the ckge/ckgp graphs, MultiKE_model.py:168,198)
 *   neg_out     optional [ (len1+len2) * K, 3 ] int32: the sampled negatives, positive-major
 *               (what base/batch.py:116 returns) -- for parity tests; NULL in production
 *   variant     0 = quarter-warp register path (default; falls back to 2 for strides without an
 *               instantiation), 1 = TMA bulk-copy / bulk-reduce path, 2 = warp-per-positive
 *               LDG / RED.v4 path (any stride <= 256), 3 = variant 0's arithmetic on the persistent
 *               row-stream schedule (pre-drawn negatives with 3 + K >= 8; other launch shapes run
 *               as variant 0)
 */
int mke_rel_step_sampled(const mke_table_t* ent, const mke_table_t* rel,
                         const int32_t* pos1, int32_t len1, const mke_kg_sampler_t* kg1,
                         const int32_t* pos2, int32_t len2, const mke_kg_sampler_t* kg2,
                         int32_t K, uint64_t seed, uint64_t step,
                         const float* w_or_null, float pos_scale,
                         double* loss_accum, int32_t* neg_out_or_null,
                         int32_t variant, mke_stream_t stream);

/*
 * Same fused kernel with the negatives supplied by the caller in structured form
 * (one corrupted entity + side bit per negative): neg_ent [n,K] int32, neg_side [n] uint32 with
 * bit j set when negative j replaces the HEAD.  Used by parity tests and by callers that keep
 * their own sampler.
 */
int mke_rel_step_structured(const mke_table_t* ent, const mke_table_t* rel,
                            const int32_t* pos, int32_t n, int32_t K,
                            const int32_t* neg_ent, const uint32_t* neg_side,
                            const float* w_or_null, float pos_scale,
                            double* loss_accum, int32_t variant, mke_stream_t stream);

/* The same for a batch given as a kg1 slice followed by a kg2 slice (base/batch.py:33-42);
 * neg_ent / neg_side are indexed by the position in the concatenated batch. */
int mke_rel_step_structured2(const mke_table_t* ent, const mke_table_t* rel,
                             const int32_t* pos1, int32_t len1, const int32_t* pos2, int32_t len2,
                             int32_t K, const int32_t* neg_ent, const uint32_t* neg_side,
                             const float* w_or_null, float pos_scale,
                             double* loss_accum, int32_t variant, mke_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Phase 2: optimizer.
 * ------------------------------------------------------------------------------------------ */

/*
 * For every flagged row: back-propagate the accumulated gradient through l2_normalize
 * (if table->normalised), apply Adagrad, zero the gradient row, clear the flag.
 *   u = v * rsqrt(max(|v|^2, 1e-12));  g_v = (g - u (u.g)) * rsqrt(max(|v|^2,1e-12))   [|v|^2 >= 1e-12]
 *   acc += g_v^2 ;  v -= lr * g_v * rsqrt(acc)
 * Replaces: the l2_normalize backward + tf.train.AdagradOptimizer.apply_gradients of
 * MultiKE_model.py:15-31 (dense ApplyAdagrad on every row; rows with zero gradient are a
 * mathematical no-op and are skipped).  `acc` is this optimizer's accumulator slot
 * ([rows,stride], initial value 0.1): the reference creates one per generate_optimizer call.
 */
int mke_rows_apply_adagrad(const mke_table_t* table, float* acc, float lr, mke_stream_t stream);

/*
 * The same for two tables in ONE launch when their strides agree (the entity and the relation
 * table of a view: both appear in every generate_optimizer var list, MultiKE_model.py:30);
 * otherwise two launches.  Each table keeps its own accumulator slot and learning rate.
 */
int mke_rows_apply_adagrad_pair(const mke_table_t* a, float* acc_a, float lr_a,
                                const mke_table_t* b, float* acc_b, float lr_b, mke_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Attribute view, the LIVE score of the reference: conv() (MultiKE_model.py:34-63) + the loss of the
 * attribute graphs (:144-149 weighted; :183 scale 2, no weights; :214-218 weighted) + full backward.
 *   loss += scale * sum_i w_i log(1 + exp(|E[ih_i] - conv(A[ia_i], V[iv_i])|^2))
 * ent = av_ent_embeds (normalised view), attr = attr_embeds (raw: "False important!", :96), val =
 * literal_embeds (constant).  theta / gtheta: the conv() instance's parameters and their gradient
 * accumulator as ONE flat fp32 vector of mke_attr_cnn_param_count(dim) floats laid out
 *   gamma[D] beta[D] k1[2][4][1][2] b1[2] k2[2][4][2][2] b2[2] wd[4D][D] bd[D]
 * (22 777 at D = 75; three independent instances exist in the reference).
 * Gradient rows go to ent->grad / attr->grad as in the other kernels; gtheta is accumulated (+=).
 * workspace: mke_attr_cnn_workspace_floats(n, dim) floats, 16-byte aligned (activations as TF32 part + remainder, their
 * transposes, partial sums: ~2 400 floats per sample at D = 75).  The dense layer (4D -> D) runs as three tcgen05 GEMMs at
 * fp32-equivalent precision, the convolutions one warp per sample (csrc/mke_cnn.cu).  Phase 2: mke_rows_apply_adagrad for
 * the tables, mke_dense_apply_adagrad for theta (acc0 = 0.1).
 * ------------------------------------------------------------------------------------------ */
int64_t mke_attr_cnn_param_count(int32_t dim);
int64_t mke_attr_cnn_workspace_floats(int32_t n, int32_t dim);
int mke_attr_cnn_fwd_bwd(const mke_table_t* ent, const mke_table_t* attr, const mke_table_t* val,
                         const int32_t* ih, const int32_t* ia, const int32_t* iv, int32_t n,
                         const float* w_or_null, float scale, const float* theta, float* gtheta,
                         float* workspace, double* loss_accum, mke_stream_t stream);
/* dense Adagrad (acc += g^2; theta -= lr g rsqrt(acc); g = 0) for a flat parameter vector */
int mke_dense_apply_adagrad(float* theta, float* grad, float* acc, int64_t n, float lr, mke_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * ITC cross-view alignment step (MultiKE_model.py:225-239 + losses.py:66-69), fused:
 *   loss += scale * sum_i ( name_weight |F_i - N_i|^2 + |F_i - R_i|^2 + |F_i - A_i|^2 ),  i = idx[.]
 * F = ent_embeds, N = name_embeds (constant: grad NULL), R = rv_ent_embeds, A = av_ent_embeds, each
 * read through its normalised view; gradients are accumulated into the trainable tables (phase 2:
 * mke_rows_apply_adagrad with args.ITC_learning_rate and this graph's accumulator slots).
 * scale = args.cv_weight, name_weight = args.cv_name_weight.
 * ------------------------------------------------------------------------------------------ */
int mke_align_fwd_bwd(const mke_table_t* shared, const mke_table_t* name, const mke_table_t* rv,
                      const mke_table_t* av, const int32_t* idx, int32_t n, float name_weight,
                      float scale, double* loss_accum, mke_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * SSL space-mapping step: _define_space_mapping_graph + train_shared_space_mapping_1epo
 * (MultiKE_model.py:241-261, :439-454), space_mapping_loss / orthogonal_loss (losses.py:53-63).
 *   loss += sum over the views X in {name, rv, av} of
 *           |F - gl2n(X M_X)|^2 + orthogonal_weight |M_X M_X^T - I|^2 + norm_w |M_X|^2
 * gl2n = l2_normalize over the WHOLE batch (tf.nn.l2_normalize without axis); every table is read through its
 * normalised view at the ids idx[0..n).  Outputs: gradient rows of the shared table (accumulated into shared->grad,
 * rows flagged; the view tables do not train: MultiKE_model.py:257) and maps_grad [3][dim][dim] (overwritten) for
 * maps [3][dim][dim] (row-major, order name, rv, av).  workspace: mke_space_mapping_workspace_floats(n, dim) floats,
 * 8-byte aligned.  Three launches (rows -> products and batch-wide sums; one block: the dim x dim algebra; rows ->
 * gradient rows), no host sync.
 * ------------------------------------------------------------------------------------------ */
int64_t mke_space_mapping_workspace_floats(int32_t n, int32_t dim);
int mke_space_mapping_fwd_bwd(const mke_table_t* shared, const mke_table_t* name, const mke_table_t* rv,
                              const mke_table_t* av, const int32_t* idx, int32_t n, const float* maps,
                              float* maps_grad, float orthogonal_weight, float norm_w, float* workspace,
                              double* loss_accum, mke_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * losses.py as free functions on already gathered [n, dim] matrices (row stride ld floats).
 * ------------------------------------------------------------------------------------------ */

/*
 * loss += scale * sum_i w_i log(1 + exp(+-|H_i + M_i - T_i|^2))  (sign: + positives, - negatives)
 * and, where the pointers are non-NULL, gH/gM/gT [n, ld] are OVERWRITTEN with d loss / d row.
 * One call per term of relation_logistic_loss / attribute_logistic_loss / *_wo_negs (losses.py:4-50).
 */
int mke_dense_logistic_fwd_bwd(const float* H, const float* M, const float* T, int32_t n, int32_t dim,
                               int32_t ld, const float* w_or_null, int32_t negative, float scale,
                               double* loss_accum, float* gH, float* gM, float* gT, mke_stream_t stream);

/* alignment_loss (losses.py:66-69): loss += scale * sum_i |A_i - B_i|^2, gA = 2 scale (A - B), gB = -gA */
int mke_dense_sqdist_fwd_bwd(const float* A, const float* B, int32_t n, int32_t dim, int32_t ld,
                             float scale, double* loss_accum, float* gA, float* gB, mke_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Step / epoch driver.
 * ------------------------------------------------------------------------------------------ */

/*
 * Everything train_relation_view_1epo (MultiKE_model.py:291-317) touches, as the kernels see it.
 * The triple lists of both KGs live either in HBM (triples1/2) or in pinned HOST memory
 * (host_triples1/2, then stage1/2 must hold two device staging buffers of b1*3 / b2*3 int32 each,
 * b1 = int(n1/(n1+n2)*batch_size), b2 = batch_size - b1, base/batch.py:36-37).
 */
typedef struct mke_rel_view {
  const mke_table_t* ent;          /* rv_ent_embeds (MultiKE_model.py:92)                       */
  const mke_table_t* rel;          /* rel_embeds    (MultiKE_model.py:93)                       */
  float*  ent_acc;                 /* Adagrad slot of this graph for ent ([rows,stride])        */
  float*  rel_acc;
  float   lr;                      /* args.learning_rate                                        */
  const int32_t* triples1;         /* device [n1,3] or NULL                                     */
  const int32_t* triples2;
  const int32_t* host_triples1;    /* pinned host [n1,3] or NULL                                */
  const int32_t* host_triples2;
  int32_t* stage1[2];              /* device staging for host-fed batches                       */
  int32_t* stage2[2];
  int32_t n1, n2;
  const mke_kg_sampler_t* kg1;
  const mke_kg_sampler_t* kg2;
  int32_t batch_size;              /* args.batch_size                                           */
  int32_t K;                       /* args.neg_triple_num                                       */
  uint64_t seed;
  int32_t*  neg_ent[2];            /* [batch_size*K] x2, or NULL => negatives drawn inside phase 1 */
  uint32_t* neg_side[2];           /* [batch_size]   x2                                         */
  double* step_loss;               /* device [>= n_steps]: mke_rel_train_steps zeroes [0, n_steps) and leaves the loss of step s in step_loss[s] */
  double* host_step_loss;          /* pinned host [>= n_steps] or NULL: step_loss[s] is copied
                                      here (async, 8 bytes) after every step                    */
  int32_t variant;                 /* phase-1 schedule; 4 = persistent step kernel (below)         */
  /* variant 4: ONE cooperative launch per `persist_chunk` (<= 128) consecutive steps (negatives, phase 1,
   * phase 2 separated by grid barriers instead of launches).  persist_ws: device workspace of
   * mke_rel_persist_workspace_bytes() bytes, zero-initialised once by the caller.  Its first 2048 bytes
   * are the barrier / queue words; at byte 2048 the last launch leaves uint64 globaltimer stamps
   * [2 * steps + 2]: launch start, after the negatives of its first step, then after phase 1 and after
   * phase 2 of every step.  Host-fed steps additionally need persist_flag_src: 256 uint32 in pinned
   * host memory holding 0..255 (the source of the 4-byte "batch k has landed" copies).            */
  void*    persist_ws;
  int64_t  persist_ws_bytes;
  int32_t  persist_chunk;
  const uint32_t* persist_flag_src;
} mke_rel_view_t;

/*
 * The relation view on the G GPUs of one box (multike_b200/sharded.py): entity table row-sharded over peer-mapped
 * memory, relation table replicated.  n_steps GLOBAL steps of `global_batch` positives (each rank trains its part:
 * sharded.py rank_parts / group_parts), issued without a collective library: xchg[k] is rank k's peer-mapped
 * exchange buffer of 2 * rel->rows * rel->stride floats (two step parities), sync[k] rank k's peer-mapped array of
 * 2 * MKE_MAX_SHARDS uint32 words (zero-initialised: barrier flags, then two words of this rank's own); *barrier_seq (host, starts at 0, the same on every rank) counts the
 * flag barriers issued so far.  Every rank must make the same sequence of calls.  With `side` != main the
 * negatives of step s+1 are drawn while step s is exchanged and applied.  step_loss[s] (device) receives this
 * rank's share of the loss of step s; positives_out the positives this rank answers for.
 * (sharded.py: ShardedRelationView; tests/multi_gpu_check.py compares with one GPU at batch world * B.)
 */
typedef struct mke_rel_sharded_view {
  const mke_table_t* ent;          /* n_shards = world, peer pointers set                                  */
  const mke_table_t* rel;          /* this rank's replica                                                  */
  float*  ent_acc;                 /* Adagrad slot of the local shard                                      */
  float*  rel_acc;
  float   lr;
  const int32_t* triples1;         /* device [n1,3]: the WHOLE lists, on every rank                        */
  const int32_t* triples2;
  int32_t n1, n2;
  const mke_kg_sampler_t* kg1;
  const mke_kg_sampler_t* kg2;
  int32_t global_batch;            /* world * per-rank batch                                               */
  int32_t K;
  uint64_t seed;
  int32_t world, rank;
  int32_t by_kg;                   /* positives trained on the ranks of their own KG (KG-block placement)  */
  int32_t owner_negs;              /* "negatives where they live" (mke_neg_keep_owned)                     */
  int32_t dummy_row;               /* a row of this shard (owner_negs)                                     */
  int32_t variant;                 /* phase-1 schedule                                                     */
  int32_t*  neg_ent[2];
  uint32_t* neg_side[2];
  uint32_t* neg_valid[2];          /* owner_negs only                                                      */
  double* step_loss;               /* device [>= n_steps]                                                  */
  float*    xchg[MKE_MAX_SHARDS];
  uint32_t* sync[MKE_MAX_SHARDS];
  /* host-fed steps (optional): the lists in pinned HOST memory; every step this rank's part is copied into
   * staging set (step & 1) ([cap * 3] int32 each, cap = the most positives one rank walks in a step) and its
   * share of the step loss is copied back to host_step_loss[s] */
  const int32_t* host_triples1;
  const int32_t* host_triples2;
  int32_t* stage1[2];
  int32_t* stage2[2];
  double*  host_step_loss;
} mke_rel_sharded_view_t;

int mke_rel_sharded_train_steps(const mke_rel_sharded_view_t* view, int32_t first_step, int32_t n_steps,
                                uint64_t first_global_step, uint32_t* barrier_seq, int64_t* positives_out,
                                mke_stream_t main, mke_stream_t side);

/* bytes of mke_rel_view_t.persist_ws for triple lists of n1 / n2 rows (0 on bad arguments) */
int64_t mke_rel_persist_workspace_bytes(int32_t n1, int32_t n2, int32_t batch_size, int32_t chunk_steps);

/*
 * n_steps consecutive training steps (batch -> negatives -> phase 1 -> phase 2), starting at step
 * `first_step` of an epoch and wrapping to step 0 after the last step of the epoch
 * (ceil((n1+n2)/batch_size) steps, MultiKE_CSL.py:40) -- the inner loop of
 * MultiKE_model.py:302-313 without any host round trip.  Step s uses RNG coordinate
 * first_global_step + s.  With neg_ent != NULL the negatives (and, for host-fed batches, the
 * H2D copy) of step s+1 run on `side` while step s trains on `main`; both streams are joined
 * again before return (the call is capturable into a CUDA graph).  positives_out (host, may be
 * NULL) receives the number of positives trained (trained_samples_num, MultiKE_model.py:311).
 */
int mke_rel_train_steps(const mke_rel_view_t* view, int32_t first_step, int32_t n_steps,
                        uint64_t first_global_step, int64_t* positives_out,
                        mke_stream_t main, mke_stream_t side);

/*
 * Measurement aid (the only entry points that synchronise): after mke_timing_enable(n) the driver
 * brackets each of its next n phase-1 launches with CUDA events on `main`; mke_timing_read waits
 * for them, returns the summed elapsed time and the number of launches, and re-arms.
 */
int mke_timing_enable(int32_t max_launches);
int mke_timing_read(double* total_ms, int32_t* launches);
/* time one phase-1 launch in `every` (default 1): the pair of timed events is itself a few microseconds of the step */
int mke_timing_stride(int32_t every);

/* ------------------------------------------------------------------------------------------
 * Sampler pieces (base/batch.py:86-116, attr_batch.py:13-25) usable on their own.
 * ------------------------------------------------------------------------------------------ */

/* Insert n (h,r,t) rows into an empty (all 0xFF) slot array. */
int mke_tripleset_build(const mke_tripleset_t* set, const int32_t* triples, int32_t n,
                        mke_stream_t stream);
/* out[i] = 1 if triples[i] is in the set else 0. */
int mke_tripleset_contains(const mke_tripleset_t* set, const int32_t* triples, int32_t n,
                           uint8_t* out, mke_stream_t stream);

/*
 * Stand-alone relation negative sampler: writes neg_out [(len1+len2)*K, 3] exactly as the
 * fused kernel would draw them (same RNG coordinates).
 */
int mke_sample_uniform(const int32_t* pos1, int32_t len1, const mke_kg_sampler_t* kg1,
                       const int32_t* pos2, int32_t len2, const mke_kg_sampler_t* kg2,
                       int32_t K, uint64_t seed, uint64_t step, int32_t* neg_out,
                       mke_stream_t stream);

/*
 * The same draws in the structured form mke_rel_step_structured consumes: neg_ent [(len1+len2), K]
 * int32 (corrupted entity of negative j) and neg_side [(len1+len2)] (bit j set: negative j
 * replaces the head).  Sampling reads no embedding table, so a driver can run it for step s+1 on
 * a second stream while step s trains (multike_b200/relation_view.py does).
 */
int mke_sample_structured(const int32_t* pos1, int32_t len1, const mke_kg_sampler_t* kg1,
                          const int32_t* pos2, int32_t len2, const mke_kg_sampler_t* kg2,
                          int32_t K, uint64_t seed, uint64_t step, int32_t* neg_ent,
                          uint32_t* neg_side, mke_stream_t stream);

/* mke_sample_structured for a launch that holds positions [index_base, index_base + len1 + len2)
 * of a larger (global) batch: the RNG coordinate of local positive i is index_base + i, so the
 * ranks of a multi-GPU step draw exactly what one GPU would draw for the whole batch. */
int mke_sample_structured_at(const int32_t* pos1, int32_t len1, const mke_kg_sampler_t* kg1,
                             const int32_t* pos2, int32_t len2, const mke_kg_sampler_t* kg2,
                             int32_t K, uint64_t seed, uint64_t step, int32_t index_base,
                             int32_t* neg_ent, uint32_t* neg_side, mke_stream_t stream);

/*
 * Attribute-view negatives (attr_batch.py:13-25 generate_neg_attribute_triples): for every positive
 * (h, a, v) K corrupted HEADS, each drawn from the candidate pool of h (neighbour list or the KG's
 * entity list) and redrawn while (h', a, v) is a known attribute triple of that KG (kgX->set built
 * over (h, a, v) rows); independent draws, i.e. with replacement, as random.choice does.
 * neg_head [(len1+len2) * K], positive-major.  The reference retries without bound; try 63 is
 * accepted unfiltered.  pos rows are (h, a, v) int32; weights ride along on the caller's side.
 */
int mke_sample_attribute_heads(const int32_t* pos1, int32_t len1, const mke_kg_sampler_t* kg1,
                               const int32_t* pos2, int32_t len2, const mke_kg_sampler_t* kg2,
                               int32_t K, uint64_t seed, uint64_t step, int32_t index_base,
                               int32_t* neg_head, mke_stream_t stream);

/*
 * "Negatives where they live", for row-sharded entity tables: every rank
 * of a KG's group walks ALL positives of the group's batch but scores only the negatives whose
 * corrupted entity it owns -- the K rows that dominate a positive's traffic never cross NVLink;
 * what still does are the two endpoint rows of a positive (when another rank owns them).
 *
 * mke_neg_keep_owned rewrites a sampled batch for one rank: negatives owned by other shards are
 * replaced by `dummy_id` (a row of this shard) and their bit in neg_valid[i] is cleared.
 * mke_rel_step_structured3 is mke_rel_step_structured2 with that mask and with the range
 * [pos_own_lo, pos_own_hi) of positives whose POSITIVE term (losses.py:7) belongs to this launch;
 * summed over the ranks of the group the gradients and the loss equal one full step.
 */
int mke_neg_keep_owned(int32_t* neg_ent, int32_t n, int32_t K, int32_t n_shards, int32_t shard_split,
                       int32_t my_shard, int32_t dummy_id, uint32_t* neg_valid, mke_stream_t stream);
int mke_rel_step_structured3(const mke_table_t* ent, const mke_table_t* rel,
                             const int32_t* pos1, int32_t len1, const int32_t* pos2, int32_t len2,
                             int32_t K, const int32_t* neg_ent, const uint32_t* neg_side,
                             const uint32_t* neg_valid_or_null, int32_t pos_own_lo, int32_t pos_own_hi,
                             const float* w_or_null, float pos_scale,
                             double* loss_accum, int32_t variant, mke_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Peer memory for row-sharded tables (one process per GPU; handles travel over torch.distributed).
 * ------------------------------------------------------------------------------------------ */
/* ------------------------------------------------------------------------------------------
 * Tensor-core GEMM of the literal auto-encoder (code/literal_encoder.py:45-69: six affine layers, forward and
 * backward) -- hand-written tcgen05 / TMA / TMEM kernel at fp32-equivalent precision (3xTF32 split).
 *   C[M,N] (ldc) = A[M,K] . B[N,K]^T (+ bias[N])     A, B row-major with K contiguous
 * Operands are passed as (hi, lo) pairs produced by mke_split_tf32 (hi: low 13 mantissa bits cleared, lo = x - hi);
 * lda / ldb multiples of 4 floats, base pointers 16-byte aligned.
 * ------------------------------------------------------------------------------------------ */
int mke_split_tf32(const float* x, float* hi, float* lo, int64_t n, mke_stream_t stream);
int mke_gemm_tf32x3(const float* a_hi, const float* a_lo, int64_t lda, const float* b_hi, const float* b_lo,
                    int64_t ldb, int32_t M, int32_t N, int32_t K, const float* bias_or_null, float* C,
                    int64_t ldc, mke_stream_t stream);

/* mke_neg_keep_owned, COMPACTED: this rank's negatives first (original order), dummies behind, the side word
 * permuted along; neg_valid[i] = low_ones(count).  With compact = 1 mke_rel_step_structured4 (otherwise
 * mke_rel_step_structured3) lets the one-wave kernel (variant 0) stop its K-loop at the longest list of a warp's
 * four positives. */
int mke_neg_keep_owned2(int32_t* neg_ent, uint32_t* neg_side, int32_t n, int32_t K, int32_t n_shards,
                        int32_t shard_split, int32_t my_shard, int32_t dummy_id, uint32_t* neg_valid,
                        mke_stream_t stream);
int mke_rel_step_structured4(const mke_table_t* ent, const mke_table_t* rel, const int32_t* pos1, int32_t len1,
                             const int32_t* pos2, int32_t len2, int32_t K, const int32_t* neg_ent,
                             const uint32_t* neg_side, const uint32_t* neg_valid_or_null, int32_t compact,
                             int32_t pos_own_lo, int32_t pos_own_hi, const float* w_or_null, float pos_scale,
                             double* loss_accum, int32_t variant, mke_stream_t stream);

/* random.sample(range(n), count) of the reference's cross-KG / entity batches (MultiKE_model.py:355-358, :377, :399,
 * :443, :462): out[i], i < count, = distinct uniform indices of [0, n) -- the first images of a keyed pseudo-random
 * permutation (Feistel network, cycle walking), a function of (seed, draw) only. */
int mke_sample_distinct(int32_t n, int32_t count, uint64_t seed, uint64_t draw, int32_t* out, mke_stream_t stream);

int mke_peer_alloc(uint64_t bytes, void** ptr);                  /* cudaMalloc + zero fill          */
int mke_peer_free(void* ptr);
int mke_ipc_export(const void* ptr, unsigned char handle[64]);   /* cudaIpcGetMemHandle             */
int mke_ipc_open(const unsigned char handle[64], void** ptr);    /* cudaIpcOpenMemHandle            */
int mke_ipc_close(void* ptr);

/* ------------------------------------------------------------------------------------------
 * Row staging for row-sharded tables (BASELINE configs[3]: the attribute-view CNN MultiKE_model.py:134-151, the
 * cross-KG graphs :158-221, ITC :225-239 and the SSL mapping :241-261 on sharded entity tables).  Those graphs
 * run unchanged on a plain local table that holds the rows of their batch:
 *   mke_table_stage_rows   staged.var[i] = raw row ids[i] of `table` (read through the owner's peer mapping when
 *                          table->n_shards > 1), staged.grad[i] = 0, for i < n <= staged->rows; ids may repeat;
 *   mke_table_commit_grads table.grad[ids[i]] += staged.grad[i] for the ids THIS rank owns (all ids of a plain
 *                          table), touched flags set; follow with mke_rows_apply_adagrad on `table`;
 *   mke_peer_barrier       rank `rank` stores seq into slot `rank` of every rank's flag array (flags[k]: >= 8
 *                          uint32 of peer-mapped memory of rank k, zero-initialised) and waits until every slot of
 *                          its own array has reached seq; seq must grow by one per call on every rank.
 * ---------------------------------------------------------------------------------------- */
int mke_table_stage_rows(const mke_table_t* table, const int32_t* ids, int32_t n, const mke_table_t* staged,
                         mke_stream_t stream);
int mke_table_commit_grads(const mke_table_t* table, const int32_t* ids, int32_t n, const mke_table_t* staged,
                           mke_stream_t stream);
int mke_peer_barrier(void* const* flags, int32_t world, int32_t rank, uint32_t seq, mke_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Similarity search on dense fp32 rows (evaluation and truncated-epsilon neighbours).
 * Inputs are row-major [*, stride] device arrays of which the first `dim` columns count
 * (dim <= 128); idx_or_null gathers rows (NULL => rows 0..n-1 in order); normalize != 0 divides
 * every gathered row by its l2 norm first (sklearn.preprocessing.normalize, base/similarity.py:
 * 31-33; a zero row stays zero).  Sims are fp32 inner products accumulated in ascending column
 * order of the embedding (one fmaf chain per pair), so bit-equal rows give bit-equal sims.
 * ------------------------------------------------------------------------------------------ */

/* Which implementation mke_sim_rank / mke_sim_topk run: 1 (default) = tcgen05 tiles at fp32-equivalent precision
 * (3xTF32 split, csrc/mke_sim_tc.cu; mke_sim_topk then never writes the similarity matrix: sampled per-row threshold,
 * candidate lists from the tile epilogue, exact select per row, exact path for the rare row whose list came out short),
 * 2 = the same tiles but mke_sim_topk materialises rows of sims and selects from them, 0 = the fp32 FMA tiles of
 * csrc/mke_sim.cu (the measured baseline).  on < 0 only queries.  Returns the previous setting.  The same tie rules hold for both (equal rows give bit-equal sims). */
int mke_sim_use_tensor_cores(int32_t on);

/* floats of workspace mke_sim_rank needs (prepared copies of both row sets + per-row scratch) */
int64_t mke_sim_rank_workspace_floats(int32_t n1, int32_t n2, int32_t dim);

/*
 * The Hits@k / MR / MRR evaluator without the similarity matrix.  Replaces
 * base/similarity.py:9-52 sim(metric='inner'), base/alignment.py:8-79 greedy_alignment and
 * :141-163 calculate_rank (callers: base/evaluation.py:6-28 valid/test from MultiKE_Late.py:14-61).
 *   gold_or_null  [n1] column (row of emb2 after the gather) aligned with each row of emb1;
 *                 NULL => gold[i] = i (what calculate_rank assumes)
 *   rank_out      [n1] number of columns ranked before the gold one in the stable descending
 *                 order of the sims: #{j : s_ij > s_ig} + #{j < g : s_ij == s_ig}
 *                 (rank_index of alignment.py:153; Hits@k <=> rank < k, MR = mean(rank + 1))
 *   top1_out      [n1] arg max_j s_ij, smallest column on ties (rank[0] of alignment.py:151,
 *                 the "alignment_rest" pairs)
 */
int mke_sim_rank(const float* emb1, const int32_t* idx1_or_null, int32_t n1,
                 const float* emb2, const int32_t* idx2_or_null, int32_t n2,
                 int32_t stride, int32_t dim, int32_t normalize, const int32_t* gold_or_null,
                 float* workspace, int32_t* rank_out, int32_t* top1_out, mke_stream_t stream);

/* floats of workspace for mke_sim_topk when `chunk_rows` rows of sims are materialised at a time */
int64_t mke_sim_topk_workspace_floats(int32_t n, int32_t dim, int32_t chunk_rows);

/*
 * Truncated-epsilon candidate lists: for every gathered row i its k most similar rows (itself
 * included, as in the reference).  Replaces base/batch.py:119-150 generate_neighbours /
 * find_neighbours (np.matmul + np.argpartition in 4 processes; caller MultiKE_CSL.py:89-99).
 *   id_list_or_null / id_base   entity id of column c: id_list[c], or id_base + c
 *   out_rows_or_null            row of neighbours_out that receives the list of gathered row i
 *                               (NULL => row i); with out_rows = the entity ids the output is the
 *                               table mke_kg_sampler_t.neighbours expects
 *   neighbours_out              [*, k] int32; a list holds the k best columns in ASCENDING column
 *                               order (ties at the k-th sim: smallest columns first) --
 *                               deterministic where np.argpartition's order is unspecified
 *   workspace, workspace_floats at least mke_sim_topk_workspace_floats(n, dim, 128); more
 *                               workspace => more rows of sims per pass
 */
int mke_sim_topk(const float* emb, const int32_t* idx_or_null, int32_t n, int32_t stride, int32_t dim,
                 int32_t normalize, int32_t k, const int32_t* id_list_or_null, int32_t id_base,
                 const int32_t* out_rows_or_null, float* workspace, int64_t workspace_floats,
                 int32_t* neighbours_out, mke_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Table utilities.
 * ------------------------------------------------------------------------------------------ */

/*
 * out[i, 0:dim] = l2_normalize(var[idx[i]]) (or the raw row if !normalised); idx == NULL => all
 * rows in order.  out is dense [n, dim].  Replaces tensor.eval(session=) /
 * tf.nn.embedding_lookup(table, ids).eval() (MultiKE_model.py:263-287, MultiKE_Late.py:16-26).
 */
int mke_table_export(const mke_table_t* table, const int32_t* idx_or_null, int32_t n,
                     float* out, mke_stream_t stream);

/* acc[r, c] = value for c < dim, 0 for pad columns (Adagrad initial_accumulator_value = 0.1). */
int mke_fill_rows(float* buf, int32_t rows, int32_t stride, int32_t dim, float value,
                  mke_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MULTIKE_B200_H_ */
