/*
 * recnow_b200.h — C ABI of librecnow_b200.so: the B200 (sm_100a) implementation of rec_now's in-batch
 * ranking-loss hot path.  This header is the drop-in boundary: plain C, plain pointers and sizes, no
 * TensorFlow / torch / C++ types.  A TensorFlow custom op (rec_now_b200/tf_ops/recnow_tf_ops.cc), the torch
 * front end (the rec_now_b200/rec_block modules via ctypes) and the tests all bind exactly these symbols.
 *
 * Reference interfaces replaced (all under /root/reference/rec_now/rec_block/):
 *   pairwise_loss_from_batch.py:228-279  pairwise_loss(...)            -> rn_pairwise_fwd_bwd
 *   pairwise_loss_from_batch.py:96-127   bpr_loss_func(...)            -> fused into rn_pairwise_fwd_bwd
 *   pairwise_loss_from_batch.py:130-151  occurance_power_weight(...)   -> rn_occurrence_power_weight
 *   pairwise_loss_from_batch.py:43-74, 206-217  generate_pair_mask + boolean_mask pair extraction
 *                                                                      -> rn_pair_indices_count / _fill
 *   listwise_loss_from_batch.py:89-148   to_listwise_sample(...)       -> rn_listwise_fwd_bwd (segmented),
 *                                                                         rn_listwise_dense (compat layout)
 *   listwise_loss_from_batch.py:151-173  listwise_loss_via_softmax_cross_entropy_with_logits
 *                                                                      -> rn_listwise_fwd_bwd
 *   embedding_util.py:239-324            embedding_using_sparse_batch_segment_ids -> rn_segment_pool_fwd / _bwd
 *
 * Conventions
 *   - Every pointer is a DEVICE pointer unless the name ends in _host.  The caller owns all memory,
 *     including the scratch arena (size from rn_*_scratch_bytes).  The library never allocates, never
 *     synchronises the stream (except the _host count read-backs, which say so).  Calls on different
 *     streams with different scratch arenas are independent.  State the library does keep, per host thread: a small
 *     cache of instantiated CUDA graphs (one per device and kernel variant, never freed), read-once getenv() tuning
 *     knobs, and the process-global measurement aid rn_profile_*.  State the CALLER may let it keep: a scratch arena
 *     declared persistent (rn_pairwise_args.scratch_persistent) carries its clean regions from call to call.
 *   - All work is enqueued on `stream` (a cudaStream_t passed as void*); calls are CUDA-graph capturable.
 *   - Return value: RN_OK or an RN_ERR_* code; rn_strerror() names it.  No exceptions, no abort().
 *   - Device pointers must be 16-byte aligned (RN_ERR_ALIGN otherwise).
 *   - Group keys are canonical int64 (equality of the int64 == equality in the reference's sense).  Float
 *     group ids (what the reference uses, pairwise_loss_from_batch.py:33-35) are canonicalised by
 *     rn_canon_keys_f32/_f64: value equality, -0.0 == +0.0, NaN/+-inf match nothing (row_ok = 0).
 */
#ifndef RECNOW_B200_H_
#define RECNOW_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RN_VERSION 104 /* 0.1.4: RN_LABEL_LUT / weight_lut (label-level weight table), RN_LABEL_LAMBDA */

enum {
  RN_OK = 0,
  RN_ERR_ARG = 1,          /* NULL / negative / inconsistent argument */
  RN_ERR_ALIGN = 2,        /* device pointer not 16-byte aligned */
  RN_ERR_SCRATCH = 3,      /* scratch arena too small */
  RN_ERR_LAUNCH = 4,       /* CUDA launch / runtime error (cudaGetLastError) */
  RN_ERR_UNSUPPORTED = 5,  /* valid request outside the fused menu */
  RN_ERR_NO_DEVICE = 6,    /* no sm_100 device / wrong architecture */
  RN_ERR_INTERNAL = 7      /* device-side consistency check failed (see rn_last_device_error) */
};

/* label_pair_to_weight_func menu (pairwise_loss_from_batch.py:175-194).
 *   STEP: C = y_i > y_j, W = 1            (the default, label_pair_to_weight_func=None)
 *   DIFF: W = (y_i - y_j)*[y_i > y_j]     (label-gain weights)
 * Optional per-sample factors rw_pos (row/positive side) and rw_neg (column/negative side) multiply W;
 * whenever any weight is present the reference's rule C = (W > 0) applies (a non-positive or NaN factor
 * removes the pairs it touches).  Anything else goes through rn_pair_indices_* + the caller's own code. */
enum { RN_LABEL_STEP = 0, RN_LABEL_DIFF = 1, RN_LABEL_GAIN2 = 2, RN_LABEL_LUT = 3, RN_LABEL_LAMBDA = 4 };
/*   GAIN2: W = (2^y_i - 2^y_j)*[y_i > y_j]   (NDCG-style exponential gains, the usual RankNet / LambdaRank gain; the
 *          sorted label column then holds 2^y.  Labels whose gains coincide in float32 although y_i > y_j keep their pair
 *          with weight 0 -- integer labels never do.)
 *   LUT:   W = weight_lut[l_i][l_j]*[y_i > y_j] with the label LEVEL l = y + 1 of integer labels -1 .. 6 (binary clicks, graded
 *          relevance): ANY label-only label_pair_to_weight_func evaluated once on the 8 x 8 grid of levels (SURVEY 8b
 *          weight_lut, hard part 5).  The table's entries with l_i > l_j must be finite and > 0, so
 *          that the pair set stays [y_i > y_j] (C = W > 0, PW:193) and the counts stay position arithmetic; the other
 *          entries are ignored.  A label outside the menu or a bad entry fails the call on the device: loss = NaN,
 *          rn_last_device_error bit 8.
 *   LAMBDA: LambdaRank, W = |delta NDCG_ij| = (2^y_i - 2^y_j) * |D(r_i) - D(r_j)| / IDCG_g * [y_i > y_j]: D(r) = 1 / log2(1 + r),
 *          r_i the 1-based rank of row i by score among the rows of its group that can pair (descending, ties by row
 *          index), IDCG_g the DCG of the group's labels in descending order with gains 2^y - 1 (a group whose ideal DCG
 *          is not positive gets weight 0).  Not expressible as a label_pair_to_weight_func of the reference (it needs the
 *          scores): an extension on the same pair set, counts and normalisation; like every weight it is a constant
 *          of the step (PW:270).  rw_pos multiplies it.  Logistic loss, one GPU, not with only_wrong / rw_neg /
 *          deterministic / the focal term.
 * pairloss_func menu (pairwise_loss_from_batch.py:229, 274; the reference ships bpr_loss_func only, :96-127):
 *   LOGISTIC: l = softplus(-x), x = (s_i - s_j)*factor            (bpr_loss_func)
 *   HINGE:    l = max(0, margin - x)                               (margin ranking loss; d l / d x = -[margin - x > 0]) */
enum { RN_LOSS_LOGISTIC = 0, RN_LOSS_HINGE = 1 };

typedef struct rn_pairwise_args {
  int64_t B;               /* rows in the batch */
  int32_t K;               /* number of group-key columns (>= 1); keys AND together (PW:68-73) */
  int32_t label_func;      /* RN_LABEL_* */
  const int64_t* keys;     /* [K][B], key k of row i at keys[k*B + i]; keys[0] is the primary key (PW:286) */
  const float* logits;     /* [B] outputs */
  const float* labels;     /* [B] */
  const uint8_t* row_ok;   /* NULL or [B]; 0 = row forms no pair (sample mask PW:154-172 AND finite key) */
  const float* rw_pos;     /* NULL or [B] */
  const float* rw_neg;     /* NULL or [B] */
  float factor;            /* bpr_loss_func factor (PW:118-119) */
  float power;             /* click_occurance_power (PW:282-291); 0 disables */
  int32_t only_wrong;      /* only_use_wrong_order_pair (PW:197-203) */
  int32_t reduce_mean;     /* bpr_loss_func reduce_mean (PW:125-126) */
  /* Work partition for the multi-GPU "global in-batch" mode: this call scores tiles
   * [part_rank, part_count) of the pair space over the SAME (all-gathered) rows; loss and dlogits are then
   * partial sums to be added across ranks, n_pair / row_pairs are already global.  Single GPU: 0, 1. */
  int32_t part_rank;
  int32_t part_count;
  /* outputs */
  float* loss;             /* [1]  sum_P w*l / (float(n)+1e-10)   (or the plain sum if !reduce_mean); NaN if the call failed
                            *      on the device (a grid barrier timed out, the group table overflowed: the arena was
                            *      not clean) -- rn_last_device_error then tells which */
  float* n_pair_f32;       /* [1]  float32(n) as PW:276 returns it */
  int64_t* n_pair;         /* [1]  exact n */
  float* dlogits;          /* [B]  d loss / d logits */
  int64_t* row_pairs;      /* NULL or [B]: number of kept pairs with row i on the positive side */
  /* Blocked rows (multi-GPU global mode; 0 = contiguous columns).  After ONE all-gather of packed per-rank blocks
   * the B rows are B / block_rows blocks of block_rows rows, block_stride BYTES apart (a multiple of 16).  Every
   * input column pointer addresses its column inside block 0: element i lives at
   * column + (i / block_rows) * block_stride bytes + (i % block_rows) elements; key column k starts block_rows
   * int64 after key column k-1.  If out_chunk > block_rows, the outputs are laid out for ONE reduce-scatter:
   * d loss / d logits of row i goes to dlogits[(i / block_rows) * out_chunk + i % block_rows] and the (partial) loss
   * to dlogits[r * out_chunk + block_rows] of every block r (remaining pad floats are zeroed); dlogits then holds
   * (B / block_rows) * out_chunk floats.  row_pairs stays [B], indexed by row. */
  int64_t block_rows;
  int64_t block_stride;
  int64_t out_chunk;
  /* Peer-memory gather (global mode over NVLink peer mappings; gather_dst == NULL: off).  peer_blocks[r] is the packed
   * row block of rank r (this rank's own included, r < B / block_rows <= 8) in peer-mapped device memory.  The FIRST
   * kernel of the call copies the blocks with peer loads into the local blocked input buffer starting at gather_dst
   * (block r at gather_dst + r * block_stride; the column pointers above address that buffer) before anything reads
   * it: the all-gather happens inside the compute path, without a collective call.  The caller makes sure that all
   * ranks have written their blocks (a device-side barrier on the stream) before the call is enqueued. */
  const void* peer_blocks[8];
  void* gather_dst;
  /* Persistent scratch arena (0 = off: the arena may hold anything, the call starts with an initialisation kernel).
   * scratch_persistent != 0 promises that the arena was zeroed once (rn_pairwise_scratch_init) and has since been
   * written by nothing but rn_pairwise_fwd_bwd calls with the same K and the same scratch_rows: every call leaves the
   * regions the next one needs clean, so the initialisation kernel is dropped, and (one key column, contiguous rows,
   * part_count == 1) group segmentation takes the sort-free counting path.  After a device-side error
   * (rn_last_device_error != 0) zero the arena again.
   * scratch_rows: the row capacity the arena was sized for (rn_pairwise_scratch_bytes(scratch_rows, K)); 0 = B.  A
   * persistent arena keeps ONE layout, so batches of different sizes B <= scratch_rows may share it. */
  int32_t scratch_persistent;
  /* Deterministic mode (0 = off).  The default path accumulates d loss / d logits with floating-point atomics and places
   * groups in arrival order: results jitter from run to run at the 1e-7 level.  deterministic != 0 gives bit-identical
   * loss and dlogits for identical inputs (same device, same library): radix segmentation with first-occurrence group
   * ids, gradient sums in 64-bit fixed point (resolution: the largest pair weight x 2^-(61 - log2 2B)), loss partials
   * summed in a fixed order.  Slower (the sort), single GPU, not with only_wrong / rw_neg (RN_ERR_UNSUPPORTED). */
  int32_t deterministic;
  int64_t scratch_rows;
  /* Focal term fused into the same pass (0 = off): the call returns
   *   loss = pairwise loss + focal_weight * focal_crossentropy_loss(labels, logits, alpha, gamma, stop_weight_gradient)
   * (rec_block/focal_loss.py:12-66 of the reference, return_mean = True: the mean over ALL B rows, masked or not) and
   * dlogits holds the gradient of that sum -- the joint pointwise + pairwise objective of a ranking model in one call, the
   * focal part riding in the pair kernel's first and last pass.  focal_alpha / focal_gamma = 0 switch the respective
   * factor off (the reference's `if alpha:` / `if gamma:`).  Not with only_wrong / rw_neg / blocked rows. */
  float focal_weight, focal_alpha, focal_gamma;
  int32_t focal_stop_weight_gradient;
  /* Pair loss (RN_LOSS_*; 0 = the logistic bpr_loss_func) and the hinge margin (>= 0).  Everything else -- pair set,
   * weights, counts, occurrence weights, 1/n -- is shared.  HINGE: not with deterministic. */
  int32_t pair_loss;
  float margin;
  /* RN_LABEL_LUT: float32[8][8] in DEVICE memory, weight_lut[l_i * 8 + l_j] = W of a positive row at label level l_i and a
   * negative row at level l_j (level = label + 1).  NULL otherwise.  Single GPU, not with deterministic. */
  const float* weight_lut;
} rn_pairwise_args;

typedef struct rn_listwise_args {
  int64_t B;
  const int64_t* keys;     /* [B] canonical group ids */
  const uint8_t* row_ok;   /* NULL or [B]; 0 = id was NaN/inf (a singleton list, never valid) */
  const float* labels;     /* [B] */
  const float* logits;     /* [B] */
  const float* list_w;     /* NULL or [>= V]: per VALID list weights, first-occurrence order (LW:168-169) */
  float pos_neg_th;        /* LW:89 default 0.5; must be >= 0 for the segmented form */
  int32_t do_reduce;       /* LW:170-172 */
  /* outputs */
  float* loss;             /* [1] mean over valid lists, 0 if none (only written if do_reduce) */
  float* list_loss;        /* NULL or [B]: per-valid-list losses, first-occurrence order (first V entries) */
  int32_t* n_valid;        /* [1] V */
  int32_t* n_group;        /* [1] G (number of distinct ids) */
  float* dlogits;          /* [B] d mean-loss / d logits (d sum of list_loss if !do_reduce) */
  /* Persistent scratch arena, as rn_pairwise_args.scratch_persistent (zeroed once with rn_pairwise_scratch_init, then
   * written by nothing but rn_listwise_fwd_bwd calls with the same B): with do_reduce and neither list_w nor
   * list_loss the call is ONE kernel without a sort (per-list statistics accumulated on hash records, gradient
   * written in row order).  rn_listwise_dense needs the sorted form: call it after a non-persistent call only. */
  int32_t scratch_persistent;
  /* 1 / temperature of the softmax (0 = 1): the loss is taken on logits * inv_temperature, dlogits is the gradient with
   * respect to the UNSCALED logits.  Not in the reference's signature (a user divides the logits before the call);
   * here it saves that elementwise pass.  rn_listwise_dense returns the member logits unscaled. */
  float inv_temperature;
} rn_listwise_args;

int rn_version(void);
const char* rn_strerror(int code);

/* Float group ids -> canonical int64 keys; row_ok[i] = 0 where the id is NaN/+-inf.  and_into is a bit set: bit 0 =
 * AND the flag into the existing row_ok (used to merge the sample mask and several key columns); bit 1 = +-inf are
 * ordinary ids (the listwise path: tf.unique compares with ==, listwise_loss_from_batch.py:109, and inf == inf; the
 * pairwise path tests g_i - g_j == 0, pairwise_loss_from_batch.py:33-35, where inf - inf is NaN). */
int rn_canon_keys_f32(const float* ids, int64_t B, int64_t* keys_out, uint8_t* row_ok, int and_into, void* stream);
int rn_canon_keys_f64(const double* ids, int64_t B, int64_t* keys_out, uint8_t* row_ok, int and_into, void* stream);

/* ---- pairwise --------------------------------------------------------------------------------------- */
size_t rn_pairwise_scratch_bytes(int64_t B, int32_t K);
/* Zero a scratch arena (one cudaMemsetAsync on `stream`): required once before the first call that declares the arena
 * persistent (rn_pairwise_args.scratch_persistent), and again after anything else wrote into it. */
int rn_pairwise_scratch_init(void* scratch, size_t scratch_bytes, void* stream);
int rn_pairwise_fwd_bwd(const rn_pairwise_args* args, void* scratch, size_t scratch_bytes, void* stream);

/* Global mode (one process per GPU): pack this rank's row block for ONE all-gather -- [keys K x int64[B_loc]]
 * [logits f32][labels f32][rw_pos f32 if given][row_ok u8 if given], every column 16-byte aligned, zero padded to
 * `stride` bytes (a multiple of 16): the layout rn_pairwise_args.block_rows / block_stride reads in place.  B_loc must
 * be a multiple of 16.  One launch. */
int rn_pack_row_block(const int64_t* keys, int32_t K, const float* logits, const float* labels, const float* rw_pos,
                      const uint8_t* row_ok, int64_t B_loc, void* block_out, int64_t stride, void* stream);

/* Global mode, return path: dst[i] = sum over r < world of peer_out[r][my_rank * chunk + i], i < chunk -- the
 * reduce-scatter of the chunked gradient buffers (rn_pairwise_args.out_chunk) read straight from the peers' mapped
 * memory.  The caller synchronises the ranks (all calls finished) before it is enqueued.  One launch. */
int rn_reduce_peer_chunks(const void* const* peer_out, int32_t world, int32_t my_rank, int64_t chunk, float* dst,
                          void* stream);

/* ---- global in-batch mode: ONE call per step and rank --------------------------------------------------------
 * pairwise_loss over the concatenation of all ranks' rows (rank order = row order; SURVEY.md 8e -- the reference has
 * no multi-GPU mode, these are the semantics BASELINE.json's north_star defines).  One process per GPU.  Every rank
 * owns one symmetric buffer of rn_global_buffer_bytes(...) bytes, zeroed once, and mapped into every process of the
 * box (cudaIpc*, CUDA VMM, torch symmetric memory: the caller's plumbing); peer_buf[r] is rank r's buffer as mapped
 * into THIS process (peer_buf[rank] = the own one).  gather_buf: local device memory of rn_global_gather_bytes(...)
 * bytes (the gathered global rows).  `local` describes THIS rank's rows exactly as for rn_pairwise_fwd_bwd (B = rows per
 * rank, a multiple of 16 and the same on every rank; device pointers; dlogits[B]; loss / n_pair are the GLOBAL values,
 * identical on all ranks; the multi-GPU fields of `local` stay 0 / {0, 1}).  step: 0, 1, 2, ... -- the same on every
 * rank, incremented by the caller after every call on this set of buffers.  The call enqueues on `stream`: pack ->
 * device-side barrier over NVLink (flag words in the peers' buffers) -> the kernels (the first one gathers every
 * rank's packed rows with peer loads; every rank segments the same rows and scores its share of the pair space) ->
 * barrier -> one kernel that sums this rank's gradient chunk over the peers' buffers.  No collective library, no host
 * synchronisation.  Score- / weight-dependent pair sets (only_wrong, rw_neg): RN_ERR_UNSUPPORTED. */
typedef struct rn_global_args {
  rn_pairwise_args local;
  int32_t world, rank;
  void* peer_buf[8];
  void* gather_buf;
  int64_t step;
} rn_global_args;
size_t rn_global_buffer_bytes(int64_t B_loc, int32_t K, int32_t world, int32_t has_rw_pos, int32_t has_row_ok);
size_t rn_global_gather_bytes(int64_t B_loc, int32_t K, int32_t world, int32_t has_rw_pos, int32_t has_row_ok);
int rn_global_pairwise_fwd_bwd(const rn_global_args* args, void* scratch, size_t scratch_bytes, void* stream);

/* ---- pairwise, HOST buffers ------------------------------------------------------------------------------
 * Front end for callers whose tensors live in host memory (a CPU-placed TF2 op, a data loader): the same call as
 * rn_pairwise_fwd_bwd, but EVERY pointer of rn_pairwise_args (weight_lut included) is a HOST pointer (pinned memory, or
 * the copies are synchronous).  The object owns `depth` slots of device memory (input columns, outputs, scratch arena), a copy-in, a
 * compute and a copy-out stream; it allocates at create time only.  submit enqueues copy-in -> kernels -> copy-out of
 * one batch and returns a ticket without waiting (it blocks only while all `depth` slots are in flight); wait returns
 * when that batch's loss, n_pair_f32, n_pair, dlogits (and row_pairs) are in the host buffers given to submit.  With
 * depth >= 2 the copies of one batch overlap the kernels of its neighbours.  The multi-GPU fields of the args
 * (part_*, block_*, out_chunk, peer_blocks, gather_dst) must be 0 / {0, 1}.  One object per thread. */
typedef struct rn_host_pairwise rn_host_pairwise;
int rn_host_pairwise_create(int64_t B_max, int32_t K, int32_t depth, rn_host_pairwise** out);
int rn_host_pairwise_submit(rn_host_pairwise* p, const rn_pairwise_args* host_args, int32_t* ticket);
/* Opt-in (environment RN_HOST_STEP_GRAPH, read at create: 1 = always, 2 = once the object has found its eager submits host
 * bound, i.e. their mean host time above RN_HOST_STEP_GRAPH_US = 25): a slot that is handed the SAME host_args (the same
 * pinned buffers, sizes and options: a loader refilling its staging buffers) a second time captures its step -- copy-in,
 * kernels, copy-out -- as one CUDA graph and launches that from then on: three driver calls per step instead of fourteen.
 * Off by default: on a quiet host the pipeline is device bound and the eager path measured faster.
 * The buffers must stay page-locked at those addresses while the object lives.  Number of submits that went out that way: */
int64_t rn_host_pairwise_graph_steps(const rn_host_pairwise* p);
int rn_host_pairwise_wait(rn_host_pairwise* p, int32_t ticket);
int rn_host_pairwise_destroy(rn_host_pairwise* p);

/* Pair materialisation in the reference's row-major order (i ascending, then j ascending; PW:217).
 * Two phases because P is data dependent: _count enqueues the segmentation + counting, then synchronises
 * the stream once to return P on the host; _fill writes pos_idx/neg_idx (int32) and, if w != NULL, the
 * pair weight (without the occurrence factor).  label_cond = 0 lists every same-group ordered pair i != j
 * (the candidates an arbitrary label_pair_to_weight_func is then evaluated on). */
size_t rn_pair_indices_scratch_bytes(int64_t B, int32_t K);
int rn_pair_indices_count(const rn_pairwise_args* args, int32_t label_cond, void* scratch, size_t scratch_bytes,
                          int64_t* n_pairs_host, void* stream);
int rn_pair_indices_fill(const rn_pairwise_args* args, int32_t label_cond, void* scratch, size_t scratch_bytes,
                         int32_t* pos_idx, int32_t* neg_idx, float* w, int64_t capacity, void* stream);

/* occurance_power_weight (PW:130-151): out[i] = count(ids == ids[i]) ^ power, float32. */
size_t rn_occurrence_scratch_bytes(int64_t N);
int rn_occurrence_power_weight(const int64_t* ids, int64_t N, float power, float* out,
                               void* scratch, size_t scratch_bytes, void* stream);

/* ---- GAUC: group AUC on the same segmentation --------------------------------------------------------------
 * The metric the in-batch ranking losses are meant to move (the reference quotes its online GAUC uplift, README.md:5, 8,
 * and ships no implementation).  Per group (same keys; rows with row_ok != 0 and a non-NaN label):
 *   pairs_g = {(i, j) : y_i > y_j}  (the pair set of pairwise_loss, pairwise_loss_from_batch.py:189),
 *   AUC_g = (#{s_i > s_j} + 1/2 #{s_i == s_j}) / |pairs_g|;  GAUC = sum |g| AUC_g / sum |g| over groups with pairs.
 * Counts are exact integers.  scratch: rn_gauc_scratch_bytes; scratch_persistent as in rn_pairwise_args. */
typedef struct rn_gauc_args {
  int64_t B;
  int32_t K;
  int32_t scratch_persistent;
  const int64_t* keys;        /* [K][B] */
  const float* scores;        /* [B] */
  const float* labels;        /* [B] */
  const uint8_t* row_ok;      /* NULL or [B] */
  float* gauc;                /* [1] */
  float* auc_mean;            /* NULL or [1]: unweighted mean of AUC_g */
  int32_t* n_valid_groups;    /* [1] groups with at least one ordered label pair */
  int64_t* n_pair;            /* NULL or [1]: sum of |pairs_g| (= n_pair of pairwise_loss with default options) */
  int64_t* concordant2;       /* NULL or [1]: 2 x concordant + ties over all pairs */
} rn_gauc_args;
size_t rn_gauc_scratch_bytes(int64_t B, int32_t K);
int rn_gauc(const rn_gauc_args* args, void* scratch, size_t scratch_bytes, void* stream);

/* ---- segment pooling of slot embeddings (the step that feeds the model whose logits the losses score) ----------
 * Replaces rec_block/embedding_util.py:239-324 (embedding_using_sparse_batch_segment_ids; helper
 * sparse_batch_segment_ids_of_targets :127-198) for an embedding_func that is a table lookup:
 *   out[b][t][:] = sum over columns c with slots[b][c] == target_slots[t] of  weights[b][c] * table[ids[b][c]][:]
 * (method 'sum' = unsorted_segment_sum; mean != 0: 'mean' = unsorted_segment_mean, the sum divided by the number of such
 * columns, empty segments 0), accumulated in column order with the product rounded before the add, i.e. bit-identical
 * to TF's CPU kernels.  One launch, nothing materialised.  With an arbitrary embedding_func the caller passes
 * table = embedding_func(unique ids) and ids = the positions' indices into it (what tf.unique returns, :306-311).
 * An id outside [0, V) contributes nothing and sets bit 3 of *err_flag (a device word, NULL = do not report). */
typedef struct rn_pool_args {
  int64_t B, C;                 /* rows and columns of slots / ids / weights */
  int32_t T, D;                 /* target slots, embedding width */
  int32_t mean, reserved0;
  const int32_t* slots;         /* [B][C] */
  const int64_t* ids;           /* [B][C] row of `table` */
  const float* weights;         /* NULL or [B][C] */
  const int32_t* target_slots;  /* [T] (device), distinct */
  const float* table;           /* [V][D] */
  int64_t V;
} rn_pool_args;
int rn_segment_pool_fwd(const rn_pool_args* args, float* out /* [B][T][D] */, uint32_t* err_flag, void* stream);
/* Gradients of a scalar through out: d_table[V][D] += (accumulated with atomics: zero it first; NULL = skip) and
 * d_weights[B][C] += <table[id], d_out[b][t]> (NULL = skip; needs weights). */
int rn_segment_pool_bwd(const rn_pool_args* args, const float* d_out, float* d_table, float* d_weights, void* stream);

/* ---- listwise --------------------------------------------------------------------------------------- */
size_t rn_listwise_scratch_bytes(int64_t B);
int rn_listwise_fwd_bwd(const rn_listwise_args* args, void* scratch, size_t scratch_bytes, void* stream);
/* Compat layout of to_listwise_sample (LW:142-145): after rn_listwise_fwd_bwd on the same scratch, fill the
 * (V,B) dense_mask / dense_labels / dense_logits the reference returns.  V rows = *n_valid (read it first). */
int rn_listwise_dense(const rn_listwise_args* args, void* scratch, size_t scratch_bytes, int64_t V,
                      uint8_t* dense_mask, float* dense_labels, float* dense_logits,
                      int32_t do_mask_logits, float value_of_masked_logit, void* stream);

/* ---- measurement helpers ---------------------------------------------------------------------------- */
/* Issues iters*3 MUFU ops (ex2, lg2, rcp) per thread on a full grid; used by bench.py to MEASURE the SFU
 * peak the pair kernel is normalised against.  mufu_ops_out_host receives the op count issued. */
int rn_bench_mufu(int32_t iters, float* sink, int64_t* mufu_ops_out_host, void* stream);
/* Per-kernel timing of the dominant (pair) kernel, for the roofline line of bench.py: while enabled, every
 * rn_pairwise_fwd_bwd call brackets its pair-kernel launch with a cudaEvent pair on `stream` (up to max_calls
 * calls).  rn_profile_collect synchronises those events and returns the elapsed milliseconds per call.
 * Measurement aid only (process-global, not thread-safe, not for use under graph capture). */
int rn_profile_enable(int32_t max_calls);
/* in_graph != 0 (what rn_profile_enable does): the timed calls stay on the default path -- one launch of a cached CUDA
 * graph that also holds two event-record nodes around the pair kernel; each timed call then waits for the stream and
 * reads the elapsed time.  in_graph == 0: plain launches with stream events between the kernels (the figure then
 * includes the launch latency of a cooperative launch behind an event). */
int rn_profile_enable_ex(int32_t max_calls, int32_t in_graph);
int rn_profile_collect(float* ms_out_host, int32_t capacity, int32_t* n_out_host);
int rn_profile_disable(void);
/* Device-side status word of the last call that used `scratch` (0 = ok); reads it back (synchronises). */
int rn_last_device_error(void* scratch, int32_t* err_host, void* stream);
/* Phase timestamps (%globaltimer, ns) of the last call that used `scratch`, written by CTA 0 of the
 * segmentation kernel (slots 0-19) and the pair kernel (20-23); reads them back (synchronises). */
int rn_debug_timestamps(void* scratch, uint64_t* ts_host, int32_t capacity, void* stream);
/* Stage timing breakdown helper: number of kernel launches the last-built pipeline enqueues per call. */
int rn_pairwise_launch_count(int64_t B, int32_t K);
int rn_listwise_launch_count(int64_t B);
/* Calls of the calling thread that were enqueued as ONE launch of a cached CUDA graph (rn_pairwise_fwd_bwd does that
 * unless RN_GRAPH=0, the stream is under capture, or rn_profile_enable is active); negative (-count - 1) after a graph
 * API failure switched the thread back to plain launches. */
int64_t rn_debug_graph_launches(void);
/* Byte offset of a region of the pairwise arena of (B, K) for the developer tools under scripts/ (which: 0 per-CTA
 * debug stamps, 1 group records, 2 J ranges, 3 piece boundaries); -1 if unknown. */
int64_t rn_debug_arena_offset(int64_t B, int32_t K, int32_t which);

#ifdef __cplusplus
}
#endif
#endif /* RECNOW_B200_H_ */
