/*
 * rechub_b200.h — C ABI of the B200 (sm_100a) embedding-and-interaction engine.
 *
 * This is the drop-in boundary for the hot path of datawhalechina/torch-rechub
 * (SURVEY.md §8b).  The reference has no FFI of its own — its boundary is the Python class
 * surface — so every entry point below names the reference function whose arithmetic it
 * replaces (paths relative to the reference root, torch-rechub v0.8.0 @ 5a73b2e).
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types cross this boundary;
 *   - every pointer marked "device" is a CUDA device pointer on the current device;
 *     the library never allocates, frees or synchronises — the caller owns all buffers;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - tables and their gradient buffers are fp32, row-major (vocab, dim);
 *   - return value: 0 = RH_OK, otherwise an rh_status; rh_last_error() returns a
 *     thread-local human-readable message for the last failing call;
 *   - out-of-range ids never touch memory: the row reads as zeros / the scatter is dropped and
 *     `*err_flag` (device int32, may be NULL) is set to 1 + the field index.  The host
 *     binding turns a non-zero flag into the reference's `IndexError: index out of range in self`.
 */
#ifndef RECHUB_B200_H_
#define RECHUB_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RH_ABI_VERSION 1

typedef enum rh_status {
  RH_OK = 0,
  RH_ERR_INVALID_ARG = 1,   /* bad shape / alignment / NULL where data is required        */
  RH_ERR_UNSUPPORTED = 2,   /* configuration outside what the kernels implement           */
  RH_ERR_CUDA = 3           /* a CUDA runtime call failed (message holds cudaGetErrorString) */
} rh_status;

#define RH_ERRFLAG_SYNC_TIMEOUT 0x7ffffff0 /* err_flag value: a cross-GPU hand-over (rh_sync) saw no signal for ~2^26 polls (tens of seconds) */
#define RH_MAX_FIELDS 64    /* id columns per rh_fields_* launch (callers chunk beyond it)  */
#define RH_MAX_DENSE  32    /* numeric columns per rh_fields_fwd launch                     */

/* One id column looked up in one table: a SparseFeature of the reference
 * (basic/features.py:42-71; lookup at basic/layers.py:83,85). */
typedef struct rh_field {
  const float* table;       /* device (vocab, dim) fp32                                          */
  float*       table_grad;  /* device (vocab, dim) fp32, backward only; NULL = frozen table      */
  const void*  ids;         /* device ids of this column, int64 (or int32 when ids_are_i32)       */
  int64_t      id_stride;   /* elements between consecutive samples (1 for a (B,) column)         */
  int32_t      ids_are_i32;
  int32_t      vocab;       /* rows of the table (bounds check)                                   */
  int32_t      padding_idx; /* row that receives no gradient (nn.Embedding padding_idx); -1 none  */
  int32_t      tile_col;    /* first column of this field in the flattened tile; -1 = not emitted */
  int32_t      fm_slot;     /* position among the FM/LR fields (LR weight offset = fm_slot*dim);
                               -1 = the field takes no part in FM/LR                              */
} rh_field;

/* One numeric column copied into the tile: a DenseFeature (basic/features.py:74-87;
 * `.float()` + concat at basic/layers.py:101,105,120). */
typedef struct rh_dense {
  const void* values;       /* device (B, width)                                                  */
  int64_t     stride;       /* elements between consecutive samples                               */
  int32_t     dtype;        /* 0 = fp32, 1 = fp64 (pandas default), 2 = int64, 3 = int32          */
  int32_t     width;        /* values per sample (DenseFeature.embed_dim)                         */
  int32_t     tile_col;     /* first destination column                                            */
} rh_dense;

int         rh_abi_version(void);
const char* rh_last_error(void);
/* Number of kernels this library has launched so far in this process (bench.py's gpu_launches). */
unsigned long long rh_launch_count(void);
/* Programmatic dependent launch for the hot-path kernels.  `mask` selects the kernel families that are LAUNCHED with the attribute
 * (their CTAs may be scheduled as soon as every CTA of the preceding kernel in the stream has issued its trigger, run their
 * memory-free prologue and block in griddepcontrol.wait until that kernel has completed and flushed):
 *   1 GEMM   2 fused BatchNorm   4 fused gather / scatter   8 optimiser updates   16 loss   32 batch copy   64 multi-GPU hand-overs
 * Every hot-path kernel issues the trigger at its top whatever the mask.  mask < 0 only queries.  Returns the previous mask. */
int rh_set_pdl(int mask);
/* Preferred shared-memory carveout (0..100 %, -1 = the driver's choice) applied to every kernel of the library at its next first launch;
 * returns the previous setting.  An experiment switch (does a uniform carveout avoid L1 / shared-memory reconfigurations between the
 * 198 KB GEMM CTAs and their neighbours?). */
int rh_set_smem_carveout(int percent);
/* L2 fetch granularity of the current device (cudaLimitMaxL2FetchGranularity): bytes L2 pulls from DRAM per sector miss.
 * set_bytes > 0 sets it first (32 / 64 / 128); returns the value in force, < 0 on a CUDA error.  Random 64-byte embedding rows
 * (basic/layers.py:83,85 lookups) cost twice their DRAM bytes at 128 — the engine's host side lowers it once per device. */
int rh_l2_fetch_granularity(int set_bytes);

/* ---------------------------------------------------------------------------------------------
 * Fused multi-field gather (+ FM second-order term + LR first-order term + flattened tile).
 *
 * Replaces, in ONE launch: EmbeddingLayer.forward's per-field index_select + unsqueeze + cat
 * (basic/layers.py:77-127), FM.forward (basic/layers.py:313-319), LR.forward on the flattened
 * embeddings (basic/layers.py:183-189 as called from models/ranking/deepfm.py:39), and the
 * dense-value concat (basic/layers.py:101-120).
 *
 *   dim        row width shared by all `fields` of this launch (callers group by dim)
 *   tile       device (batch, tile_ld) fp32 or NULL; field f fills columns
 *              [tile_col, tile_col+dim), dense column j fills [tile_col, tile_col+width)
 *   lr_weight  device (n_fm*dim) fp32, lr_bias device (1) fp32      (NULL when y_lr is NULL)
 *   y_fm       device (batch) fp32 or NULL: 0.5 * sum_d[(sum_f e)^2 - sum_f e^2]
 *   y_lr       device (batch) fp32 or NULL: <flatten(e), lr_weight> + lr_bias
 *   field_sum  device (batch, dim) fp32 or NULL: sum_f e over FM fields (saved for backward)
 * ------------------------------------------------------------------------------------------- */
int rh_fields_fwd(const rh_field* fields, int n_fields, int dim,
                  const rh_dense* dense, int n_dense,
                  int batch,
                  float* tile, int64_t tile_ld,
                  const float* lr_weight, const float* lr_bias,
                  float* y_fm, float* y_lr, float* field_sum,
                  int32_t* err_flag, void* stream);

/* Fused gather + all-to-all over peer memory (multi-GPU field sharding): the owner of the tables gathers
 * its fields for the GLOBAL batch and writes each sample's rows STRAIGHT into the tile of the GPU that holds that sample
 * (NVLink peer stores from inside the gather kernel) — no local staging buffer, no NCCL all-to-all.
 *   dest_tiles[d]  device pointer (peer-mapped) of destination d's tile block; sample i goes to dest i / rows_per_dest,
 *                  row i % rows_per_dest, row stride tile_ld; field f fills columns [tile_col, tile_col + dim)
 * Callers order the exchange with cross-GPU barriers.  The backward direction needs no new entry point: rh_fields_bwd's
 * rh_field.table_grad may be a peer pointer (vector RED over NVLink). */
int rh_fields_fwd_p2p(const rh_field* fields, int n_fields, int dim, int batch,
                      float* const* dest_tiles, int n_dest, int rows_per_dest, int64_t tile_ld,
                      int32_t* err_flag, void* stream);

/* ids of my samples -> the owners' id buffers (peer memory).  cols[c] (ids / id_stride / ids_are_i32 are read) is the c-th id
 * column, col_dest[c] its destination GPU (columns sorted by destination; the order inside a destination is the slot order);
 * dest_base[d] points at MY block of destination d's id buffer:  dest_base[d][slot * slot_stride + j * sample_stride] = ids[j]
 * (field-major buffers, sample_stride 1, give contiguous NVLink stores and unit-stride ids for the owner's gather;
 * sample-major is slot_stride 1, sample_stride fmax).  At most fmax columns per destination.
 * rh_field.table_grad may equally be a peer pointer in rh_fields_bwd: the sample's GPU then REDs row gradients straight
 * into the owner's gradient buffer at the row id (no staging buffer, no owner-side scatter pass). */
int rh_ids_scatter(const rh_field* cols, int n_cols, const int32_t* col_dest, int batch,
                   int64_t* const* dest_base, int n_dest, int fmax, int64_t slot_stride, int64_t sample_stride,
                   void* stream);

/* The exchange's cross-GPU hand-overs folded INTO the launches (no barrier kernels between them).  Every rank keeps, per phase, a
 * flags array of 8 int32 in peer-mapped memory: flags[s] = the last step for which rank s's data of that phase has landed here.
 *   producer side  sig_flags[d] = destination d's flags array of the phase this launch PRODUCES (d < sig_world); when the grid's last
 *                  CTA has finished, the step number goes to sig_flags[d][sig_rank] for every d (one system fence, st.release.sys);
 *                  ticket: one zeroed unsigned per launch site, left zeroed.
 *   consumer side  wait_flags = MY flags array of the phase this launch CONSUMES: every CTA waits (ld.acquire.sys) until
 *                  wait_flags[s] >= *step for each s in wait_mask before it reads peer-written memory.
 *   step           device int32, advanced once per exchange by rh_ids_scatter_signal (which publishes *step + 1).
 *   id_snapshot_delta (bytes, owner-side gather): every id read at address a is also stored at a + delta — the local copy the
 *                  owner's backward / optimiser use while the peers may already be refilling the id buffer.
 * Step numbers only grow: no flag is ever reset.  All pointers NULL / masks 0 = no hand-over on that side. */
typedef struct rh_sync {
  const int32_t*  wait_flags;
  const int32_t*  step;
  int32_t* const* sig_flags;
  int32_t*        ticket;
  int32_t         wait_mask;
  int32_t         sig_world;
  int32_t         sig_rank;
  int32_t         reserved;
  int64_t         id_snapshot_delta;
} rh_sync;

/* rh_fields_fwd (dest_tiles == NULL) or rh_fields_fwd_p2p (dest_tiles != NULL: dense / lr / fm arguments must be NULL / 0) with the
 * hand-overs of `sync` inside the launch. */
int rh_fields_fwd_sync(const rh_field* fields, int n_fields, int dim,
                       const rh_dense* dense, int n_dense, int batch,
                       float* tile, int64_t tile_ld,
                       const float* lr_weight, const float* lr_bias,
                       float* y_fm, float* y_lr, float* field_sum,
                       float* const* dest_tiles, int n_dest, int rows_per_dest,
                       const rh_sync* sync, int32_t* err_flag, void* stream);

/* rh_ids_scatter that advances the exchange's step counter and publishes it: when the last CTA has stored its ids, *step_dev + 1 goes
 * to peer_flags[d][rank] for every d < world (the owners' id-phase flags) and into *step_dev. */
int rh_ids_scatter_signal(const rh_field* cols, int n_cols, const int32_t* col_dest, int batch,
                          int64_t* const* dest_base, int n_dest, int fmax, int64_t slot_stride, int64_t sample_stride,
                          int32_t* const* peer_flags, int rank, int world, int32_t* step_dev, int32_t* ticket_dev,
                          void* stream);

/* Backward of rh_fields_fwd: the sparse-gradient scatter-add into the tables' gradient buffers.
 *
 * Replaces autograd's aten::embedding_dense_backward per lookup (reference: implicit in
 * nn.Embedding(sparse=False), basic/initializers.py:17) plus the FM / LR / cat backward
 * elementwise chain.  For row e = table[id] of FM field f:
 *     g = d_tile[b, cols(f)] + d_y_fm[b] * (field_sum[b] - e) + d_y_lr[b] * lr_weight[f]
 * and table_grad[id] += g (vector RED, duplicates accumulate; padding_idx rows are skipped).
 *
 *   tile        the forward tile (rows are re-read from it when the field was emitted) or NULL
 *               (rows are re-gathered from the table)
 *   d_tile      device (batch, d_tile_ld) fp32 or NULL (any row stride; 16-byte aligned rows go faster)
 *   d_y_fm, d_y_lr   device (batch) fp32 or NULL
 *   d_lr_weight device (n_fm*dim) fp32, ACCUMULATED into (caller zeroes) or NULL
 *   d_lr_bias   device (1) fp32, accumulated into, or NULL
 */
int rh_fields_bwd(const rh_field* fields, int n_fields, int dim, int batch,
                  const float* tile, int64_t tile_ld,
                  const float* d_tile, int64_t d_tile_ld,
                  const float* d_y_fm, const float* d_y_lr,
                  const float* lr_weight, const float* field_sum,
                  float* d_lr_weight, float* d_lr_bias,
                  int32_t* err_flag, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Single-table lookups of arbitrary id shape (FieldTable.forward; SequenceFeature paths).
 * ------------------------------------------------------------------------------------------- */

/* out[i, :] = table[ids[i], :] for i < n  — nn.Embedding.forward as used at basic/layers.py:83-99
 * (sequence features with pooling="concat": basic/layers.py:91-92,204-205). */
int rh_rows_gather(const float* table, int vocab, int dim,
                   const void* ids, int ids_are_i32, int64_t n,
                   float* out, int32_t* err_flag, void* stream);

/* table_grad[ids[i], :] += d_out[i, :], skipping padding_idx (embedding_dense_backward). */
int rh_rows_scatter_add(float* table_grad, int vocab, int dim, int padding_idx,
                        const void* ids, int ids_are_i32, int64_t n,
                        const float* d_out, int32_t* err_flag, void* stream);

/* Masked sum / mean pooling of a padded id sequence: SumPooling / AveragePooling with the
 * InputMask rule (basic/layers.py:148-161, 209-251): position l counts when
 * ids[b,l] != mask_id (mask_id = padding_idx, or -1 when the feature has none);
 * mean divides by (count + 1e-16).   mode: 1 = sum, 2 = mean.   out: (batch, dim) with row stride
 * out_ld (so the pooled vector can land directly in a tile column block). */
int rh_seq_pool_fwd(const float* table, int vocab, int dim,
                    const void* ids, int ids_are_i32, int batch, int seq_len,
                    int mode, int64_t mask_id,
                    float* out, int64_t out_ld, int32_t* err_flag, void* stream);

int rh_seq_pool_bwd(float* table_grad, int vocab, int dim, int padding_idx,
                    const void* ids, int ids_are_i32, int batch, int seq_len,
                    int mode, int64_t mask_id,
                    const float* d_out, int64_t d_out_ld, int32_t* err_flag, void* stream);

/* table_grad[ids[i], :] = 0 — sparse replacement for zero-filling a dense (vocab, dim) gradient
 * (the reference spends 66 % of its CPU step in aten::fill_). */
int rh_rows_zero(float* table_grad, int vocab, int dim,
                 const void* ids, int ids_are_i32, int64_t n, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Row-wise (lazy) optimisers over the rows touched by one batch.
 * For every DISTINCT id in `ids` (first claimant wins through `stamp`, a device int32[vocab]
 * holding the last step that updated the row):
 *     g = table_grad[id]; table_grad[id] = 0;   then the update below on row `id` only.
 *   kind 0 SGD      w -= lr * (g + wd*w)
 *   kind 1 Adam     torch.optim.Adam arithmetic (L2-style weight_decay, bias correction with the
 *                   row-independent global step) applied to touched rows only
 *   kind 2 Adagrad  state1 += g^2 ; w -= lr * g / (sqrt(state1) + eps)       (g includes wd*w)
 * state1/state2: device (vocab, dim) fp32 moment buffers (Adam: m, v; Adagrad: state1 = sum).
 * stamp: device int32[vocab], zero-initialised.  step_dev / lr_dev: DEVICE scalars (int32 / fp32) so
 * that a captured CUDA graph sees the step counter and a scheduler-updated learning rate on replay;
 * rh_opt_advance increments *step_dev and refreshes the Adam bias corrections (call it once per optimiser step,
 * before the updates).
 * ------------------------------------------------------------------------------------------- */
int rh_rowwise_update(float* table, float* table_grad, float* state1, float* state2,
                      int32_t* stamp, int vocab, int dim,
                      const void* ids, int ids_are_i32, int64_t n,
                      int kind, const int32_t* step_dev, const float* lr_dev, const float* bias_corr_dev,
                      int64_t state_row_stride,
                      float beta1, float beta2, float eps, float weight_decay, void* stream);
/* ++*step_dev, and (Adam) bias_corr_dev[0] = 1 - beta1^step, bias_corr_dev[1] = sqrt(1 - beta2^step) in fp64, once
 * per optimiser step instead of once per thread.  bias_corr_dev: device float[2], may be NULL for SGD/Adagrad. */
int rh_opt_advance(int32_t* step_dev, float* bias_corr_dev, float beta1, float beta2, void* stream);

/* The same update for ALL tables of one batch in a single launch (grid.y = field): fields[i] gives
 * table_grad / ids / id_stride / vocab of table i; tables / state1 / state2 / stamp are host arrays
 * of n_fields device pointers.  Needs dim % 4 == 0.  state_row_stride: floats between consecutive rows of state1 / state2
 * (2*dim when m and v are interleaved as one 128-byte record per row — one DRAM burst instead of two random ones).  This is the optimiser half of the
 * "backward sparse-grad scatter-add into the tables" (BASELINE.json north_star). */
int rh_fields_rowwise_update(const rh_field* fields, int n_fields, int dim, int batch,
                             float* const* tables, float* const* state1, float* const* state2,
                             int32_t* const* stamp, int kind,
                             const int32_t* step_dev, const float* lr_dev, const float* bias_corr_dev,
                             int64_t state_row_stride,
                             float beta1, float beta2, float eps, float weight_decay, void* stream);

/* L2 prefetch (prefetch.global.L2, nothing loaded or stored) of the rows the ids select in up to three arrays per field:
 * rh_field.table (vocab, dim), rh_field.table_grad (vocab, dim; may be NULL) and state[f] (vocab, state_row_stride floats; the
 * interleaved optimiser records; `state` may be NULL).  Meant for the NEXT batch, one step ahead on a copy stream: the
 * reference has no counterpart (its lookups are synchronous index_select calls, basic/layers.py:83). */
int rh_fields_prefetch(const rh_field* fields, int n_fields, int dim, int batch,
                       float* const* state, int64_t state_row_stride, void* stream);

/* table_grad[ids] = 0 for all tables of one batch in a single launch (sparse zero_grad). */
int rh_fields_zero(const rh_field* fields, int n_fields, int dim, int batch, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Interactions on dense tiles.
 * ------------------------------------------------------------------------------------------- */

/* FM.forward on a materialised (batch, n_fields, dim) tensor (basic/layers.py:313-319).
 * reduce_sum != 0: y (batch);  reduce_sum == 0: y (batch, dim). */
int rh_fm_fwd(const float* x, int batch, int n_fields, int dim, int reduce_sum,
              float* y, void* stream);
int rh_fm_bwd(const float* x, const float* d_y, int batch, int n_fields, int dim, int reduce_sum,
              float* d_x, void* stream);

/* FM + LR on an already materialised tile (batch, n_fields*dim) — the sharded multi-GPU front end, where rows arrive
 * over NVLink instead of being gathered locally.  Same arithmetic as the FM/LR part of rh_fields_fwd / rh_fields_bwd:
 *   y_fm = FM(tile) (basic/layers.py:313-319), y_lr = tile . lr_weight + lr_bias (layers.py:183-189), field_sum = sum_f e.
 * Backward: d_tile[b, f] (= or +=, `accumulate`) d_y_fm[b] (field_sum[b] - e) + d_y_lr[b] lr_weight[f];
 * d_lr_weight / d_lr_bias are accumulated into (caller zeroes). */
int rh_tile_fm_lr_fwd(const float* tile, int64_t tile_ld, int batch, int n_fields, int dim,
                      const float* lr_weight, const float* lr_bias,
                      float* y_fm, float* y_lr, float* field_sum, void* stream);
int rh_tile_fm_lr_bwd(const float* tile, int64_t tile_ld, int batch, int n_fields, int dim,
                      const float* lr_weight, const float* field_sum,
                      const float* d_y_fm, const float* d_y_lr,
                      float* d_tile, int64_t d_tile_ld, int accumulate,
                      float* d_lr_weight, float* d_lr_bias, void* stream);

/* CrossNetwork.forward (basic/layers.py:412-420): x_{l+1} = x0 * <w_l, x_l> + b_l + x_l for
 * l < n_layers (<= 16), all layers in one launch, the row held in registers throughout.
 *   w, b       host arrays of n_layers DEVICE pointers, each (width) fp32 — the reference keeps one
 *              Linear(width,1,bias=False).weight and one bias Parameter per layer (layers.py:409-410)
 *   xw_saved   device (n_layers, batch) fp32: the per-layer scalars <w_l, x_l> (kept for backward) */
int rh_cross_fwd(const float* x0, int64_t x_ld, int batch, int width, int n_layers,
                 const float* const* w, const float* const* b,
                 float* out, int64_t out_ld, float* xw_saved, void* stream);

/* Backward of rh_cross_fwd.  d_w, d_b: host arrays of n_layers device pointers, each (width) fp32,
 * accumulated into (caller zeroes). */
int rh_cross_bwd(const float* x0, int64_t x_ld, int batch, int width, int n_layers,
                 const float* const* w, const float* const* b, const float* xw_saved,
                 const float* d_out, int64_t d_out_ld,
                 float* d_x0, int64_t d_x0_ld, float* const* d_w, float* const* d_b, void* stream);

/* ---------------------------------------------------------------------------------------------
 * MLP glue: BatchNorm1d (train / eval) + activation + dropout in one pass
 * (basic/layers.py:276-292: Linear -> BatchNorm1d -> activation -> Dropout).
 * ------------------------------------------------------------------------------------------- */

/* Column statistics of h (rows, cols) for BatchNorm1d's training forward.
 *   stats    device (2*cols + 1) fp32: [0,cols) mean, [cols,2*cols) biased variance, [2*cols] the bit pattern of this
 *            forward's step counter (num_batches_tracked after the increment) = the dropout stream id of the layer
 *   scratch  device (2*cols + 1) fp32, zeroed ONCE by the caller; the kernel leaves it zeroed
 * running_mean/var (may be NULL) are updated with `momentum` (unbiased variance, as torch does);
 * num_batches_tracked (device int64, may be NULL) is incremented. */
int rh_colstats(const float* h, int64_t h_ld, int64_t rows, int cols,
                float* stats, float* scratch,
                float* running_mean, float* running_var, int64_t* num_batches_tracked,
                float momentum, void* stream);

/* y = dropout(act(gamma * (h - mean) / sqrt(var + eps) + beta)).   mean/var NULL: no normalisation.
 *   act: 0 identity, 1 ReLU, 2 Dice (basic/activation.py:15-25: per-ROW statistics over `cols`;
 *        act_param = device alpha (1)), 3 PReLU (act_param = single slope), 4 sigmoid, 5 LeakyReLU(0.01)
 *   dropout (p_drop > 0): the keep decision is a pure function of (dropout_seed, *dropout_counter, element index),
 *        regenerated in the backward pass — no mask is stored.  dropout_counter: device fp32 holding an int bit pattern
 *        (stats[2*cols] of rh_colstats) or NULL.  Kept values are scaled by 1/(1-p_drop). */
int rh_bn_act_fwd(const float* h, int64_t h_ld, int64_t rows, int cols,
                  const float* mean, const float* var, float bn_eps,
                  const float* gamma, const float* beta,
                  int act, const float* act_param, float dice_eps,
                  float p_drop, uint32_t dropout_seed, const float* dropout_counter,
                  float* y, int64_t y_ld, void* stream);

/* Backward of rh_bn_act_fwd through dropout, activation and batch norm; the pre-activation is
 * recomputed from h.  d_h: gradient w.r.t. the Linear output h.
 *   training != 0: batch statistics take part in the gradient (two launches: column sums, apply);
 *                  d_gamma/d_beta (cols) double as the column-sum buffers and MUST be zero on entry.
 *   training == 0: BN is affine (running statistics); d_gamma/d_beta are accumulated into if non-NULL.
 *   d_act_param (1): accumulated into (Dice alpha / PReLU slope), may be NULL. */
int rh_bn_act_bwd(const float* h, int64_t h_ld, int64_t rows, int cols,
                  const float* mean, const float* var, float bn_eps,
                  const float* gamma, const float* beta,
                  int act, const float* act_param, float dice_eps,
                  float p_drop, uint32_t dropout_seed, const float* dropout_counter,
                  const float* d_y, int64_t d_y_ld, int training,
                  float* d_h, int64_t d_h_ld,
                  float* d_gamma, float* d_beta, float* d_act_param, void* stream);

/* Training-mode BatchNorm1d + activation + dropout of one tower layer in ONE launch (rh_colstats + rh_bn_act_fwd fused): every
 * CTA keeps its rows of h in registers across a grid-wide barrier (column sums published by atomics, then the apply).
 * Replaces [BatchNorm1d -> activation -> Dropout] of MLP.forward (basic/layers.py:282-285), batch statistics included.
 *   stats (2 cols + 1): mean | biased variance | step-counter bits (out; the backward's input, same layout as rh_colstats)
 *   scratch: rh_bn_fused_scratch_floats(cols) floats, zeroed ONCE by the caller; the kernel leaves it zeroed
 *   running_mean / running_var / num_batches_tracked: updated as nn.BatchNorm1d does (may be NULL)
 *   head mode (head_w != NULL): the layer is the LAST hidden layer and the tower's output layer Linear(cols, 1) + side terms +
 *   sigmoid (rh_head_fwd's arithmetic) consumes y in registers: head_out[r] = f(<y[r], head_w> + head_b + extra0[r] + extra1[r]);
 *   y may then be NULL (the activation never reaches HBM).
 * Shapes: rh_bn_fused_supported(rows, cols, head) != 0 (cols % 4 == 0, cols <= 512 (256 in head mode), rows <= SMs * 64 / ceil(cols/128)). */
int64_t rh_bn_fused_scratch_floats(int cols);
int rh_bn_fused_supported(int64_t rows, int cols, int head);
int rh_bn_act_fused_fwd(const float* h, int64_t h_ld, int64_t rows, int cols, float bn_eps,
                        const float* gamma, const float* beta, int act, const float* act_param, float dice_eps,
                        float p_drop, uint32_t dropout_seed,
                        float* running_mean, float* running_var, int64_t* num_batches_tracked, float momentum,
                        float* stats, float* scratch, float* y, int64_t y_ld,
                        const float* head_w, const float* head_b, const float* extra0, const float* extra1,
                        int apply_sigmoid, float* head_out, void* stream);

/* Backward of rh_bn_act_fused_fwd in ONE launch (both passes of rh_bn_act_bwd; in head mode also rh_head_bwd): d_h is final,
 * d_gamma / d_beta / d_act_param (/ d_head_w (cols) / d_head_b (1)) are WRITTEN (no zeroing by the caller), d_extra (rows) or NULL
 * receives the gradient of every side term, d_lin_bias (cols) or NULL is written with the gradient of the Linear bias in front of
 * the BatchNorm (exactly 0: batch statistics remove any per-column shift).  Plain mode reads d_y (rows, cols); head mode reads d_head_out (rows) and the saved
 * head_out (sigmoid derivative). */
int rh_bn_act_fused_bwd(const float* h, int64_t h_ld, int64_t rows, int cols, const float* stats, float bn_eps,
                        const float* gamma, const float* beta, int act, const float* act_param, float dice_eps,
                        float p_drop, uint32_t dropout_seed,
                        const float* d_y, int64_t d_y_ld,
                        const float* head_w, const float* head_out, const float* d_head_out, int apply_sigmoid,
                        float* scratch, float* d_h, int64_t d_h_ld,
                        float* d_gamma, float* d_beta, float* d_act_param,
                        float* d_head_w, float* d_head_b, float* d_extra, float* d_lin_bias, void* stream);

/* Output head of a ranking tower in one pass: the MLP's output layer nn.Linear(k, 1) (reference basic/layers.py:279-280) fused
 * with the model's tail  sigmoid(y_deep + extra0 + extra1)  — DeepFM's  y_linear + y_fm + y_deep  (models/ranking/deepfm.py:41-43).
 *   x (rows, k) row stride x_ld;  w (k);  bias (1) or NULL;  extra0/extra1 (rows) or NULL;  out (rows)
 *   out[r] = f(<x[r], w> + bias + extra0[r] + extra1[r]),  f = sigmoid if apply_sigmoid else identity.   k <= 1024. */
int rh_head_fwd(const float* x, int64_t x_ld, int64_t rows, int k, const float* w, const float* bias,
                const float* extra0, const float* extra1, int apply_sigmoid, float* out, void* stream);

/* Backward of rh_head_fwd.  d_logit[r] = d_out[r] * out[r] * (1 - out[r]) (or d_out[r] without the sigmoid);
 *   d_x[r, :] = d_logit[r] * w (or NULL);  d_w (k) += sum_r d_logit[r] * x[r, :];  d_b (1) += sum_r d_logit[r]  (both must
 *   arrive zeroed, either may be NULL);  d_extra (rows) = d_logit (the gradient of every extra term), or NULL. */
int rh_head_bwd(const float* x, int64_t x_ld, int64_t rows, int k, const float* w, const float* out,
                const float* d_out, int apply_sigmoid, float* d_x, int64_t d_x_ld, float* d_w, float* d_b,
                float* d_extra, void* stream);

/* One launch copying n (1..4) contiguous device buffers: a packed batch (id block, numeric block, sequence block, labels) into the
 * captured step's static buffers (the reference moves one tensor per column, trainers/ctr_trainer.py:84).  dst/src/bytes: HOST arrays
 * of device pointers / byte counts; buffers must not overlap. */
int rh_copy_segments(int n, void* const* dst, const void* const* src, const int64_t* bytes, void* stream);

/* torch.nn.BCELoss(reduction="mean") on probabilities (the CTR trainer's criterion, trainers/ctr_trainer.py:68,88) in one launch each way:
 *   *loss = mean_i -(y_i max(log p_i, -100) + (1 - y_i) max(log(1 - p_i), -100));   d_prob_i = *d_loss * (p_i - y_i) / max(p_i (1 - p_i), 1e-12) / n
 * scratch: 65 floats, zeroed once by the caller (block partials + a ticket the kernel leaves zero); sums run in a fixed order. */
int rh_bce_fwd(const float* prob, const float* target, int64_t n, float* scratch, float* loss, void* stream);
int rh_bce_bwd(const float* prob, const float* target, const float* d_loss, int64_t n, float* d_prob, void* stream);

/* One launch of SGD / Adam / Adagrad (kinds as rh_rowwise_update) over n_tensors small dense tensors — the tower's
 * weights (the dense half of optimizer.step(), trainers/ctr_trainer.py:99).  params/grads/state1/state2: host arrays of
 * device pointers; numel: host array.  lr_dev / bias_corr_dev: the device scalars of rh_opt_advance. */
int rh_dense_update(int n_tensors, float* const* params, const float* const* grads,
                    float* const* state1, float* const* state2, const int64_t* numel,
                    int kind, const float* lr_dev, const float* bias_corr_dev,
                    float beta1, float beta2, float eps, float weight_decay, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Multi-GPU dense step over NVLink peer memory (replaces nn.DataParallel's gradient reduction + optimizer.step() for the
 * replicated parameters, trainers/ctr_trainer.py:53-55,97-99): all-reduce of the tower gradients fused with the optimiser update.
 *   staging buffer of a rank: 2 slots of rh_dense_stage_floats(n_tensors, numel) floats (peer-mapped, double-buffered by step parity)
 *   flags of a rank: 8 int32 (peer-mapped, zero-initialised): flags[s] = last step rank s has published
 *   epoch_dev: device int32, number of completed reductions (zero-initialised; advanced by rh_dense_reduce_update)
 *   ticket_dev: device uint32, zero-initialised, left zero
 * ------------------------------------------------------------------------------------------- */
int64_t rh_dense_stage_floats(int n_tensors, const int64_t* numel);
/* Cross-GPU barrier of the sharded exchange (the step's ids / rows / gradient hand-overs): publish *epoch_dev + 1 into this rank's slot
 * of every rank's flags (8 int32, peer-mapped, zero-initialised; peer_flags[s] = rank s's array), wait until every rank has published it
 * into my_flags, then store it to *epoch_dev.  Everything this GPU wrote before the call is visible to a peer that passes the barrier. */
int rh_peer_barrier(int32_t* const* peer_flags, const int32_t* my_flags, int rank, int world, int32_t* epoch_dev, void* stream);
/* Wait until every rank's flag (my_flags[0..world), written by the peers' rh_dense_pack_signal of this step) has reached
 * *epoch_dev + 1.  Because a rank publishes behind its backward kernels and a system fence, this also orders the peers' row-gradient
 * REDs before whatever follows on `stream` — the step's third barrier without a signal round of its own. */
int rh_peer_wait(const int32_t* my_flags, int world, const int32_t* epoch_dev, void* stream);

/* Push this rank's gradients (grads[i] NULL = zeros) and up to 4 extra device scalars into slot [rank] of EVERY rank's staging buffer
 * (peer_stage[s] = rank s's buffer, peer-mapped: 2 step parities x world slots of rh_dense_stage_floats() floats), then publish the
 * step number to every rank's flags (peer_flags[s] = rank s's flags array). */
int rh_dense_pack_signal(int n_tensors, const float* const* grads, const int64_t* numel,
                         const float* const* extra, int n_extra, float* const* peer_stage,
                         int32_t* const* peer_flags, int rank, int world,
                         const int32_t* epoch_dev, int32_t* ticket_dev, void* stream);
/* Wait for every rank's publication of this step, sum the W slots of THIS rank's staging buffer element-wise in rank order and apply
 * the SGD / Adam / Adagrad update (kinds and device scalars as rh_dense_update); extra_out (n_extra) receives the summed extras.
 * Every rank ends with bit-identical parameters. */
int rh_dense_reduce_update(int n_tensors, float* const* params, float* const* state1, float* const* state2, const int64_t* numel,
                           int n_extra, float* extra_out, const float* stage, const int32_t* flags,
                           int rank, int world, int32_t* epoch_dev, int32_t* ticket_dev,
                           int kind, const float* lr_dev, const float* bias_corr_dev,
                           float beta1, float beta2, float eps, float weight_decay, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Tower GEMM on the tcgen05 tensor cores, fp32-accurate (3xTF32: hi*hi + hi*lo + lo*hi, fp32 accumulators in TMEM).
 *   C[M, N] (+)= A[M, K] * B[N, K]^T (+ bias[N])          — the Linear layers of MLP.forward / backward
 *   (reference basic/layers.py:281-292; ATen addmm / mm there).
 *   a_mn_major == 0: A is stored row-major [M][lda] (K contiguous);  != 0: A is stored [K][lda] (M contiguous), i.e. the
 *   caller passes the matrix whose TRANSPOSE is the operand (dW = dH^T X reads dH and X as stored).  Same for B.
 *   lda, ldb must be multiples of 4 floats and A, B 16-byte aligned (TMA); C any ldc >= N.  Rows of C that are 16-byte aligned
 *   (ldc % 4 == 0, C 16-byte aligned, ldc >= 4 * ceil(N / 4)) leave through bulk tensor stores whose clipping works on 16-byte
 *   pieces: columns N .. 4 * ceil(N / 4) - 1 of such a row are PADDING and may be overwritten (with the zero products of the
 *   out-of-range columns).  Any other C is written element-exactly by the lane-per-row epilogue.
 *   split_k > 1: K is cut into split_k slices accumulated with red.global.add — C must be zero on entry.
 * ------------------------------------------------------------------------------------------- */
int rh_gemm_tf32x3(const float* A, int64_t lda, int a_mn_major,
                   const float* B, int64_t ldb, int b_mn_major,
                   float* C, int64_t ldc, int M, int N, int K,
                   const float* bias, int split_k, void* stream);
/* Output-tile width of rh_gemm_tf32x3: 0 = chosen per problem (64 when 128-wide tiles would leave more than half of the SMs idle and
 * the 64-wide grid still fits one wave), 64 / 128 = forced (A/B runs, tests).  Returns the setting in force. */
int rh_gemm_tile_n(int set);
/* Kernel variants of rh_gemm_tf32x3 (A/B runs, tests); an argument < 0 leaves that switch alone; returns bit 0 | bit 1 in force.
 *   tma_epilogue (bit 0, default on): the accumulators leave through shared memory and cp.async.bulk.tensor stores (cp.reduce ... add
 *     for split-K) issued by all 8 warps, instead of lane-per-row st.global / red.global from 4 warps (kept for C with ldc % 4 != 0);
 *   concat_b (bit 1, default on): the (hi, lo) twins of the B tile lie back to back, so hi*hi and hi*lo are ONE tcgen05.mma of width
 *     2 BN — 8 MMAs and 5 operand-tile reads per k-block instead of 12 and 6.  Same sums either way. */
int rh_gemm_options(int tma_epilogue, int concat_b);


/* The same GEMM (split_k = 1) with BatchNorm1d's training-mode column statistics of C = A B^T + bias computed in the epilogue —
 * rh_gemm_tf32x3 followed by rh_colstats in one launch (MLP.forward, basic/layers.py:282-283: Linear then BatchNorm1d).
 * Every CTA reduces its 128 rows to a per-column (mean, M2) pair (two-pass inside each warp, Chan's merge above it); the
 * last CTA of each 128-column block merges the row tiles in a fixed order and writes
 *   stats (2N + 1): mean | biased variance | step-counter bits          — exactly what rh_colstats produces,
 * updates running_mean / running_var (unbiased) with `momentum` and increments num_batches_tracked (all may be NULL).
 *   scratch: rh_gemm_stats_scratch_floats(M, N) floats, zeroed ONCE by the caller (the kernel leaves its tickets zero). */
int64_t rh_gemm_stats_scratch_floats(int M, int N);
int rh_gemm_tf32x3_stats(const float* A, int64_t lda, int a_mn_major,
                         const float* B, int64_t ldb, int b_mn_major,
                         float* C, int64_t ldc, int M, int N, int K, const float* bias,
                         float* stats, float* scratch,
                         float* running_mean, float* running_var, int64_t* num_batches_tracked,
                         float momentum, void* stream);

/* ---------------------------------------------------------------------------------------------
 * DCN-v2 cross networks: CrossNetMix (mixture of low-rank experts, basic/layers.py:447-506, forward at :470) and
 * CrossNetV2 (basic/layers.py:423-444).  A CrossNetMix layer is three rh_gemm_tf32x3 products over PACKED operands
 *     [A | G] = x_l Wcat^T      Wcat (E r + E, W): rows (e, j) = V_e[:, j] (v_list[l][e]), rows E r + e = gating[e].weight
 *     P       = t1 Cbd^T        Cbd  (E r, E r)  : block diagonal of C_e (c_list[l][e])
 *     u       = z Ucat^T        Ucat (W, E r)    : [w, (e, j)] = U_e[w, j] (u_list[l][e])
 * joined by the maps below; CrossNetV2 is one product per layer + rh_crossmix_out_fwd.  All matrices row-major fp32.
 * ------------------------------------------------------------------------------------------- */

/* Packed operands of ALL layers in one launch.  u/v/c/wcat/cbd/ucat: host arrays of n_layers device pointers
 * (u, v: (E, W, r); c: (E, r, r)); gate: host array of n_experts device pointers ((W) each, shared by all layers, :466);
 * wcat rows have stride ld_w >= W (columns >= W are zero-filled).  n_layers <= 8, n_experts <= 8. */
int rh_crossmix_pack(int n_layers, int n_experts, int width, int rank,
                     const float* const* u, const float* const* v, const float* const* c, const float* const* gate,
                     float* const* wcat, int64_t ld_w, float* const* cbd, float* const* ucat, void* stream);
/* The reverse for the gradients: d_u/d_v/d_c in the reference's parameter layout from the packed d_ucat / d_wcat / d_cbd
 * (only the diagonal blocks of d_cbd are read); d_gate[e] (W) = sum over layers of d_wcat[l][E r + e, :] (written). */
int rh_crossmix_unpack_grads(int n_layers, int n_experts, int width, int rank,
                             const float* const* d_wcat, int64_t ld_dw, const float* const* d_cbd, int64_t ld_dc,
                             const float* const* d_ucat, int64_t ld_du,
                             float* const* d_u, float* const* d_v, float* const* d_c, float* const* d_gate, void* stream);
/* t1 (batch, E r) = tanh(ag[:, :E r]);  s (batch, E) = softmax(ag[:, E r : E r + E])      (layers.py:486-488,496-498) */
int rh_crossmix_mid1_fwd(const float* ag, int64_t ld_ag, int64_t batch, int n_experts, int rank, float* t1, float* s, void* stream);
/* t2 (batch, E r) = tanh(P);  z[:, (e, j)] = s[:, e] * t2[:, (e, j)]                      (layers.py:489-490,502) */
int rh_crossmix_mid2_fwd(const float* P, const float* s, int64_t batch, int n_experts, int rank, float* t2, float* z, void* stream);
/* The cross step.  bias_outside == 0: out = x0 * (u + bias) + xl  (CrossNetMix, :494,503);  != 0: out = x0 * u + bias + xl
 * (CrossNetV2, :443);  bias (width) or NULL */
int rh_crossmix_out_fwd(const float* x0, int64_t ld0, const float* xl, int64_t ldl, const float* u, int64_t ldu, const float* bias,
                        int bias_outside, int64_t batch, int width, float* out, int64_t ldo, void* stream);
/* Backward of the cross step for an incoming gradient g = g1 (+ g2, may be NULL):  g_sum = g (or NULL);  d_u = g * x0;
 * d_x0_acc += g * (u + bias);  d_bias (width, zeroed by the caller, or NULL) += column sums of d_u
 * (bias_outside: d_x0_acc += g * u;  d_bias += column sums of g). */
int rh_crossmix_out_bwd(const float* g1, int64_t ldg1, const float* g2, int64_t ldg2, const float* x0, int64_t ld0,
                        const float* u, int64_t ldu, const float* bias, int bias_outside, int64_t batch, int width,
                        float* g_sum, int64_t ldgs, float* d_u, int64_t lddu, float* d_x0_acc, int64_t ldx, float* d_bias, void* stream);
/* d_P = d_z * s_e * (1 - t2^2);  d_ag[:, E r + e] = softmax backward of d_s_e = sum_j d_z[(e, j)] t2[(e, j)] */
int rh_crossmix_mid2_bwd(const float* d_z, int64_t ld_dz, const float* s, const float* t2, int64_t batch, int n_experts, int rank,
                         float* d_P, float* d_ag, int64_t ld_dag, void* stream);
/* d_ag[:, :E r] = d_t1 * (1 - t1^2) */
int rh_crossmix_mid1_bwd(const float* d_t1, int64_t ld_dt1, const float* t1, int64_t batch, int n_experts, int rank,
                         float* d_ag, int64_t ld_dag, void* stream);
/* out = a + b + c over (rows, cols) views with their own row strides (b, c may be NULL) */
int rh_sum3(const float* a, int64_t lda, const float* b, int64_t ldb, const float* c, int64_t ldc, int64_t rows, int cols,
            float* out, int64_t ldo, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Two-tower in-batch negatives (trainers/match_trainer.py:118-140; utils/match.py:104-161).
 * ------------------------------------------------------------------------------------------- */

/* picks (batch, k) int64: for every row k DISTINCT columns other than the row's own, uniformly at random (order included) —
 * utils/match.py:136-145's one randperm(B - 1) per row, for all rows in one launch.  The stream is keyed by *seed_dev (device int64: the
 * caller draws it from its torch Generator, no host sync).  k == batch - 1 lists every other column; otherwise k <= 128. */
int rh_inbatch_sample_random(int batch, int k, const int64_t* seed_dev, int64_t* picks, void* stream);
/* Hard negatives (utils/match.py:131-135): the k best-scoring off-diagonal columns of every row of scores (batch, batch; row stride ld),
 * best first, ties to the lower column. */
int rh_inbatch_sample_hard(const float* scores, int64_t ld, int batch, int k, int64_t* picks, void* stream);
/* Cross entropy over [positive | sampled negatives] with the logits taken as dot products of the tower outputs
 * (gather_inbatch_logits + CrossEntropyLoss(mean), utils/match.py:150-161 + match_trainer.py:137-140):
 *   prob (batch, 1 + k) = softmax([<u_i, v_i>, <u_i, v_picks[i, :]>]);  loss_rows (batch) = -log prob[:, 0]. */
int rh_inbatch_ce_fwd(const float* user, int64_t ldu, const float* item, int64_t ldv, int dim, const int64_t* picks,
                      int batch, int k, float* prob, float* loss_rows, void* stream);
/* Backward for loss = mean(loss_rows) with upstream gradient *d_loss (device scalar): d_user (batch, dim) is written, d_item
 * (batch, dim, ZEROED by the caller) accumulates by REDs; the (batch, batch) score gradient is never formed.  dim <= 256. */
int rh_inbatch_ce_bwd(const float* user, int64_t ldu, const float* item, int64_t ldv, int dim, const int64_t* picks,
                      const float* prob, const float* d_loss, int batch, int k,
                      float* d_user, int64_t lddu, float* d_item, int64_t lddv, void* stream);

/* ---------------------------------------------------------------------------------------------
 * DIN target attention (models/ranking/din.py:77-93, ActivationUnit.forward).
 * ------------------------------------------------------------------------------------------- */

/* att_in[b*L + l, :] = [t, h, t-h, t*h] with h = hist_table[hist_ids[b,l]], t = tgt_table[tgt_ids[b]]
 * (din.py:80-81) — the gather of the behaviour sequence fused with the feature construction.
 * Also emits hist (batch, L, dim) and target (batch, dim) when non-NULL. */
int rh_din_attn_input_fwd(const float* hist_table, int hist_vocab,
                          const float* tgt_table, int tgt_vocab, int dim,
                          const void* hist_ids, const void* tgt_ids, int ids_are_i32,
                          int64_t tgt_id_stride, int batch, int seq_len,
                          float* att_in, float* hist_out, float* tgt_out,
                          int32_t* err_flag, void* stream);

/* out[b,:] = sum_l w[b,l] * hist[b,l,:]   (din.py:92); optional softmax over l first (din.py:86-87). */
int rh_din_weighted_sum_fwd(const float* att_w, const float* hist, int batch, int seq_len, int dim,
                            int use_softmax, float* w_used, float* out, void* stream);

/* Backward of the attention-pooling tail + the att_in construction, ending in the scatter-add into
 * the two tables' gradient buffers:
 *   d_out (batch, dim), d_att_in (batch*L, 4*dim) -> d_att_w (batch, L) [written],
 *   hist_grad[hist_ids] += ..., tgt_grad[tgt_ids] += ... (+ d_tgt_extra (batch, dim), the gradient
 *   reaching the target embedding from the final MLP's concat, may be NULL). */
int rh_din_weighted_sum_bwd(const float* w_used, const float* hist, const float* d_out,
                            int batch, int seq_len, int dim, int use_softmax,
                            float* d_att_w, float* d_hist, void* stream);

int rh_din_attn_input_bwd(float* hist_grad, int hist_vocab, int hist_padding_idx,
                          float* tgt_grad, int tgt_vocab, int tgt_padding_idx, int dim,
                          const void* hist_ids, const void* tgt_ids, int ids_are_i32,
                          int64_t tgt_id_stride, int batch, int seq_len,
                          const float* hist, const float* tgt,
                          const float* d_att_in, const float* d_hist, const float* d_tgt_extra,
                          int32_t* err_flag, void* stream);

#ifdef __cplusplus
}
#endif
#endif  /* RECHUB_B200_H_ */
