/* rb200.h — C ABI of librb200.so, the B200 (sm_100a) implementation of recommendit's two-tower /
 * IVF hot path.
 *
 * The reference (sarihammad/recommendit) has no FFI of its own: its hot path is Python that calls
 * PyTorch and FAISS.  Each entry point below names the reference lines whose arithmetic it
 * replaces (paths relative to the reference repo).  The Python classes in recommendit_b200/
 * (TwoTowerModel, FAISSIndex) bind these with ctypes; INTEGRATION.md shows the stub.
 *
 * Conventions
 *   - every function returns 0 on success, a negative RB200_ERR_* otherwise; rb200_last_error()
 *     returns a thread-local message.  Nothing aborts the process.
 *   - every pointer is a DEVICE pointer into caller-owned memory unless the name ends in _host.
 *     The library never allocates device memory; workspaces are sized by *_workspace_bytes().
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).  All work is
 *     enqueued asynchronously; entry points are CUDA-graph capturable.
 *   - matrices are row-major, contiguous, fp32; ids are int64 (torch.long).
 *   - Linear weights use the torch layout W[out][in].
 */
#ifndef RB200_H
#define RB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RB200_VERSION 100

#define RB200_OK 0
#define RB200_ERR_INVALID (-1)     /* bad argument / unsupported shape */
#define RB200_ERR_CUDA (-2)        /* CUDA runtime error, message in rb200_last_error() */
#define RB200_ERR_WORKSPACE (-3)   /* workspace too small */

int rb200_version(void);
const char* rb200_last_error(void);
/* number of SMs of the current device (148 on B200); used by callers to size workspaces */
int rb200_sm_count(void);
/* sizeof() of the ABI structs (0 tower_job, 1 tower_bwd_job, 2 opt_state, 3 step_params,
 * 4 step_views, 5 sumsq_seg) so that bindings can verify their mirrors */
size_t rb200_sizeof(int which);
/* kernels of this library launched so far by this process (host-side count; cub's internal launches
 * and graph replays are not included) */
uint64_t rb200_launch_count(void);
/* measurement aid: a one-thread kernel on `stream` writes the device's %globaltimer (ns) to out[idx] (device memory).
 * Capturable, so phase boundaries can be timed inside a CUDA-graph replay (bench_sharded.py's stage table). */
int rb200_stamp(uint64_t* out, int idx, void* stream);

/* ------------------------------------------------------------------------------------------ *
 * Towers — src/models/two_tower.py:39-42 (UserTower.forward) and :68-72 (ItemTower.forward)
 *   y = normalize( W2 · dropout(relu(W1 · [table[ids] ; extra] + b1)) + b2 ),  eps = 1e-12
 * Supported widths: D in {32,64,128}, H in {64,128,256}, extra_dim <= 64 (18 genres in the
 * reference), subject to the 227 KB shared-memory budget (D=128 needs H<=128).
 * ------------------------------------------------------------------------------------------ */
typedef struct rb200_tower_job {
    const float* table;        /* [n_rows, D] embedding table                                  */
    const int64_t* ids;        /* [B]                                                          */
    const float* extra;        /* [B, extra_dim] appended input columns (genres) or NULL       */
    const float* W1;           /* [H, D+extra_dim]                                             */
    const float* b1;           /* [H]                                                          */
    const float* W2;           /* [D, H]                                                       */
    const float* b2;           /* [D]                                                          */
    float* out;                /* [B, D]  L2-normalised tower output                           */
    float* hid;                /* [B, H]  post-ReLU/dropout hidden, saved for backward; or NULL */
    float* denom;              /* [B]     max(||pre||, 1e-12), saved for backward; or NULL     */
    const uint8_t* keep_mask;  /* [B, H]  optional explicit dropout keep-mask (parity runs)    */
    int64_t n_rows;            /* rows in table (ids outside [0,n_rows) raise the error flag)  */
    int B;
    int extra_dim;
    int extra_by_id;           /* 0: extra is [B, extra_dim] per sample (the reference's batch
                                  format); 1: extra is [n_rows, extra_dim], looked up by id      */
    const void* img;           /* tensor-core modes: optional weight image from rb200_tower_prep
                                  (NULL ⇒ staged internally from the workspace)                  */
} rb200_tower_job;

/* Runs up to 3 tower evaluations (user / positive items / negative items) in ONE launch; the
 * 148 SMs are partitioned between the jobs in proportion to their flops.
 * dropout_p = 0 ⇒ eval mode.  With dropout_p > 0 and keep_mask == NULL a Philox4x32-10 stream
 * keyed by (seed, offset + 3*offset_dev[0] + job index) generates the mask in-kernel; offset_dev
 * (optional device int64, e.g. &opt_state.step) lets a replayed CUDA graph draw a fresh mask.
 * err_flag: optional device int, bit 0 is set when an id is out of range. */
int rb200_tower_fwd(const rb200_tower_job* jobs_host, int n_jobs, int D, int H, float dropout_p,
                    uint64_t seed, uint64_t offset, const int64_t* offset_dev, int mode, int* err_flag,
                    void* workspace, size_t workspace_bytes, void* stream);
/* workspace of rb200_tower_fwd: 0 bytes in mode 0; room for the weight images of the jobs whose `img` is NULL otherwise */
size_t rb200_tower_fwd_workspace_bytes(int n_jobs, int D, int H, int extra_dim, int mode);
/* Tensor-core modes consume the MLP weights as ready-made UMMA operand images (hi/lo-split, canonical K-major layout,
 * also the transposed ones the backward needs), which a CTA pulls into shared memory with ONE bulk asynchronous copy
 * (TMA engine).  rb200_tower_prep builds the image of one weight set (W1 [H, D+extra_dim], W2 [D, H]) into `img`
 * (rb200_tower_img_bytes bytes, 16-byte aligned); it must be re-run whenever the weights change. */
size_t rb200_tower_img_bytes(int D, int H, int extra_dim);
int rb200_tower_prep(const float* W1, const float* W2, int D, int H, int extra_dim, void* img, void* stream);
/* mode (both tower entry points): 0 = fp32 FFMA kernels (parity mode, any supported width);
 *   1 = tcgen05 tensor cores, single TF32 (fast mode, ~1e-3 relative error on activations/gradients);
 *   2 = tcgen05 tensor cores, 3xTF32 error-compensated (fp32-grade, meets the 1e-5 parity bound).
 * Tensor-core modes currently cover D = 64, H = 128, extra_dim <= 24 (the reference's production widths).
 * All modes draw the same dropout mask for the same (seed, offset). */

/* Backward of one tower evaluation (autograd of the above; src/training/train_embeddings.py:190).
 * Step 1 (data): from dY [B,D] (gradient w.r.t. the normalised output) and the saved y/denom/hid
 *   computes dpre [B,D], dact [B,H] (gradient at the first Linear's output) and dRows [B,D]
 *   (gradient w.r.t. the gathered embedding rows).
 * Step 2 (weights): dW1 = dactᵀ·X, db1 = Σ dact, dW2 = dpreᵀ·hid, db2 = Σ dpre, reduced over the
 *   batch with a fixed-order two-stage sum (deterministic).  grads_out is the flat block
 *   [W1 (H·Din) | b1 (H) | W2 (D·H) | b2 (D)]; accumulate != 0 adds to it instead of overwriting.
 * Several jobs that share weights (positive + negative items) are reduced into one grads_out. */
typedef struct rb200_tower_bwd_job {
    const float* table; const int64_t* ids; const float* extra; int64_t n_rows; int B; int extra_dim;
    int extra_by_id;
    const float* W1; const float* W2;
    const float* dY;      /* [B,D] */
    const float* y;       /* [B,D] saved forward output */
    const float* denom;   /* [B]   saved */
    const float* hid;     /* [B,H] saved */
    float* dpre;          /* [B,D] scratch/out */
    float* dact;          /* [B,H] scratch/out */
    float* dRows;         /* [B,D] out */
    const void* img;      /* tensor-core modes: optional weight image from rb200_tower_prep, or NULL */
} rb200_tower_bwd_job;

size_t rb200_tower_bwd_workspace_bytes(int D, int H, int extra_dim);
int rb200_tower_bwd(const rb200_tower_bwd_job* jobs_host, int n_jobs, int D, int H, float dropout_p,
                    int mode, float* grads_out, int accumulate, void* workspace, size_t workspace_bytes,
                    void* stream);

/* ------------------------------------------------------------------------------------------ *
 * Losses
 * ------------------------------------------------------------------------------------------ */
/* TwoTowerModel.bpr_loss — src/models/two_tower.py:117-130:
 *   loss = mean_i softplus(-(u_i·p_i - u_i·n_i)).  Writes loss[0]; du/dp/dn (each [B,D]) are the
 *   gradients of `loss` scaled by grad_scale (pass 1.0f; may be NULL to skip backward). */
int rb200_bpr_pair(const float* u, const float* p, const float* n, int B, int D, float* loss,
                   float* du, float* dp, float* dn, float grad_scale, void* workspace,
                   size_t workspace_bytes, void* stream);
size_t rb200_bpr_pair_workspace_bytes(int B);

/* TwoTowerModel.in_batch_bpr_loss — src/models/two_tower.py:132-160, closed form
 *   loss = Σ_{i≠j} softplus(S_ij − S_ii) / (B(B−1)),  S = U·Iᵀ  (S is never written to HBM).
 * dU/dI may be NULL (forward only).  mode: 0 = fp32 FFMA (parity mode, 1e-5),
 * 1 = tcgen05 TF32 (fast mode; stated bound 2e-3 on gradients), 2 = tcgen05 3xTF32 (1e-5). */
size_t rb200_bpr_inbatch_workspace_bytes(int B, int D);
int rb200_bpr_inbatch(const float* U, const float* I, int B, int D, int mode, float* loss, float* dU,
                      float* dI, float grad_scale, void* workspace, size_t workspace_bytes,
                      void* stream);

/* ------------------------------------------------------------------------------------------ *
 * Sparse embedding gradient — autograd of nn.Embedding(padding_idx=0) (two_tower.py:27,54):
 * deterministic sorted-segment sum of per-sample row gradients.
 *   ids [B], rows [B,D] → for every distinct id != padding_idx the sum of its rows, added in
 *   ascending sample order.
 *   dense_grad (optional) [n_rows, D]: the sums are ADDED into it (caller zero-fills).
 *   uniq_ids / uniq_grads / n_uniq (optional): compact output, ascending id.
 *   row_slot (optional) [n_rows] int32: row_slot[id] = index into uniq_* (caller keeps it at -1;
 *   rb200_scatter_reset_slots restores it).
 * ------------------------------------------------------------------------------------------ */
size_t rb200_scatter_workspace_bytes(int B, int64_t n_rows);
int rb200_scatter_rows(const int64_t* ids, const float* rows, int B, int D, int64_t n_rows,
                       int64_t padding_idx, float* dense_grad, int64_t* uniq_ids, float* uniq_grads,
                       int* n_uniq, int* row_slot, void* workspace, size_t workspace_bytes,
                       void* stream);
/* The same result in two phases: rb200_scatter_plan needs only the ids (sort, segment heads, compact id list: uniq_ids / n_uniq /
 * row_slot) and may run on another stream before the gradient rows exist; rb200_scatter_apply then adds the rows up.  Both use the
 * same workspace (rb200_scatter_workspace_bytes), which must stay untouched in between.  n_rows < 2^31. */
int rb200_scatter_plan(const int64_t* ids, int B, int64_t n_rows, int64_t padding_idx, int64_t* uniq_ids, int* n_uniq,
                       int* row_slot, void* workspace, size_t workspace_bytes, void* stream);
int rb200_scatter_apply(const float* rows, int B, int D, int64_t n_rows, float* dense_grad, const int64_t* uniq_ids,
                        float* uniq_grads, const int* n_uniq, void* workspace, size_t workspace_bytes, void* stream);
int rb200_scatter_reset_slots(const int64_t* uniq_ids, const int* n_uniq, int max_uniq, int* row_slot,
                              void* stream);
/* row_slot[uniq_ids[i]] = i for i < n_uniq[0] (when the compact list was produced without a slot map) */
int rb200_scatter_set_slots(const int64_t* uniq_ids, const int* n_uniq, int max_uniq, int* row_slot,
                            void* stream);
/* Device-side batch producer: the sample stream of the reference's UserItemDataset + DataLoader
 * (src/training/train_embeddings.py:23-79, 144-151: positives shuffled per epoch with drop_last, one negative per sample drawn
 * uniformly from the catalog and rejected while the user has rated it).  pos_users/pos_items: the n_pos positive pairs;
 * rated_offsets [max user id + 2] / rated_items: CSR of every user's rated items, ascending within a user; catalog: the
 * n_cat candidate item ids.  Batch `step` of `epoch` (0-based) is a pure function of `seed`.  Several data-parallel ranks share one
 * epoch: rank r of `world` takes slots [(step·world + r)·B, +B) of the epoch's permutation, so the ranks' batches are disjoint and
 * a step consumes world·B positives ((step+1)·world·B <= n_pos; world = 1, rank = 0 for a single process). */
int rb200_sample_batch(const int64_t* pos_users, const int64_t* pos_items, int64_t n_pos, const int64_t* rated_offsets,
                       const int64_t* rated_items, const int64_t* catalog, int64_t n_cat, int B, uint64_t seed,
                       int64_t epoch, int64_t step, int64_t rank, int64_t world, int64_t* out_users, int64_t* out_pos,
                       int64_t* out_neg, void* stream);
/* The same producer as a description the fused step can run by itself (rb200_step_params.next_batch): at the end of step t it
 * writes the batch of step t+1 into the step's own id buffers, on a side stream under the optimizer kernels.  The global batch
 * index g is read on the device from the optimizer's step counter (opt->step = number of steps begun so far), epoch =
 * g / batches_per_epoch, step = g % batches_per_epoch — nothing comes from the host, so the whole epoch is graph replays.
 * rated_bitmap (optional, [n_users + 1][bitmap_words] uint32, bit i of a user's row = "has rated item i") replaces the binary
 * search in the CSR by one load; results are identical. */
typedef struct rb200_sampler {
    const int64_t *pos_users, *pos_items; int64_t n_pos;
    const int64_t *rated_offsets, *rated_items;
    const uint32_t* rated_bitmap; int64_t bitmap_words;
    const int64_t* catalog; int64_t n_cat;
    uint64_t seed; int64_t batches_per_epoch;
    int64_t rank, world;               /* data-parallel slice of every step (world <= 1: single process) */
} rb200_sampler;
/* batch g = *counter_dev of the stream described by `s` (what the fused step runs); world·B·batches_per_epoch <= n_pos */
int rb200_sample_batch_dev(const rb200_sampler* s, int B, const int64_t* counter_dev, int64_t* out_users, int64_t* out_pos,
                           int64_t* out_neg, void* stream);
/* Exchange plan of the row-sharded step (tables sharded by id mod world, both shards of a rank in
 * one tensor: user rows first, item rows behind them).  Requests = [user_ids | item_ids] (n = n_user + n_item) in sample
 * order; bucket order = stable by owner.  perm[j] = sample of bucket position j, inv = its inverse, local_rows[j] = row in
 * the owner's combined shard (id / world, + user_rows_by_rank[owner] for items), send_counts[w] = requests owned by w.
 * All device pointers; deterministic. */
size_t rb200_route_plan_workspace_bytes(int64_t n, int world);
int rb200_route_plan(const int64_t* user_ids, int64_t n_user, const int64_t* item_ids, int64_t n_item, int world,
                     const int64_t* user_rows_by_rank, int64_t* perm, int64_t* inv, int64_t* local_rows,
                     int64_t* send_counts, void* workspace, size_t workspace_bytes, void* stream);
/* Fixed-capacity form of the same plan (the CUDA-graph-captured sharded step: every all-to-all gets equal, host-known splits, so
 * nothing is read back from the device).  Bucket w of the request list occupies slots [w·capacity, (w+1)·capacity) of a padded
 * buffer: slot_of_sample[i] = slot of request i (sample order), send_rows[world·capacity] = the owner-local row in each slot,
 * -1 in empty slots.  A request that does not fit its bucket gets the DUMMY slot world·capacity (one past the buffer: the caller
 * keeps a zero row there on the way in and drops that gradient row on the way out) and is counted in *overflow
 * (device int64, accumulated — the caller reports it: recommendit_b200/sharded.py check_exchange). */
size_t rb200_route_plan_padded_workspace_bytes(int64_t n, int world);
int rb200_route_plan_padded(const int64_t* user_ids, int64_t n_user, const int64_t* item_ids, int64_t n_item, int world,
                            const int64_t* user_rows_by_rank, int64_t capacity, int64_t* slot_of_sample,
                            int64_t* send_rows, int64_t* overflow, void* workspace, size_t workspace_bytes,
                            void* stream);
/* out[i,:] = table[rows[i],:] — the owner-side row gather of row-sharded tables; rows outside
 * [0, n_table_rows) yield zeros */
int rb200_gather_rows(const float* table, const int64_t* rows, int64_t n, int D, int64_t n_table_rows,
                      float* out, void* stream);

/* Peer-memory form of SURVEY.md §8e steps 1-3 and 5 on one NVLink / NVSwitch box: every rank's combined shard (user rows, then
 * item rows; owner(id) = id % world, local row = id / world) and its gradient receive buckets are mapped into every process
 * (torch symmetric memory; the caller passes HOST arrays of the `world` device pointers, and user_rows_by_rank — the number of
 * user rows in each rank's shard — as a HOST array too).  The kernels are collective-free: the
 * caller brackets them with cross-GPU barriers (recommendit_b200/sharded.py, exchange="p2p").
 *   gather: out[r,:] = shard[owner(id_r)][local(id_r),:] for the requests in sample order [user ids | item ids] — replaces the id
 *           all-to-all, the owner-side gather and the row all-to-all.  Ids outside their table read row 0 and set bit 0 of
 *           *err_flag (may be NULL).
 *   push:   gradient row r → owner's grad bucket, slot rank·capacity + off, where slot_of_sample[r] = owner·capacity + off comes
 *           from rb200_route_plan_padded (slot world·capacity = overflowed: dropped); send_rows (the plan's owner-local row per
 *           slot, -1 = empty) is written to the owners' row buckets alongside — replaces the gradient all-to-all.  Ends with a
 *           system-scope fence. */
#define RB200_MAX_PEERS 16
int rb200_gather_rows_sharded(const void* const* shard_ptrs, const int64_t* user_rows_by_rank, int world,
                              const int64_t* user_ids, int64_t n_user, const int64_t* item_ids, int64_t n_item,
                              int64_t n_user_rows, int64_t n_item_rows, int D, float* out, int* err_flag, void* stream);
int rb200_push_rows_sharded(void* const* grad_bucket_ptrs, void* const* row_bucket_ptrs, int world, int rank, int64_t capacity,
                            const float* drows, const int64_t* slot_of_sample, int64_t n, int D, const int64_t* send_rows,
                            void* stream);
/* The plan's row list alone (send_rows [world·capacity] → the owners' row buckets; rb200_push_rows_sharded then takes
 * send_rows = NULL and row_bucket_ptrs may hold NULLs): it depends only on the ids, so it can be sent at the start of the step and
 * the owners run rb200_scatter_plan on it under the towers.  Ends with a system-scope fence. */
int rb200_push_row_lists_sharded(void* const* row_bucket_ptrs, int world, int rank, int64_t capacity, const int64_t* send_rows,
                                 void* stream);
/* Reductions of the same step over peer memory ("ncclAllReduce of MLP grads and of {norm², loss}"),
 * deterministic and identical on every rank: each rank reads all ranks' buffers and adds them in rank order.
 *   allreduce_oneshot: out[i] = Σ_k src_ptrs[k][i] (n floats, n % 4 == 0; HOST array of `world` device pointers).
 *   scalars_publish:   slot[0..3] = {hi(st->sumsq), lo(st->sumsq), loss[0]·scale, 0} (this rank's partial, peer-readable).
 *   scalars_reduce:    st->sumsq = Σ_k (slot_k[0] + slot_k[1]) in fp64, st->loss = Σ_k slot_k[2].
 * The caller orders them with cross-GPU barriers (publish → barrier → reduce). */
int rb200_allreduce_oneshot(const void* const* src_ptrs, int world, int64_t n, float* out, void* stream);
/* Two-shot all-reduce IN PLACE over peer memory (the data-parallel step: one buffer of dense gradients per rank, mapped into every
 * process): rank r sums slice r of all ranks' buffers in rank order and writes the sum into slice r of every rank's buffer — all
 * ranks end up with bit-identical sums.  buf_ptrs: HOST array of the `world` device pointers; n floats, n % 4 == 0.  The caller puts
 * a cross-GPU barrier before (all buffers complete) and after (all slices delivered).  Replaces ncclAllReduce inside the step. */
int rb200_allreduce_twoshot(void* const* buf_ptrs, int world, int rank, int64_t n, void* stream);
/* The same reduction inside the NVSwitch (NVLS multicast): multicast_ptr = the multicast address of the ranks' buffers (torch symmetric
 * memory: handle.multicast_ptr; 0 when the fabric has no multicast support — use the two-shot form then).  Rank r reduces slice r with
 * multimem.ld_reduce and broadcasts the sum with multimem.st.  Same barriers around it. */
int rb200_allreduce_multimem(void* multicast_ptr, int world, int rank, int64_t n, void* stream);
/* (rb200_sharded_scalars_publish / _reduce are declared below, after rb200_opt_state) */

/* ------------------------------------------------------------------------------------------ *
 * clip_grad_norm_ + Adam — src/training/train_embeddings.py:160,191-192
 * Optimiser scalars live in a small device block so that a captured CUDA graph can be replayed
 * step after step without host involvement.
 * ------------------------------------------------------------------------------------------ */
typedef struct rb200_opt_state {   /* device-resident, 128 bytes; host fills the first block */
    double lr, beta1, beta2;        /* python doubles, as torch.optim.Adam holds them               */
    double sumsq;                   /* running Σ g² of the current step (fp64)                      */
    int64_t step;                   /* completed optimiser steps                                    */
    float eps, weight_decay, max_norm;
    float one_minus_beta1, one_minus_beta2, beta2_f;
    float step_size;                /* lr / (1 - beta1^step)          (rb200_opt_begin_step)        */
    float bias_corr2_sqrt;          /* sqrt(1 - beta2^step)           (rb200_opt_begin_step)        */
    float clip_coef;                /* min(1, max_norm/(norm+1e-6))   (rb200_grad_norm_clip)        */
    float total_norm;
    float loss;                     /* last loss, convenience                                       */
    unsigned ticket;        /* internal: last-block election of the grad-norm reduction (always 0 between launches) */
    float pad[10];
} rb200_opt_state;

/* step += 1, recompute bias corrections, sumsq = 0 */
int rb200_opt_begin_step(rb200_opt_state* st, void* stream);
/* peer-memory reduction of the step's scalars, see rb200_allreduce_oneshot above */
int rb200_sharded_scalars_publish(const rb200_opt_state* st, const float* loss, float scale, float* slot, void* stream);
int rb200_sharded_scalars_reduce(const void* const* slot_ptrs, int world, rb200_opt_state* st, void* stream);
/* scalars_reduce + the sum of squares of the replicated dense gradient (n floats, counted once) + the clip coefficient, in one
 * launch: st->sumsq = Σ_k slot_k.sumsq + Σ dense_grad², st->loss = Σ_k slot_k.loss, st->total_norm, st->clip_coef. */
int rb200_sharded_scalars_finish(const void* const* slot_ptrs, int world, const float* dense_grad, int64_t n, rb200_opt_state* st,
                                 void* stream);
/* sumsq += Σ x² over up to 4 segments.  A segment is n floats at x; when `count` (device int) is
 * set the length is count[0]*row_len instead (compact unique-row gradients) and n is only the
 * capacity used to size the grid.  Deterministic (fixed partition, fixed-order fp64 finalisation). */
typedef struct rb200_sumsq_seg { const float* x; int64_t n; const int* count; int row_len; } rb200_sumsq_seg;
int rb200_sumsq_accumulate(rb200_opt_state* st, const rb200_sumsq_seg* segs_host, int n_segs,
                           void* workspace, size_t workspace_bytes, void* stream);
size_t rb200_sumsq_workspace_bytes(void);
/* total_norm = sqrt(sumsq); clip_coef = min(1, max_norm / (total_norm + 1e-6)) */
int rb200_grad_norm_clip(rb200_opt_state* st, void* stream);
/* torch.optim.Adam (coupled L2) on n contiguous elements; g may be NULL (gradient 0). */
int rb200_adam_dense(float* w, const float* g, float* m, float* v, int64_t n,
                     const rb200_opt_state* st, void* stream);
/* Parity mode for embedding tables: every row gets an Adam update; rows with row_slot[r] >= 0
 * take uniq_grads[row_slot[r]], all others gradient 0 (only the weight-decay term) — identical
 * to the reference's dense nn.Embedding gradient + Adam(weight_decay). */
int rb200_adam_table_dense(float* w, float* m, float* v, int64_t n_rows, int D, const int* row_slot,
                           const float* uniq_grads, const rb200_opt_state* st, void* stream);
/* Throughput mode: Adam only on the touched rows (documented divergence from F5). */
int rb200_adam_rows(float* w, float* m, float* v, int D, const int64_t* uniq_ids,
                    const float* uniq_grads, const int* n_uniq, int max_uniq,
                    const rb200_opt_state* st, void* stream);
/* rb200_adam_rows on a table (shard) and rb200_adam_dense on up to two dense blocks (n0 / n1 floats, 16-byte aligned, may be 0) in ONE
 * launch — the optimizer of the row-sharded step. */
int rb200_adam_rows_dense2(float* w, float* m, float* v, int D, const int64_t* uniq_ids, const float* uniq_grads, const int* n_uniq,
                           int max_uniq, float* w0, const float* g0, float* m0, float* v0, int64_t n0, float* w1, const float* g1,
                           float* m1, float* v1, int64_t n1, const rb200_opt_state* st, void* stream);

/* ------------------------------------------------------------------------------------------ *
 * The whole training step of src/training/train_embeddings.py:183-192 as ONE host call:
 *   towers fwd (1 launch) → loss (+grad) → towers bwd → sorted-segment scatter → global-norm clip
 *   → Adam on MLPs and tables.  Everything is enqueued on `stream`, nothing synchronises, so the
 *   call can be captured into a CUDA graph and replayed.
 * MLP parameters of one tower are ONE flat block [W1 (H·Din) | b1 (H) | W2 (D·H) | b2 (D)]; the
 * Adam moments use the same layout.
 * ------------------------------------------------------------------------------------------ */
typedef struct rb200_step_params {
    int D, H, extra_dim, B;
    int64_t n_user_rows, n_item_rows;
    float *user_table, *user_table_m, *user_table_v;      /* [n_user_rows, D]                    */
    float *item_table, *item_table_m, *item_table_v;      /* [n_item_rows, D]                    */
    float *user_mlp, *user_mlp_m, *user_mlp_v;            /* flat MLP blocks                     */
    float *item_mlp, *item_mlp_m, *item_mlp_v;
    int *user_row_slot, *item_row_slot;                   /* [n_rows] int32, all -1 between steps (dense mode) */
    rb200_opt_state* opt;
    const int64_t *user_ids, *pos_ids, *neg_ids;          /* [B]                                 */
    const float *pos_extra, *neg_extra;                   /* [B, extra_dim], or one [n_item_rows, extra_dim]
                                                             table in both when extra_by_id      */
    int extra_by_id;
    int loss_kind;      /* 0: bpr_loss with sampled negatives (what train_embeddings.py runs)
                           1: in_batch_bpr_loss (neg_* unused)                                   */
    int inbatch_mode;   /* precision mode of rb200_bpr_inbatch                                   */
    int tower_mode;     /* precision mode of rb200_tower_fwd / rb200_tower_bwd (0 FFMA fp32, 1 TF32, 2 3xTF32) */
    int adam_mode;      /* 0: dense — every table row is updated, as torch.optim.Adam(weight_decay) on the
                              reference's dense nn.Embedding gradients (parity mode)
                           1: touched rows only (throughput mode, documented divergence)         */
    float dropout_p; uint64_t seed;
    const uint8_t *keep_mask_user, *keep_mask_pos, *keep_mask_neg;   /* optional explicit masks */
    int64_t padding_idx;
    float* loss;        /* device scalar, also copied to opt->loss                               */
    int* err_flag;      /* optional                                                              */
    void* workspace; size_t workspace_bytes;
    void* const* stage_events_host;   /* optional: 7 cudaEvent_t recorded at the stage boundaries
                                         start | towers fwd | loss | towers bwd | scatter | clip | adam */
    float grad_scale;   /* scale of the loss gradient (0 ⇒ 1).  Data-parallel replicas pass 1/world.         */
    float* dp_grads;    /* data-parallel mode (replicated tables): when set, rb200_bpr_step stops after the
                           gradients and leaves them DENSE in this flat buffer
                           [user MLP | item MLP | user table (n_user_rows·D) | item table (n_item_rows·D)]
                           for the caller to all-reduce; rb200_bpr_apply then clips and runs Adam from it.  */
    const rb200_sampler* next_batch;   /* optional (host pointer): produce the NEXT step's user/pos/neg ids on the device at the
                           end of this step (device-side batch producer; needs extra_by_id or extra_dim == 0) */
} rb200_step_params;

size_t rb200_bpr_step_workspace_bytes(int B, int D, int H, int extra_dim, int64_t n_user_rows,
                                      int64_t n_item_rows, int loss_kind);
int rb200_bpr_step(const rb200_step_params* params_host, void* stream);
/* second half of a data-parallel step: Σg² over dp_grads → clip coefficient → Adam on MLPs and tables (every row) */
int rb200_bpr_apply(const rb200_step_params* params_host, void* stream);
/* number of floats in dp_grads */
size_t rb200_bpr_dp_grad_floats(int D, int H, int extra_dim, int64_t n_user_rows, int64_t n_item_rows);
/* Debug/test access to the step's intermediate gradients inside the workspace (device pointers,
 * valid after rb200_bpr_step on the same workspace): flat MLP grads of both towers and the
 * compact unique-row gradients. */
typedef struct rb200_step_views {
    float *user_mlp_grad, *item_mlp_grad;
    int64_t *user_uniq_ids, *item_uniq_ids;
    float *user_uniq_grads, *item_uniq_grads;
    int *user_n_uniq, *item_n_uniq;
    float *user_emb, *pos_emb, *neg_emb;
} rb200_step_views;
int rb200_bpr_step_views(const rb200_step_params* params_host, rb200_step_views* out_host);

/* ------------------------------------------------------------------------------------------ *
 * Tensor-core score product: C[M,N] (row stride ldc) = A[M,K] · B[N,K]ᵀ, fp32 in / fp32 out, tcgen05.mma kind::tf32 with
 * fp32 accumulation in TMEM.  mode 1 = single TF32 (~1e-3 relative error, fast mode), mode 2 = 3xTF32 error-compensated
 * (~1e-6, fp32-grade).  K % 4 == 0, 16-byte aligned operands.  err_flag (optional device int): bit 1 is set if the
 * tensor-core pipeline timed out (never expected; results are then undefined).  Serves the IVF coarse quantizer
 * (faiss_index.py:113 → IndexFlatIP.search) and exhaustive-search score chunks.
 * ------------------------------------------------------------------------------------------ */
int rb200_gemm_nt(const float* A, int M, const float* B, int N, int K, int mode, float* C, int64_t ldc,
                  int* err_flag, void* stream);

/* ------------------------------------------------------------------------------------------ *
 * IVFFlat inner-product index — src/models/faiss_index.py
 * ------------------------------------------------------------------------------------------ */
/* x / max(||x||, eps) per row (faiss_index.py:64-65,109-110,141-142; eps = 1e-8). out may alias x */
int rb200_normalize_rows(const float* x, int64_t n, int D, float eps, float* out, void* stream);

/* quantizer.assign (top-1 inner product, lowest index on ties) used by index.add
 * (faiss_index.py:74) and by k-means (index.train, :73).  best_score optional. */
int rb200_ivf_assign(const float* x, int64_t n, int D, const float* centroids, int nlist,
                     int32_t* assign_out, float* best_score, void* stream);
/* one spherical k-means update: centroid = normalize(mean of assigned rows); empty lists keep
 * their old centroid and are reported in empty_count (device int).  Deterministic. */
size_t rb200_kmeans_update_workspace_bytes(int64_t n, int nlist);
int rb200_kmeans_update(const float* x, int64_t n, int D, const int32_t* assign, int nlist,
                        float* centroids, int32_t* counts, void* workspace, size_t workspace_bytes,
                        void* stream);
/* index.add: counting sort into CSR inverted lists, insertion order kept inside each list.
 * offsets [nlist+1] int64, list_ids [n] int64 (internal row numbers), list_vecs [n,D]. */
size_t rb200_ivf_build_workspace_bytes(int64_t n, int nlist);
int rb200_ivf_build(const float* x, int64_t n, int D, const int32_t* assign, int nlist, int64_t* offsets,
                    int64_t* list_ids, float* list_vecs, void* workspace, size_t workspace_bytes,
                    void* stream);

/* index.search (faiss_index.py:113,145): coarse top-nprobe by IP → scan of the probed lists →
 * top-k by (score desc, scan position asc).  q must already be normalised.
 * Two calls, because the candidate buffer is sized by the lists the batch actually probes:
 *   rb200_ivf_search_plan  coarse scores, top-nprobe lists per query, (list → queries) inverse map;
 *                          SYNCHRONISES the stream once to return Σ and max of per-query candidate
 *                          counts to the host — unless both host pointers are NULL: then nothing is read
 *                          back (CUDA-graph capturable) and the caller sizes rb200_ivf_search_run by the
 *                          upper bounds total = nq·nprobe·max_list_len, max = nprobe·max_list_len.
 *   rb200_ivf_search_run   list-major scan into the candidate buffer + per-query top-k.
 *   rb200_ivf_search_status  SYNCHRONISES the stream and reports a timed-out tensor-core pipeline of the searches that
 *                          used this plan workspace since its last rb200_ivf_search_plan (RB200_ERR_INVALID + message);
 *                          the host-facing search calls it before it hands results out.
 * out_scores [nq,k] (-FLT_MAX padding), out_ids [nq,k] internal row numbers (-1 padding); k <= 2048. */
size_t rb200_ivf_plan_workspace_bytes(int nq, int nlist, int nprobe);
int rb200_ivf_search_plan(const float* q, int nq, int D, const float* centroids, int nlist, int nprobe,
                          const int64_t* offsets, void* plan_ws, size_t plan_ws_bytes,
                          int64_t* total_candidates_host, int64_t* max_candidates_host, void* stream);
size_t rb200_ivf_search_workspace_bytes(int64_t total_candidates);
int rb200_ivf_search_status(const void* plan_ws, size_t plan_ws_bytes, int nq, int nlist, int nprobe, void* stream);
/* tile_list / tile_idx (optional, int32 [n_tiles]): the index's table of (list, 64-vector tile) work items — one CTA per
 * non-empty tile instead of a (max tiles) × nlist grid full of empty CTAs.  It depends only on `offsets`; build it once.
 * n_vectors: rows of list_vecs (= offsets[nlist]; the D = 64 scan fetches vector tiles with a tensor map over [n_vectors, D]). */
int rb200_ivf_search_run(const float* q, int nq, int D, int nlist, int nprobe, const int64_t* offsets,
                         const int64_t* list_ids, const float* list_vecs, int64_t n_vectors, int64_t max_list_len,
                         const int32_t* tile_list, const int32_t* tile_idx, int n_tiles, int k,
                         void* plan_ws, size_t plan_ws_bytes, int64_t total_candidates,
                         int64_t max_candidates, float* out_scores, int64_t* out_ids, void* workspace,
                         size_t workspace_bytes, void* stream);

/* IndexFlatIP.search — exhaustive inner-product top-k (BASELINE cfg 5): (score desc, row asc).
 * id_base is added to the row numbers (row-sharded databases). */
size_t rb200_flat_search_workspace_bytes(int nq, int64_t n, int k);
int rb200_flat_search(const float* q, int nq, const float* x, int64_t n, int D, int k, int64_t id_base,
                      float* out_scores, int64_t* out_ids, void* workspace, size_t workspace_bytes,
                      void* stream);
/* merge `parts` sorted top-k lists per query (scores/ids laid out [parts, nq, k]) into one
 * [nq, k] list — the per-shard select + allgather merge of BASELINE cfg 5.  Ties keep the
 * lower part first. */
size_t rb200_topk_merge_workspace_bytes(int parts, int nq, int k);
int rb200_topk_merge(const float* scores, const int64_t* ids, int parts, int nq, int k,
                     float* out_scores, int64_t* out_ids, void* workspace, size_t workspace_bytes,
                     void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RB200_H */
