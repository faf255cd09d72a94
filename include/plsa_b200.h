/*
 * plsa_b200 — C ABI of the B200-native pLSA EM engine (libplsa_b200.so).
 *
 * This is the drop-in boundary for the EM hot path of lmcinnes/enstop.  The reference has
 * no FFI of its own; its raw-array seam is the pair of numba entry points
 *     enstop/plsa.py:516-640   plsa_fit_inner(X_rows, X_cols, X_vals, p_w_given_z,
 *                              p_z_given_d, sample_weight, n_iter, n_iter_per_test,
 *                              tolerance, e_step_thresh, use_sample_weights)
 *     enstop/plsa.py:819-920   plsa_refit_inner(X_rows, X_cols, X_vals, topics,
 *                              p_z_given_d, sample_weight, n_iter, n_iter_per_test,
 *                              tolerance, e_step_thresh)
 * (contiguous int32 / float32 buffers, factors mutated in place, scalars by value), and
 * the precedent for swapping the backend is enstop/enstop_.py:52-53,92-114 where
 * `enstop.cuda_plsa.plsa_fit` replaces `enstop.plsa.plsa_fit`.  The one-shot entry points
 * below (plsa_b200_fit_inner / plsa_b200_refit_inner) bind exactly that seam; the context
 * API underneath is what the Python host layer (enstop_b200/plsa.py) uses to keep the
 * corpus resident across calls (ensemble members, transform after fit).
 *
 * Conventions
 *   - every function returns 0 on success, a PLSA_E* code otherwise; no C++ exception or
 *     abort crosses this boundary; plsa_last_error() gives the message.
 *   - all pointers are HOST pointers unless the name says `_device`; the caller owns
 *     every host buffer; the library owns all device memory behind the opaque context.
 *   - layouts are the reference's: P(z|d) is [n, k] row-major float32, P(w|z) is [k, m]
 *     row-major float32 (the device keeps P(w|z) transposed and padded; that is private).
 *   - a context is bound to one device and one stream and must not be used from two
 *     threads at once; different contexts are independent (one per ensemble worker).
 */
#ifndef PLSA_B200_H
#define PLSA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PLSA_OK 0
#define PLSA_EINVAL 1   /* bad argument / state                                     */
#define PLSA_ECUDA 2    /* a CUDA runtime call or kernel failed                      */
#define PLSA_ENOMEM 3   /* host or device allocation failed                          */
#define PLSA_ENCCL 4    /* NCCL could not be loaded or a collective failed           */

#define PLSA_MAX_K 1024 /* same ceiling as the incumbent GPU path (cuda_plsa.py:135) */

typedef struct plsa_ctx plsa_ctx;
typedef struct plsa_comm plsa_comm;

/* ---- library / device ------------------------------------------------------------ */
int plsa_version(void);                      /* 100 * major + minor                   */
int plsa_device_count(int *count);
const char *plsa_last_error(const plsa_ctx *ctx); /* ctx == NULL: last error of a call
                                                     that had no context (thread local) */

int plsa_ctx_create(int device, plsa_ctx **ctx);
int plsa_ctx_destroy(plsa_ctx *ctx);

/* ---- corpus ------------------------------------------------------------------------ */
/* Upload a CSR doc-term matrix (what enstop/plsa.py:714 turns into COO triplets).  The
 * transposed (term-major) copy used by the P(w|z) pass is built on the device. */
int plsa_upload_csr(plsa_ctx *ctx, const int32_t *indptr, const int32_t *indices,
                    const float *data, int64_t n_docs, int64_t n_terms, int64_t nnz);
/* Same, with the values in the caller's dtype (PLSA_F32/F64/I32/I64); the float32 cast of
 * plsa.py:714 is done on the device. */
#define PLSA_F32 0
#define PLSA_F64 1
#define PLSA_I32 2
#define PLSA_I64 3
int plsa_upload_csr_typed(plsa_ctx *ctx, const int32_t *indptr, const int32_t *indices,
                          const void *data, int32_t dtype, int64_t n_docs, int64_t n_terms,
                          int64_t nnz);
/* Same from row-sorted COO triplets — the argument form of plsa_fit_inner. */
int plsa_upload_coo(plsa_ctx *ctx, const int32_t *rows, const int32_t *cols,
                    const float *vals, int64_t n_docs, int64_t n_terms, int64_t nnz);
/* Bootstrap resample (enstop/enstop_.py:86-88, B = A[bootstrap_sample_indices]): the
 * working corpus becomes rows `row_idx[0..n_rows)` of the uploaded one, gathered on the
 * device.  row_idx == NULL restores the uploaded corpus. */
int plsa_bootstrap(plsa_ctx *ctx, const int32_t *row_idx, int64_t n_rows);
int plsa_corpus_shape(const plsa_ctx *ctx, int64_t *n_docs, int64_t *n_terms, int64_t *nnz);

/* ---- model state -------------------------------------------------------------------- */
/* p_z_given_d [n_docs, k], p_w_given_z [k, n_terms] (plsa.py:709-710 float32 C-order). */
int plsa_set_factors(plsa_ctx *ctx, const float *p_z_given_d, const float *p_w_given_z,
                     int32_t k);
/* Page-locked host memory owned by the context for the two initial factors ([n_docs, k] and
 * [k, n_terms] float32; valid until the next call with larger sizes or the context's end):
 * a host-side initialisation that writes into it skips the page faults of fresh memory and
 * plsa_set_factors copies from it at full PCIe speed.  Safe to call while another thread is
 * inside plsa_upload_csr* / plsa_prepare on the same context. */
int plsa_pinned_factors(plsa_ctx *ctx, int64_t n_docs, int64_t n_terms, int32_t k,
                        float **p_z_given_d, float **p_w_given_z);
/* Page-locked host memory for the caller's own buffers (cudaHostAlloc, portable): an upload
 * whose source lies in it is one DMA transfer, without the staging copy pageable memory needs. */
int plsa_host_alloc(int64_t bytes, void **ptr);
int plsa_host_free(void *ptr);
/* sample_weight [n_docs] or NULL for all ones (plsa.py:1144). */
int plsa_set_sample_weight(plsa_ctx *ctx, const float *sample_weight);
/* Either output may be NULL. */
int plsa_get_factors(plsa_ctx *ctx, float *p_z_given_d, float *p_w_given_z);
/* ---- EM ------------------------------------------------------------------------------- */
/* The loop of plsa_fit_inner (refit == 0; plsa.py:583-640) or plsa_refit_inner
 * (refit != 0; plsa.py:884-920) on the resident corpus and factors.
 *   use_sample_weights  plsa.py:606 — P(w|z) accumulates s*sample_weight[d]
 *   iters_run           EM iterations actually carried out (early stop, plsa.py:635)
 *   ll_trace/ll_cap     optional: log-likelihood before the loop, then every tested value
 *   n_ll                number of entries the trace would hold */
int plsa_em(plsa_ctx *ctx, int32_t n_iter, int32_t n_iter_per_test, double tolerance,
            float e_step_thresh, int32_t refit, int32_t use_sample_weights,
            int32_t *iters_run, double *ll_trace, int32_t ll_cap, int32_t *n_ll);
/* Build what a later plsa_em needs from the corpus alone (work items; for a full fit also
 * the term-major copy; items are sized for k topics) — lets the host overlap its RNG
 * initialisation with it. */
int plsa_prepare(plsa_ctx *ctx, int32_t refit, int32_t k);
/* log_likelihood (plsa.py:329-386) of the resident model, float64 reduction. */
int plsa_log_likelihood(plsa_ctx *ctx, double *ll);

/* ---- measurement ------------------------------------------------------------------------ */
/* Device time of the last plsa_em loop (CUDA events on the context's stream). */
int plsa_last_em_ms(const plsa_ctx *ctx, float *ms);
/* With profiling on, every kernel launch inside plsa_em is bracketed by CUDA events on
 * the context's stream.  Slots: see PLSA_PROF_* . */
#define PLSA_PROF_DOC_PASS 0   /* E-step + P(z|d) M-step over the doc-major copy        */
#define PLSA_PROF_WORD_PASS 1  /* E-step + P(w|z) M-step + column sums, term-major copy */
#define PLSA_PROF_FIXUP 2      /* ordered sums of split rows (both factors)            */
#define PLSA_PROF_NORMALIZE 3  /* sharded fit: all-reduce of P(w|z) + its column sums   */
#define PLSA_PROF_LOGLIK 4     /* log-likelihood pass                                  */
#define PLSA_PROF_DOC_HEAD 5   /* tiled doc pass: the shared-memory tile kernel alone (also
                                  counted in PLSA_PROF_DOC_PASS)                        */
#define PLSA_PROF_TERM_HEAD 6  /* tiled term pass: the tile kernel alone (also counted in
                                  PLSA_PROF_WORD_PASS)                                  */
#define PLSA_PROF_SLOTS 7
int plsa_set_profiling(plsa_ctx *ctx, int32_t on);
int plsa_get_profile(plsa_ctx *ctx, double *ms /*[PLSA_PROF_SLOTS]*/,
                     int64_t *launches /*[PLSA_PROF_SLOTS]*/);
/* Kernel launches issued by this context since creation (bench "gpu_launches"). */
int plsa_launch_count(const plsa_ctx *ctx, int64_t *launches);
/* Tunables: "chunk" (max stored entries per work item, 32..4096, 0 = automatic; longer rows
 * are split),
 * "texture" (1: gather factor rows through the texture pipe when they fit, 0: plain loads),
 * "fuse_ll" (1: the periodic log-likelihood rides on the next doc pass, 0: separate pass),
 * "tiled" (-1: automatic by corpus size, 0: never, 1: wherever possible — the doc pass reads
 * the rows of the most frequent terms from a TMA-staged shared-memory tile, csrc/plsa_tile.cuh),
 * "tile_kb" (shared memory of that tile per CTA, 1..220),
 * "presort" (1: plsa_upload_csr starts the sort of the entries by term — needed by a full fit,
 * not by a refit — on a second stream as soon as the column indices have arrived, while the
 * values are still being copied; default 0),
 * "p2p_timeout_ms" (sharded fit: bound of a rank's wait for a peer's partial sums),
 * "p2p_two_shot" (sharded fit over peer memory: -1 = from 4 ranks up every rank adds only its
 * slice of the terms and the finished slices are exchanged, 0 = every rank adds everything,
 * 1 = always slices; same bits either way). */
int plsa_set_option(plsa_ctx *ctx, const char *name, int64_t value);
/* Host-only (no device): the work items a pass over a CSR with these row pointers would launch,
 * in launch order.  A row longer than `chunk` entries is cut into equal chunks that write
 * partial sums into consecutive slots; with `align` > 1 an item starts on a multiple of
 * `align` entries and its first `skip` entries belong to the row before.  Writes at most
 * `cap` items into the optional arrays; *n_items is the full count. */
int plsa_plan_items(const int32_t *indptr, int64_t rows, int64_t chunk, int32_t align,
                    int64_t cap, int64_t *start, int32_t *row, int32_t *len,
                    int32_t *slot, int32_t *skip, int64_t *n_items, int32_t *n_split,
                    int32_t *n_slots);

/* Test hook: the work items the context holds on the device (which: 0 doc pass, 1 term pass,
 * 2 tail part of the tiled doc pass) and the row pointers they were planned from. */
int plsa_debug_items(plsa_ctx *ctx, int32_t which, int64_t cap, int64_t *start, int32_t *row,
                     int32_t *len, int32_t *slot, int32_t *skip, int64_t *n_items, int32_t *n_split,
                     int32_t *n_slots, int64_t *chunk, int32_t *align, int32_t *indptr_out,
                     int64_t indptr_cap);

/* ---- host helper: the reference's seeded random initialisation, faster ------------------------ */
/* plsa.py:454-456 + :510-511 + :709-710: draw rows*cols doubles from a numpy legacy
 * RandomState (MT19937 state key[624], *pos — from rng.get_state(), written back for
 * rng.set_state()), L1-normalise each row in float64 (utils.py:22-41) and store float32
 * (and optionally the float64 values).  Bit-identical to rng.rand + normalize + astype. */
int plsa_host_random_rows(uint32_t *key, int32_t *pos, int64_t rows, int64_t cols, float *out,
                          double *out_f64);

/* ---- one-shot drop-ins for the reference's raw-array seam --------------------------------- */
/* plsa.py:516-640.  p_w_given_z [k, m] and p_z_given_d [n, k] are updated in place. */
int plsa_b200_fit_inner(const int32_t *X_rows, const int32_t *X_cols, const float *X_vals,
                        int64_t nnz, float *p_w_given_z, float *p_z_given_d,
                        const float *sample_weight, int64_t n_docs, int64_t n_terms,
                        int32_t k, int32_t n_iter, int32_t n_iter_per_test,
                        double tolerance, float e_step_thresh, int32_t use_sample_weights,
                        int32_t device, int32_t *iters_run);
/* plsa.py:819-920.  topics [k, m] is read only; p_z_given_d [n, k] is updated in place. */
int plsa_b200_refit_inner(const int32_t *X_rows, const int32_t *X_cols, const float *X_vals,
                          int64_t nnz, const float *topics, float *p_z_given_d,
                          const float *sample_weight, int64_t n_docs, int64_t n_terms,
                          int32_t k, int32_t n_iter, int32_t n_iter_per_test,
                          double tolerance, float e_step_thresh, int32_t device,
                          int32_t *iters_run);

/* ---- ensemble: topic stash + gather (enstop_.py:209-231) ------------------------------------ */
/* plsa_topics (enstop_.py:56-115) returns only P(w|z); ensemble_of_topics stacks the
 * members' results with np.vstack (enstop_.py:231).  Here every finished member leaves its
 * P(w|z) [k, n_terms] (reference layout, float32) in slot `slot` of an n_slots-deep
 * device-side stash owned by the context, and one gather moves all of them to the host. */
int plsa_stash_topics(plsa_ctx *ctx, int32_t slot, int32_t n_slots);
/* The stash as a DEVICE pointer (valid until the next stash/destroy on this context). */
int plsa_topics_device(plsa_ctx *ctx, void **device_ptr, int64_t *floats_per_slot);

/* Single process, one context per device: slots [0, n_slots[i]) of ctxs[i] are sent to
 * ctxs[0]'s device with NCCL send/recv over NVLink (ncclCommInitAll) and copied to `out`
 * (host, [sum(n_slots) * k, n_terms], context order then slot order).  One context, or no
 * remote slots, never touches NCCL. */
int plsa_gather_topics(plsa_ctx **ctxs, int32_t n_ctx, const int32_t *n_slots, float *out);
/* Create (and keep, per device list) the communicators plsa_gather_topics will use, e.g. from
 * a helper thread while the members are still being fitted: ncclCommInitAll takes longer
 * than an ensemble member. */
int plsa_gather_warmup(const int32_t *devices, int32_t n_devices);
/* Append slots [0, n_src) of src's stash to slots [0, n_dst) of dst's (both on one device):
 * lets several contexts share a GPU during the fits and still gather one stash per device. */
int plsa_stash_append(plsa_ctx *dst, plsa_ctx *src, int32_t n_dst, int32_t n_src);

/* One process per GPU (torchrun-style launch).  Rank 0 obtains a unique id, the caller's
 * own rendezvous hands its PLSA_NCCL_ID_BYTES bytes to every rank, each rank creates a
 * communicator; plsa_comm_gather_topics sends slots [0, n_per_rank[rank]) of ctx's stash to
 * `root`, which receives them in rank order into `out` (host; ignored on other ranks). */
#define PLSA_NCCL_ID_BYTES 128
int plsa_nccl_unique_id(char *id /*[PLSA_NCCL_ID_BYTES]*/);
int plsa_comm_create(int device, int32_t n_ranks, int32_t rank,
                     const char *id /*[PLSA_NCCL_ID_BYTES]*/, plsa_comm **comm);
int plsa_comm_destroy(plsa_comm *comm);
/* Abort outstanding collectives (ncclCommAbort): lets the surviving ranks of a failed sharded
 * fit return an error instead of waiting.  plsa_comm_destroy must still follow. */
int plsa_comm_abort(plsa_comm *comm);
int plsa_comm_gather_topics(plsa_comm *comm, plsa_ctx *ctx, const int32_t *n_per_rank,
                            int32_t root, float *out);

/* ---- ensemble: all-pairs distances between the stacked topics (enstop_.py:234-263) ------------
 * topics [n_topics, n_terms] float32 (the np.vstack of the members' P(w|z)); out [n_topics,
 * n_topics] float64.  kind 0: Hellinger distance (umap.distances.hellinger as used by
 * all_pairs_hellinger_distance, enstop_.py:253-263; rows need not sum to 1; an all-zero row is
 * at distance 1 from every non-zero row and 0 from another all-zero row).  kind 1:
 * all_pairs_kl_divergence (enstop_.py:234-250), out[i, j] = sum over terms where both are
 * positive of a log2(a / b).  The reference evaluates both serially in O(N^2 m). */
#define PLSA_DIST_HELLINGER 0
#define PLSA_DIST_KL 1
int plsa_topic_distances(int32_t device, const float *topics, int64_t n_topics, int64_t n_terms,
                         int32_t kind, double *out);
/* The same on the stack of topic matrices the last plsa_gather_topics / plsa_comm_gather_topics
 * left on this (root) context's device: the clustering stage of the ensemble
 * (enstop_.py:266-351) gets its distance matrix without uploading the stack again.
 * out == NULL: only *n_topics is returned. */
int plsa_gathered_distances(plsa_ctx *ctx, int32_t kind, double *out, int64_t *n_topics);
/* Device time (CUDA events) of the kernels of this thread's last distance call. */
int plsa_last_distances_ms(float *kernel_ms);

/* ---- one fit over several GPUs: documents sharded by rows -------------------------------------
 * Generalises the row blocking of enstop/block_parallel_plsa.py:156-185 and
 * enstop/distributed_plsa.py:116-131 (per-block E-step + partial M-step sums, then a sum over
 * the blocks).  Every rank uploads ITS rows of X as its corpus (plsa_upload_csr*), sets
 * P(z|d) for its rows and the same full P(w|z) (plsa_set_factors), attaches the communicator
 * and runs plsa_em with identical arguments: the doc pass is local, the term pass leaves the
 * shard's raw P(w|z) sums, which are added over the ranks once per EM iteration (NCCL
 * all-reduce over NVLink, issued behind the term pass so that it overlaps the doc pass); the
 * log-likelihood is the sum of the shards' values, so every rank takes the same early-stop
 * decision (plsa.py:630-638).  plsa_get_factors returns the rank's rows of P(z|d) and the
 * full P(w|z).  comm == NULL detaches.  Every shard must hold at least one document. */
int plsa_set_shard(plsa_ctx *ctx, plsa_comm *comm);
/* Peer-memory all-reduce (optional; without it plsa_em uses ncclAllReduce).  After
 * plsa_set_shard and plsa_set_factors every rank (1) prepares its exchange block — two
 * buffers for its raw P(w|z) sums and a row of signal words, one allocation, returned as
 * a device address — (2) hands the peers either that address (same process; the library
 * enables peer access) or the 64-byte CUDA IPC handle from plsa_shard_p2p_export (other
 * processes), and (3) attaches every peer's block.  Once ALL ranks have attached ALL peers
 * (the caller's barrier), plsa_em adds the ranks' sums with one kernel per rank that reads
 * the peers' buffers over NVLink, in rank order (bit-identical on every rank), and takes the
 * column sums in the same pass.  Cross-GPU ordering uses signal words in peer memory with a
 * bounded wait: a missing peer yields PLSA_ENCCL, not a hang.  Option "p2p" = 0 forces NCCL. */
#define PLSA_IPC_HANDLE_BYTES 64
int plsa_shard_p2p_prepare(plsa_ctx *ctx, uint64_t *device_address, int64_t *bytes);
int plsa_shard_p2p_export(plsa_ctx *ctx, char *handle /*[PLSA_IPC_HANDLE_BYTES]*/);
int plsa_shard_p2p_attach(plsa_ctx *ctx, int32_t peer_rank, int32_t peer_device,
                          uint64_t device_address, const char *ipc_handle /* or NULL */);
/* Unmap the peers' blocks; call on every rank, then a barrier, before any rank frees its own
 * block (plsa_set_shard(ctx, NULL) or plsa_ctx_destroy). */
int plsa_shard_p2p_detach(plsa_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* PLSA_B200_H */
