/*
 * plsa_b200.cu — host side of libplsa_b200.so: the C ABI declared in include/plsa_b200.h.
 *
 * Replaces the raw-array seam of the reference (enstop/plsa.py:516-640 plsa_fit_inner,
 * :819-920 plsa_refit_inner) and keeps a corpus resident on one B200 for repeated fits
 * (ensemble members, transform).  Device code is in plsa_kernels.cuh.
 *
 * Data layout in HBM (per context):
 *   doc-major CSR   indptr[n+1] i32, ent[nnz] {term i32, count f32}     (working corpus)
 *   term-major CSR  ent[nnz] {doc i32, count f32} (+ a sample-weighted copy), built on device
 *   P(z|d)          A[2][n, strideA] f32, ping-pong
 *   P(w|z)^T        B[2][m, strideB] f32, ping-pong, RAW column-unnormalised sums;
 *                   scale[kp] = 1 / column sum is folded in by the next pass
 *   work items      one per row, rows longer than `chunk` entries are split
 */
#include <cub/cub.cuh>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <sched.h>

#include <algorithm>
#include <set>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/plsa_b200.h"
#include "plsa_kernels.cuh"
#include "plsa_tile.cuh"

using namespace plsa;

#define API extern "C" __attribute__((visibility("default")))

static thread_local std::string g_err;

namespace {

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        /* a buffer that has to grow gets 1/16 of slack: bootstrapped ensemble members differ
         * by a fraction of a percent in size, and cudaFree + cudaMalloc of tens of megabytes
         * (with its device-wide synchronisation) per member costs more than the kernels */
        const size_t want = p ? bytes + bytes / 16 : bytes;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMalloc(&p, want ? want : 16);
        if (e != cudaSuccess && want > bytes) {
            cudaGetLastError();
            e = cudaMalloc(&p, bytes);
            if (e == cudaSuccess) cap = bytes;
            return e;
        }
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

struct Corpus {
    int64_t n = 0, m = 0, nnz = 0;
    DevBuf indptr, ent; /* ent: int2 {column, value bits} [nnz + ENT_SLACK], slack zeroed */
    std::vector<int32_t> h_indptr;
};

/* readable, zeroed entries past the end: kernels read ahead of a row's end, the group-per-row
 * kernel by up to one work-item length */
constexpr int64_t ENT_PAD = ENT_SLACK + 4096;
static inline size_t ent_bytes(int64_t nnz) { return (size_t)(nnz + ENT_PAD) * sizeof(int2); }

struct ItemSet {
    DevBuf items, split_rows, slot_begin;
    DevBuf plan[9]; /* device planner scratch: 5 counts, their 5 scans share [5..], keys, perm, unsorted */
    int64_t n_items = 0;
    int32_t n_split = 0, n_heavy = 0, n_slots = 0;
    bool ready = false;
    int64_t chunk = 0;
    int align = 1; /* items start on multiples of this many entries (Item::skip) */
    void release()
    {
        items.release(); split_rows.release(); slot_begin.release();
        for (DevBuf &b : plan) b.release();
        ready = false;
    }
};

/* Tiled doc pass (plsa_tile.cuh): the doc-major corpus split into a "head" CSR whose columns
 * are slots of the shared-memory tile (the most frequent terms) and a "tail" CSR with the
 * rest, which stays with the group-per-row kernel. */
struct TileSet {
    bool ready = false;
    int32_t kp = 0, tile_rows = 0, pitch_f = 0;
    int64_t n = 0, head_slots = 0, tail_nnz = 0;
    DevBuf col_count, col_ids, col_count_sorted, col_sorted, slot_of, head_len, head_mem, tail_len,
        head_indptr, tail_indptr, head_ent, tail_ent, order, order_keys, row_ids, cub_tmp, head_sum,
        img[2], scale_raw, ll_part, ll_head, ll_ticket, headers;
    std::vector<int32_t> h_tail_indptr;
    ItemSet tail_items;
    void release()
    {
        for (DevBuf *b : {&col_count, &col_ids, &col_count_sorted, &col_sorted, &slot_of, &head_len, &head_mem,
                          &tail_len, &head_indptr, &tail_indptr, &head_ent, &tail_ent, &order,
                          &order_keys, &row_ids, &cub_tmp, &head_sum, &img[0], &img[1], &scale_raw,
                          &ll_part, &ll_head, &ll_ticket, &headers})
            b->release();
        tail_items.release();
        ready = false;
    }
};

/* Tiled term pass: the frequent terms' rows of the term-major corpus cut at blocks of
 * documents; item v = (tiled term, block) gathers P(z|d) rows of ONE block, which a CTA holds
 * in shared memory; the per-block partial sums of a term are added by fixup_kernel.  The other
 * terms stay with the group-per-row kernel (tail_items skips the tiled rows). */
struct TermTiles {
    bool ready = false, weighted = false;
    int32_t kp = 0, block_rows = 0, n_blocks = 0, n_tiled = 0, pitch_f = 0, grid = 0;
    int64_t n = 0, m = 0, n_items = 0, head_slots = 0;
    DevBuf flag, tiled_at, tiled_terms, beg, end, own_row, head_len, head_mem, head_indptr, head_ent, keys,
        keys_sorted, ids, order, work, work_prefix, cta_begin, block_begin, partial, slot_begin, cub_tmp,
        acomp[2], headers;
    std::vector<int32_t> h_flag;
    ItemSet tail_items;
    void release()
    {
        for (DevBuf *b : {&flag, &tiled_at, &tiled_terms, &beg, &end, &own_row, &head_len, &head_mem, &head_indptr,
                          &head_ent, &keys, &keys_sorted, &ids, &order, &work, &work_prefix, &cta_begin,
                          &block_begin, &partial, &slot_begin, &cub_tmp, &acomp[0], &acomp[1], &headers})
            b->release();
        tail_items.release();
        ready = false;
    }
};

} // namespace

struct plsa_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::string err;

    Corpus base, boot;
    bool use_boot = false;
    Corpus &cur() { return use_boot ? boot : base; }
    const Corpus &cur() const { return use_boot ? boot : base; }

    /* term-major copy of the working corpus */
    DevBuf t_ent, t_entw, t_indptr, up_cols, up_vals, flag, scratch[7];
    bool device_plan = true; /* option "device_plan": work items planned on the device */
    std::vector<int32_t> h_tindptr;
    bool t_ready = false, t_weighted_ready = false;

    ItemSet doc_items, term_items;
    TileSet tiles;
    TermTiles tterm;
    int term_tiled_opt = 1;             /* option "term_tiled": tile the term pass too (with "tiled") */
    int64_t term_tile_min = 12;         /* option "term_tile_min": entries per (term, block) item, on average, for a term to be tiled */
    bool a_comp[2] = {false, false};    /* tterm.acomp[i] is the compact image of A[i] */
    int tiled_opt = -1;                 /* option "tiled": -1 by corpus size (tiled_wanted), 0 off, 1 on where possible */
    int64_t tile_bytes = 200 * 1024;    /* option "tile_kb": shared memory of the tile        */
    bool b_norm[2] = {false, false};    /* B[i] is column-normalised in place, tiles.img[i] holds its tile rows */
    int n_sms = 0;
    int64_t chunk_user = 0;  /* option "chunk": 0 = chosen per corpus size and k */
    int64_t chunk_built = 0; /* work-item length the item sets were built with */
    int32_t k_hint = 0;      /* plsa_prepare: k the items should be sized for */
    bool use_texture = true; /* gather through the texture pipe when the factor fits */
    bool vec_entries = true; /* items aligned to entry blocks, one wide load per block */
    bool fuse_ll = true;     /* take the periodic log-likelihood from the next doc pass */
    double *mail = nullptr;  /* pinned host mailbox {ll, flag} */
    cudaEvent_t ev_ll = nullptr;
    bool overlap = true;     /* doc pass and term pass of an iteration on two streams */
    cudaStream_t stream2 = nullptr;
    bool presort_opt = false;      /* option "presort": sort by term while the values upload */
    bool presorted = false;        /* the base corpus' term sort is queued on stream2 */
    cudaEvent_t ev_presort = nullptr, ev_cols = nullptr, ev_d2h = nullptr;
    /* pinned staging for host->device copies of pageable memory (see h2d_fast) */
    static constexpr int H2D_THREADS = 4; /* at most; h2d_threads() of them are used */
    static constexpr size_t H2D_CHUNK = (size_t)4 << 20;
    char *pin[H2D_THREADS] = {nullptr, nullptr, nullptr, nullptr};
    cudaStream_t pin_stream[H2D_THREADS] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t pin_ev[H2D_THREADS][2] = {};
    cudaEvent_t ev_a = nullptr, ev_b = nullptr, ev_go = nullptr;
    /* pinned host block the caller may draw the initial factors into (plsa_pinned_factors) */
    char *pin_factors = nullptr;
    size_t pin_factors_cap = 0;
    cudaTextureObject_t texA[2] = {0, 0}, texB[2] = {0, 0};
    std::map<cudaTextureObject_t *, std::pair<void *, size_t>> tex_bound; /* what each is bound to */
    size_t tex_max_texels = 0;

    /* model */
    int32_t k = 0, kp = 0, strideA = 0, strideB = 0;
    DevBuf A[2], B[2], scale, ones, colnorm, colpart, partialA, partialB;
    DevBuf sw, ll_part, ll_out, stage, topics_dev, tickets;
    DevBuf gathered;                /* stack of topic matrices the last gather brought to this device */
    int64_t gathered_n = 0, gathered_m = 0;
    int32_t stash_slots = 0;
    size_t stash_per = 0;
    int curA = 0, curB = 0;
    bool have_factors = false, have_sw = false;

    /* document-sharded fit: this context holds one shard of the documents; the ranks' raw
     * P(w|z) sums and log-likelihoods are added over `shard` (NCCL) inside plsa_em */
    plsa_comm *shard = nullptr;
    DevBuf ll2, colpart2;
    /* exchange block of the peer-memory all-reduce: [partial 0 | partial 1 | signal words] */
    struct P2P {
        DevBuf block, err;
        size_t part_bytes = 0;
        void *peer_base[SHARD_MAX_RANKS] = {};
        bool peer_ipc[SHARD_MAX_RANKS] = {};
        int n_attached = 0;
        unsigned int seq = 0;
        size_t red_bytes = 0;       /* one finished-slice buffer of the two-shot exchange */
        int two_shot = -1;          /* option "p2p_two_shot": -1 from 4 ranks up, 0 never, 1 always */
        bool enabled = true; /* option "p2p" */
        int64_t timeout_ms = 30000; /* option "p2p_timeout_ms": wait for a peer's signal */
    } p2p;

    /* measurement */
    float last_em_ms = 0.f;
    int64_t launches = 0;
    bool profiling = false;
    double prof_ms[PLSA_PROF_SLOTS] = {};
    int64_t prof_n[PLSA_PROF_SLOTS] = {};
    struct ProfRec { int slot; cudaEvent_t a, b; };
    std::vector<ProfRec> prof_pending;
    std::vector<cudaEvent_t> ev_pool;

    int fail(int code, const std::string &msg)
    {
        err = msg;
        g_err = msg;
        return code;
    }
};

#define CK(expr)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (expr);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return ctx->fail(e_ == cudaErrorMemoryAllocation ? PLSA_ENOMEM : PLSA_ECUDA,      \
                             std::string(#expr) + ": " + cudaGetErrorString(e_));            \
    } while (0)

#define CHECK_CTX(ctx)                                                                        \
    do {                                                                                      \
        if (!(ctx)) {                                                                         \
            g_err = "null context";                                                           \
            return PLSA_EINVAL;                                                               \
        }                                                                                     \
        cudaError_t e_ = cudaSetDevice((ctx)->device);                                        \
        if (e_ != cudaSuccess)                                                                \
            return (ctx)->fail(PLSA_ECUDA, std::string("cudaSetDevice: ") +                   \
                                               cudaGetErrorString(e_));                       \
    } while (0)

static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

/* ---- fast host -> device copy of pageable memory ------------------------------------------
 * cudaMemcpy from pageable memory is staged by the driver on one thread (~10 GB/s here).
 * Large uploads are instead cut into four slices, each moved by its own host thread through
 * a private pinned double buffer and stream, so the host-side memcpy runs four wide and
 * overlaps the DMA. */
/* Upload threads of this process: the cores it may run on, shared with the other ranks of the
 * box (LOCAL_WORLD_SIZE, set by torchrun) and with the fit's own helper threads (seeded init,
 * value check) — eight ranks with four upload threads each oversubscribed a 32-core host. */
static int h2d_threads()
{
    static const int n = [] {
        cpu_set_t set;
        int cores = 0;
        if (sched_getaffinity(0, sizeof(set), &set) == 0) cores = CPU_COUNT(&set);
        if (cores <= 0) cores = (int)std::thread::hardware_concurrency();
        int ranks = 1;
        if (const char *e = getenv("LOCAL_WORLD_SIZE")) ranks = std::max(1, atoi(e));
        if (const char *e = getenv("ENSTOP_B200_H2D_THREADS")) return std::max(1, std::min(4, atoi(e)));
        return std::max(1, std::min(plsa_ctx::H2D_THREADS, cores / ranks / 2));
    }();
    return n;
}

static cudaError_t h2d_fast(plsa_ctx *ctx, void *dst, const void *src, size_t bytes)
{
    const int T = h2d_threads();
    constexpr size_t CH = plsa_ctx::H2D_CHUNK;
    if (bytes < 8 * CH) return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream);
    cudaError_t e;
    {   /* a source that is already page-locked (plsa_host_alloc, cudaHostRegister) goes by DMA as
         * it is: no staging copy, no host threads */
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, src) == cudaSuccess && attr.type == cudaMemoryTypeHost)
            return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream);
        cudaGetLastError(); /* an ordinary pointer is not an error */
    }
    for (int t = 0; t < T; ++t) {
        if (!ctx->pin[t]) {
            if ((e = cudaHostAlloc((void **)&ctx->pin[t], 2 * CH, cudaHostAllocDefault)) != cudaSuccess ||
                (e = cudaStreamCreateWithFlags(&ctx->pin_stream[t], cudaStreamNonBlocking)) != cudaSuccess ||
                (e = cudaEventCreateWithFlags(&ctx->pin_ev[t][0], cudaEventDisableTiming)) != cudaSuccess ||
                (e = cudaEventCreateWithFlags(&ctx->pin_ev[t][1], cudaEventDisableTiming)) != cudaSuccess)
                return e;
        }
    }
    /* everything queued on ctx->stream so far (e.g. buffer memsets) precedes the copies */
    if ((e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess) return e;
    const size_t slice = (bytes / T + 255) / 256 * 256;
    cudaError_t errs[plsa_ctx::H2D_THREADS];
    std::thread th[plsa_ctx::H2D_THREADS];
    const int device = ctx->device;
    int started = 0;
    try { /* no C++ exception crosses the C ABI */
    for (int t = 0; t < T; ++t, ++started) {
        errs[t] = cudaSuccess;
        th[t] = std::thread([=, &errs]() {
            cudaSetDevice(device);
            const size_t lo = std::min(bytes, slice * (size_t)t), hi = std::min(bytes, lo + slice);
            int b = 0;
            for (size_t off = lo; off < hi; off += CH, b ^= 1) {
                const size_t n = std::min(CH, hi - off);
                cudaError_t r = cudaEventSynchronize(ctx->pin_ev[t][b]); /* buffer free again */
                if (r == cudaSuccess) {
                    memcpy(ctx->pin[t] + (size_t)b * CH, (const char *)src + off, n);
                    r = cudaMemcpyAsync((char *)dst + off, ctx->pin[t] + (size_t)b * CH, n,
                                        cudaMemcpyHostToDevice, ctx->pin_stream[t]);
                }
                if (r == cudaSuccess) r = cudaEventRecord(ctx->pin_ev[t][b], ctx->pin_stream[t]);
                if (r != cudaSuccess) {
                    errs[t] = r;
                    return;
                }
            }
            errs[t] = cudaStreamSynchronize(ctx->pin_stream[t]);
        });
    }
    } catch (...) {
    }
    for (int t = 0; t < started; ++t) th[t].join();
    if (started < T) return cudaErrorOperatingSystem; /* could not start a copy thread */
    for (int t = 0; t < T; ++t)
        if (errs[t] != cudaSuccess) return errs[t];
    return cudaSuccess;
}

/* ---- profiling helpers ------------------------------------------------------------------- */
static cudaEvent_t get_event(plsa_ctx *ctx)
{
    if (!ctx->ev_pool.empty()) {
        cudaEvent_t e = ctx->ev_pool.back();
        ctx->ev_pool.pop_back();
        return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}

struct ProfScope {
    plsa_ctx *ctx;
    int slot;
    cudaEvent_t a = nullptr, b = nullptr;
    cudaStream_t st;
    ProfScope(plsa_ctx *c, int s, cudaStream_t stream = nullptr)
        : ctx(c), slot(s), st(stream ? stream : c->stream)
    {
        if (ctx->profiling) {
            a = get_event(ctx);
            b = get_event(ctx);
            cudaEventRecord(a, st);
        }
    }
    ~ProfScope()
    {
        if (ctx->profiling) {
            cudaEventRecord(b, st);
            ctx->prof_pending.push_back({slot, a, b});
        }
    }
};

static void prof_collect(plsa_ctx *ctx)
{
    for (auto &r : ctx->prof_pending) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
            ctx->prof_ms[r.slot] += ms;
            ctx->prof_n[r.slot] += 1;
        }
        ctx->ev_pool.push_back(r.a);
        ctx->ev_pool.push_back(r.b);
    }
    ctx->prof_pending.clear();
}

/* ---- document-sharded fit: collectives over the shard communicator (defined with the NCCL
 * bindings at the end of this file) ------------------------------------------------------------ */
static int shard_allreduce(plsa_ctx *ctx, void *buf, size_t count, bool f64, cudaStream_t stream);
static int shard_n_ranks(const plsa_ctx *ctx);
static int shard_rank(const plsa_ctx *ctx);
static void p2p_release(plsa_ctx *ctx);

/* ---- kernel dispatch ------------------------------------------------------------------------ */
typedef void (*pass_fn)(const PassArgs);

template <int G, int KV, bool VEC> static pass_fn pick_mode_v(int mode, bool tex)
{
    switch (mode) {
    case MODE_DOC: return tex ? row_pass_kernel<G, KV, MODE_DOC, true, VEC> : row_pass_kernel<G, KV, MODE_DOC, false, VEC>;
    case MODE_TERM: return tex ? row_pass_kernel<G, KV, MODE_TERM, true, VEC> : row_pass_kernel<G, KV, MODE_TERM, false, VEC>;
    case MODE_DOC_LL: return tex ? row_pass_kernel<G, KV, MODE_DOC_LL, true, VEC> : row_pass_kernel<G, KV, MODE_DOC_LL, false, VEC>;
    default: return tex ? row_pass_kernel<G, KV, MODE_LOGLIK, true, VEC> : row_pass_kernel<G, KV, MODE_LOGLIK, false, VEC>;
    }
}

/* vec: the items start on entry-block boundaries (ItemSet::align > 1) */
template <int G, int KV> static pass_fn pick_mode(int mode, bool tex, bool vec)
{
    if constexpr (pass_block_entries(KV) > 1) {
        if (vec) return pick_mode_v<G, KV, true>(mode, tex);
    }
    return pick_mode_v<G, KV, false>(mode, tex);
}

/* lanes per work item for a factor row of kp floats */
static int pass_group_lanes(int kp)
{
    const int nv = kp / 4;
    return nv <= 8 ? nv : nv <= 16 ? 16 : 32;
}

/* entries per block of the kernel that serves kp (pass_block_entries of its KV) */
static int pass_entry_block(int kp)
{
    const int nv = kp / 4;
    return nv <= 32 ? 4 : nv <= 64 ? 2 : 1;
}

static pass_fn pick_kernel(int kp, int mode, bool tex, bool vec)
{
    const int nv = kp / 4; /* float4 vectors per factor row */
    switch (nv) {
    case 1: return pick_mode<1, 1>(mode, tex, vec);
    case 2: return pick_mode<2, 1>(mode, tex, vec);
    case 3: return pick_mode<3, 1>(mode, tex, vec);
    case 4: return pick_mode<4, 1>(mode, tex, vec);
    case 5: return pick_mode<5, 1>(mode, tex, vec);
    case 6: return pick_mode<6, 1>(mode, tex, vec);
    case 7: return pick_mode<7, 1>(mode, tex, vec);
    case 8: return pick_mode<8, 1>(mode, tex, vec);
    default: break;
    }
    if (nv <= 16) return pick_mode<16, 1>(mode, tex, vec);
    if (nv <= 32) return pick_mode<32, 1>(mode, tex, vec);
    if (nv <= 64) return pick_mode<32, 2>(mode, tex, vec);
    if (nv <= 128) return pick_mode<32, 4>(mode, tex, vec);
    return pick_mode<32, 8>(mode, tex, vec);
}

static int64_t pass_grid(int64_t n_items, int kp)
{
    return cdiv(n_items, (int64_t)(PLSA_PASS_THREADS / 32) * (32 / pass_group_lanes(kp))); /* 8 warps per CTA */
}

static int launch_pass(plsa_ctx *ctx, int mode, const PassArgs &a, bool vec,
                       cudaStream_t stream = nullptr)
{
    if (a.n_items == 0) return PLSA_OK;
    pass_fn fn = pick_kernel(a.kp, mode, ctx->use_texture && a.gat_tex != 0, vec);
    {   /* The pass lives on L1 hits of the gathered factor rows and uses next to no shared
         * memory (the term pass 5 KB per CTA for its column sums): ask for the smallest
         * shared-memory carve-out that still holds the resident CTAs (ncu showed 64 KB being
         * set aside for the term pass by default). */
        static std::mutex mu;
        static std::map<std::pair<int, pass_fn>, bool> configured; /* function attributes are per device */
        std::lock_guard<std::mutex> lock(mu);
        if (!configured[std::make_pair(ctx->device, fn)]) {
            cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout,
                                 mode == MODE_TERM ? 15 : 5);
            cudaGetLastError(); /* a hint: failure is not an error */
            configured[std::make_pair(ctx->device, fn)] = true;
        }
    }
    fn<<<(unsigned)pass_grid(a.n_items, a.kp), PLSA_PASS_THREADS, 0, stream ? stream : ctx->stream>>>(a);
    ctx->launches++;
    CK(cudaGetLastError());
    return PLSA_OK;
}

/* ---- work items --------------------------------------------------------------------------- */
/* One item per row; rows longer than `chunk` stored entries are cut into chunks whose
 * partial sums are added in order by fixup_kernel.  Items are ordered longest first
 * (counting sort) so that the 8 warps of a CTA carry rows of similar length and the
 * hardware CTA scheduler sees the heavy work first. */
/* Work-item length.  A group walks its item serially (U entries per ~0.45 us iteration), so
 * the item length bounds the pass from below; short items cost a header chain and a partial
 * sum each.  Measured on B200 (profiles/r1_kernel_experiments.md): 256 is best at 10 M
 * entries with 6 items per warp, 64 or less at 0.2 M entries, 2048 when a warp carries a
 * single item (k > 64). */
static int64_t choose_chunk(const plsa_ctx *ctx, int kp)
{
    if (ctx->chunk_user > 0) return ctx->chunk_user;
    const int64_t nnz = ctx->cur().nnz;
    const bool one_item_per_warp = kp > 0 && 32 / pass_group_lanes(kp) < 3;
    const int64_t cap = one_item_per_warp ? 2048 : 256;
    const int64_t want = nnz / (one_item_per_warp ? 4096 : 32768);
    return std::max<int64_t>(32, std::min(cap, (want + 31) / 32 * 32));
}

/* The plan of a pass, host side only (no CUDA): the items in launch order, the split rows
 * (heavy first) and the first partial-sum slot of each.  Same-length items stay in row order:
 * the chunks of a split row sit next to each other.  (Two other launch orders — chunks grouped
 * by their position inside the row, for L1 reuse inside a CTA, and in bands of positions, for
 * L2 reuse — were measured at C2 and gained nothing: profiles/r2a_ab_c2.txt.) */
struct ItemPlan {
    std::vector<Item> sorted;
    std::vector<int32_t> split_rows, slot_begin;
    int32_t n_heavy = 0, slots = 0;
};

static void plan_items(const int32_t *indptr, int64_t rows, const int64_t chunk_asked, int align,
                       ItemPlan &plan, const int32_t *skip_row = nullptr /* rows that get no item */)
{
    const int64_t chunk = std::max<int64_t>(align, chunk_asked / align * align); /* chunks of a split row stay aligned */
    std::vector<Item> items;
    items.reserve((size_t)rows + 1024);
    /* align > 1: an item starts on a multiple of `align` entries at or before its first
     * entry; the entries in between (Item::skip of them, they belong to the row before) are
     * walked with value 0.  Lengths below include them. */
    auto lead = [&](int64_t r) { return (int64_t)indptr[r] & (int64_t)(align - 1); };
    auto skipped = [&](int64_t r) { return skip_row != nullptr && skip_row[r] != 0; };
    auto span = [&](int64_t r) {
        return skipped(r) ? (int64_t)0 : (int64_t)indptr[r + 1] - indptr[r] + lead(r);
    };
    /* A split row is cut into equal pieces (not full chunks plus a short remainder), so that
     * the pieces of a row sort next to each other and none of them is a tiny item. */
    auto piece = [&](int64_t len) {
        const int64_t nc = cdiv(len, chunk);
        return std::min(chunk, cdiv(cdiv(len, nc), (int64_t)align) * align);
    };
    /* split rows, those with more than 32 chunks first (fixup_kernel gives them a CTA) */
    std::vector<int64_t> heavy, light;
    for (int64_t r = 0; r < rows; ++r) {
        const int64_t len = span(r);
        if (len > chunk) (cdiv(len, piece(len)) > 32 ? heavy : light).push_back(r);
    }
    std::vector<int32_t> &split_rows = plan.split_rows, &slot_begin = plan.slot_begin;
    split_rows.clear();
    slot_begin.clear();
    slot_begin.push_back(0);
    std::vector<int32_t> first_slot((size_t)rows, -1);
    int32_t slots = 0;
    for (const std::vector<int64_t> *lst : {&heavy, &light})
        for (int64_t r : *lst) {
            first_slot[(size_t)r] = slots;
            slots += (int32_t)cdiv(span(r), piece(span(r)));
            split_rows.push_back((int32_t)r);
            slot_begin.push_back(slots);
        }
    for (int64_t r = 0; r < rows; ++r) {
        const int64_t skip = lead(r), s = indptr[r] - skip, len = span(r);
        if (skipped(r)) continue;
        if (len <= chunk) {
            items.push_back(Item{s, (int32_t)r, (int32_t)len, -1, (int32_t)skip});
        } else {
            const int64_t per = piece(len), nc = cdiv(len, per); /* per is a multiple of align */
            for (int64_t c = 0; c < nc; ++c) {
                const int64_t b = c * per;
                const Item it{s + b, (int32_t)r, (int32_t)std::min(per, len - b),
                              first_slot[(size_t)r] + (int32_t)c,
                              (int32_t)(c == 0 ? (skip | ITEM_FIRST) : 0)};
                items.push_back(it);
            }
        }
    }
    /* counting sort by length, descending, stable */
    std::vector<int64_t> cnt((size_t)chunk + 2, 0);
    for (const Item &it : items) cnt[(size_t)(chunk - it.len)]++;
    int64_t run = 0;
    for (auto &c : cnt) {
        const int64_t t = c;
        c = run;
        run += t;
    }
    std::vector<Item> &sorted = plan.sorted;
    sorted.resize(items.size());
    for (const Item &it : items) sorted[(size_t)cnt[(size_t)(chunk - it.len)]++] = it;
    plan.n_heavy = (int32_t)heavy.size();
    plan.slots = slots;
}

static int build_items(plsa_ctx *ctx, const std::vector<int32_t> &indptr, int64_t rows,
                       ItemSet &out, const int64_t chunk_asked, int align,
                       cudaStream_t stream = nullptr, const int32_t *h_skip_row = nullptr)
{
    if (!stream) stream = ctx->stream;
    ItemPlan plan;
    try { /* no C++ exception crosses the C ABI */
        plan_items(indptr.data(), rows, chunk_asked, align, plan, h_skip_row);
    } catch (const std::bad_alloc &) {
        return ctx->fail(PLSA_ENOMEM, "work items: out of host memory");
    }
    const std::vector<Item> &sorted = plan.sorted;
    const std::vector<int32_t> &split_rows = plan.split_rows, &slot_begin = plan.slot_begin;
    const int32_t slots = plan.slots;

    out.n_items = (int64_t)sorted.size();
    out.n_split = (int32_t)split_rows.size();
    out.n_heavy = plan.n_heavy;
    out.n_slots = slots;
    CK(out.items.ensure(sorted.size() * sizeof(Item)));
    CK(cudaMemcpyAsync(out.items.p, sorted.data(), sorted.size() * sizeof(Item),
                       cudaMemcpyHostToDevice, stream));
    CK(out.split_rows.ensure(split_rows.size() * sizeof(int32_t)));
    CK(out.slot_begin.ensure(slot_begin.size() * sizeof(int32_t)));
    if (!split_rows.empty())
        CK(cudaMemcpyAsync(out.split_rows.p, split_rows.data(),
                           split_rows.size() * sizeof(int32_t), cudaMemcpyHostToDevice,
                           stream));
    CK(cudaMemcpyAsync(out.slot_begin.p, slot_begin.data(), slot_begin.size() * sizeof(int32_t),
                       cudaMemcpyHostToDevice, stream));
    CK(cudaStreamSynchronize(stream)); /* host vectors die here */
    out.ready = true;
    out.chunk = chunk_asked;
    out.align = align;
    return PLSA_OK;
}

/* The same plan built on the device from the device copy of the row pointers: no host pass over
 * the rows, one 20-byte read-back (the totals).  Bit-identical to plan_items: the items are
 * generated in row order and put longest first by a STABLE radix sort, which is what the host's
 * counting sort does (tests/test_gpu_parity.py::test_device_plan_is_the_host_plan). */
static int build_items_device(plsa_ctx *ctx, const int32_t *d_indptr, int64_t rows, ItemSet &out,
                              const int64_t chunk_asked, int align, cudaStream_t stream = nullptr,
                              const int32_t *d_skip_row = nullptr /* rows that get no item */)
{
    if (!stream) stream = ctx->stream;
    const int32_t chunk = (int32_t)std::max<int64_t>(align, chunk_asked / align * align);
    const int T = 256;
    const size_t cnt_bytes = (size_t)(rows + 1) * 4;
    DevBuf &counts = out.plan[0], &scans = out.plan[1], &keys = out.plan[2], &keys_out = out.plan[3],
           &perm_in = out.plan[4], &perm_out = out.plan[5], &unsorted = out.plan[6], &tmp = out.plan[7];
    CK(counts.ensure(cnt_bytes * 5));
    CK(scans.ensure(cnt_bytes * 5));
    int32_t *c0 = counts.as<int32_t>(), *s0 = scans.as<int32_t>();
    const size_t stride = (size_t)rows + 1;
    plan_count_kernel<<<(unsigned)cdiv(rows + 1, T), T, 0, stream>>>(d_indptr, rows, chunk, align, d_skip_row, c0, c0 + stride,
                                                                   c0 + 2 * stride, c0 + 3 * stride,
                                                                   c0 + 4 * stride);
    size_t tmp_bytes = 0;
    CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, c0, s0, (int)(rows + 1), stream));
    CK(tmp.ensure(tmp_bytes));
    for (int q = 0; q < 5; ++q) {
        size_t tb = tmp.cap;
        CK(cub::DeviceScan::ExclusiveSum(tmp.p, tb, c0 + q * stride, s0 + q * stride, (int)(rows + 1), stream));
    }
    int32_t totals[5] = {0, 0, 0, 0, 0};
    for (int q = 0; q < 5; ++q)
        CK(cudaMemcpyAsync(&totals[q], s0 + q * stride + rows, 4, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    const int64_t n_items = totals[0];
    const int32_t n_heavy = totals[1], n_light = totals[2], slots = totals[3] + totals[4];
    out.n_items = n_items;
    out.n_split = n_heavy + n_light;
    out.n_heavy = n_heavy;
    out.n_slots = slots;
    CK(out.items.ensure((size_t)std::max<int64_t>(n_items, 1) * sizeof(Item)));
    CK(unsorted.ensure((size_t)std::max<int64_t>(n_items, 1) * sizeof(Item)));
    CK(keys.ensure((size_t)std::max<int64_t>(n_items, 1) * 4));
    CK(keys_out.ensure((size_t)std::max<int64_t>(n_items, 1) * 4));
    CK(perm_in.ensure((size_t)std::max<int64_t>(n_items, 1) * 4));
    CK(perm_out.ensure((size_t)std::max<int64_t>(n_items, 1) * 4));
    CK(out.split_rows.ensure((size_t)std::max(out.n_split, 1) * 4));
    CK(out.slot_begin.ensure((size_t)(out.n_split + 1) * 4));
    CK(cudaMemsetAsync(out.slot_begin.p, 0, 4, stream)); /* rows == 0: the closing entry */
    if (rows > 0 && n_items > 0) {
        plan_emit_kernel<<<(unsigned)cdiv(rows, T), T, 0, stream>>>(
            d_indptr, rows, chunk, align, d_skip_row, s0, s0 + stride, s0 + 2 * stride, s0 + 3 * stride, s0 + 4 * stride,
            unsorted.as<Item>(), keys.as<int32_t>(), out.split_rows.as<int32_t>(), out.slot_begin.as<int32_t>());
        iota_kernel<<<(unsigned)cdiv(n_items, T), T, 0, stream>>>(perm_in.as<int32_t>(), n_items);
        int end_bit = 1;
        while (((int64_t)1 << end_bit) <= chunk) ++end_bit;
        size_t sb = 0;
        CK(cub::DeviceRadixSort::SortPairs(nullptr, sb, keys.as<int32_t>(), keys_out.as<int32_t>(),
                                           perm_in.as<int32_t>(), perm_out.as<int32_t>(), (int)n_items, 0,
                                           end_bit, stream));
        CK(tmp.ensure(sb));
        sb = tmp.cap;
        CK(cub::DeviceRadixSort::SortPairs(tmp.p, sb, keys.as<int32_t>(), keys_out.as<int32_t>(),
                                           perm_in.as<int32_t>(), perm_out.as<int32_t>(), (int)n_items, 0,
                                           end_bit, stream));
        plan_gather_kernel<<<(unsigned)cdiv(n_items, T), T, 0, stream>>>(unsorted.as<Item>(), perm_out.as<int32_t>(),
                                                                       n_items, out.items.as<Item>());
        ctx->launches += 4;
        CK(cudaGetLastError());
    }
    out.ready = true;
    out.chunk = chunk_asked;
    out.align = align;
    return PLSA_OK;
}

/* ---- term-major copy ---------------------------------------------------------------------- */
/* Stable radix sort of the entry numbers by column: within a term the documents stay in
 * ascending order, so the summation order — and the result — is reproducible. */
/* First half of the term-major build: the stable sort of the stored entries by term — keys and
 * permutation only, no values — and the term row pointers.  `cols` (the uploaded column
 * indices, still in their staging buffer) lets plsa_upload_csr start it on `s` = the second
 * stream as soon as the indices have arrived, while the values are still crossing PCIe
 * (option "presort"); otherwise the keys are read from the interleaved entries. */
static int term_major_sort(plsa_ctx *ctx, const Corpus &c, const int32_t *cols, cudaStream_t s)
{
    const int64_t nnz = c.nnz, n = c.n, m = c.m;
    /* scratch (20 B per entry) is kept with the context for the next corpus unless it is
     * large: repeated fits then skip seven cudaMalloc/cudaFree pairs */
    DevBuf &rows_exp = ctx->scratch[0], &keys_in = ctx->scratch[1], &perm_in = ctx->scratch[2],
           &perm_out = ctx->scratch[3], &keys_out = ctx->scratch[4], &tindptr = ctx->t_indptr,
           &tmp = ctx->scratch[6];
    const size_t nz = (size_t)std::max<int64_t>(nnz, 1);
    CK(tindptr.ensure((size_t)(m + 1) * 4));
    if (nnz == 0) {
        CK(cudaMemsetAsync(tindptr.p, 0, (size_t)(m + 1) * 4, s));
        return PLSA_OK;
    }
    CK(rows_exp.ensure(nz * 4));
    CK(keys_in.ensure(nz * 4));
    CK(perm_in.ensure(nz * 4));
    CK(perm_out.ensure(nz * 4));
    CK(keys_out.ensure(nz * 4));
    const int T = 256;
    expand_rows_kernel<<<(unsigned)cdiv(n * 32, T), T, 0, s>>>(
        c.indptr.as<int32_t>(), n, cols ? nullptr : c.ent.as<int2>(), cols, rows_exp.as<int32_t>(),
        keys_in.as<int32_t>());
    iota_kernel<<<(unsigned)cdiv(nnz, T), T, 0, s>>>(perm_in.as<int32_t>(), nnz);
    ctx->launches += 2;
    int end_bit = 1;
    while (((int64_t)1 << end_bit) < m) ++end_bit;
    size_t tmp_bytes = 0;
    CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys_in.as<int32_t>(),
                                       keys_out.as<int32_t>(), perm_in.as<int32_t>(),
                                       perm_out.as<int32_t>(), (int)nnz, 0, end_bit, s));
    CK(tmp.ensure(tmp_bytes));
    CK(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, keys_in.as<int32_t>(),
                                       keys_out.as<int32_t>(), perm_in.as<int32_t>(),
                                       perm_out.as<int32_t>(), (int)nnz, 0, end_bit, s));
    lower_bound_kernel<<<(unsigned)cdiv(m + 1, T), T, 0, s>>>(keys_out.as<int32_t>(), nnz, m,
                                                             tindptr.as<int32_t>());
    ctx->launches++;
    CK(cudaGetLastError());
    return PLSA_OK;
}

static void release_sort_scratch(plsa_ctx *ctx, int64_t nnz)
{
    if (nnz > ((int64_t)32 << 20))
        for (DevBuf &b : ctx->scratch) b.release();
}

static int build_term_major(plsa_ctx *ctx)
{
    Corpus &c = ctx->cur();
    const int64_t nnz = c.nnz, m = c.m;
    cudaStream_t s = ctx->stream;
    /* the sort may already be under way (or done) on the second stream: plsa_upload_csr */
    const bool presorted = ctx->presorted && !ctx->use_boot;
    ctx->presorted = false;
    int rc = PLSA_OK;
    if (presorted) {
        cudaError_t e = cudaStreamWaitEvent(s, ctx->ev_presort, 0);
        if (e != cudaSuccess) rc = ctx->fail(PLSA_ECUDA, std::string("presort wait: ") + cudaGetErrorString(e));
    } else {
        if (ctx->ev_presort) cudaStreamWaitEvent(s, ctx->ev_presort, 0); /* same scratch buffers */
        rc = term_major_sort(ctx, c, nullptr, s);
    }
    auto failed = [&](cudaError_t e, const char *what) {
        release_sort_scratch(ctx, nnz);
        return ctx->fail(e == cudaErrorMemoryAllocation ? PLSA_ENOMEM : PLSA_ECUDA,
                         std::string(what) + ": " + cudaGetErrorString(e));
    };
    if (rc != PLSA_OK) {
        release_sort_scratch(ctx, nnz);
        return rc;
    }
    cudaError_t e;
    if ((e = ctx->t_ent.ensure(ent_bytes(nnz))) != cudaSuccess) return failed(e, "term-major entries");
    if ((e = cudaMemsetAsync(ctx->t_ent.as<int2>() + nnz, 0, ENT_PAD * sizeof(int2), s)) != cudaSuccess)
        return failed(e, "term-major padding");
    ctx->h_tindptr.assign((size_t)m + 1, 0);
    if (nnz > 0) {
        const int T = 256;
        permute_kernel<<<(unsigned)cdiv(nnz, T), T, 0, s>>>(ctx->scratch[3].as<int32_t>(), nnz,
                                                           ctx->scratch[0].as<int32_t>(),
                                                           c.ent.as<int2>(), ctx->t_ent.as<int2>());
        ctx->launches++;
        if ((e = cudaGetLastError()) != cudaSuccess) return failed(e, "permute_kernel");
        if ((e = cudaMemcpyAsync(ctx->h_tindptr.data(), ctx->t_indptr.p, (size_t)(m + 1) * 4,
                                 cudaMemcpyDeviceToHost, s)) != cudaSuccess)
            return failed(e, "term row pointers");
        if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return failed(e, "term-major build");
    }
    release_sort_scratch(ctx, nnz);
    ctx->term_items.ready = false;
    ctx->t_ready = true;
    ctx->t_weighted_ready = false;
    return PLSA_OK;
}

static int ensure_weighted_vals(plsa_ctx *ctx)
{
    if (ctx->t_weighted_ready) return PLSA_OK;
    const int64_t nnz = ctx->cur().nnz;
    CK(ctx->t_entw.ensure(ent_bytes(nnz)));
    CK(cudaMemsetAsync(ctx->t_entw.as<int2>() + nnz, 0, ENT_PAD * sizeof(int2), ctx->stream));
    if (nnz > 0) {
        weight_vals_kernel<<<(unsigned)cdiv(nnz, 256), 256, 0, ctx->stream>>>(
            ctx->t_ent.as<int2>(), ctx->sw.as<float>(), ctx->t_entw.as<int2>(), nnz);
        ctx->launches++;
        CK(cudaGetLastError());
    }
    ctx->t_weighted_ready = true;
    return PLSA_OK;
}

/* ---- tiled doc pass: head / tail split of the doc-major corpus (plsa_tile.cuh) --------------- */
static bool tiled_possible(const plsa_ctx *ctx, int kp)
{
    const Corpus &c = ctx->cur();
    return ctx->tiled_opt != 0 && ctx->shard == nullptr && kp >= 4 && kp <= 24 && c.n > 0 && c.m > 0 &&
           c.nnz > 0 && c.nnz < ((int64_t)1 << 28); /* padded head slots (< 8 nnz) stay int32 */
}

/* Automatic choice (option "tiled" = -1, the default).  Measured on B200
 * (profiles/r2_kernel_experiments.md): at C2 (10 M entries, both factors deep inside the L2) the
 * tiled passes are no faster than the group-per-row passes, whose gathers mostly hit L1/L2; at
 * C5 (200 M entries, P(z|d) = 128 MB > L2) they take 3.22 instead of 4.42 ms per iteration. */
static bool tiled_wanted(const plsa_ctx *ctx, int kp)
{
    if (!tiled_possible(ctx, kp)) return false;
    if (ctx->tiled_opt > 0) return true;
    return ctx->cur().nnz >= 64'000'000;
}

typedef void (*tile_fn)(const TileArgs);
template <int KC> static tile_fn pick_tile_ll(bool ll)
{
    return ll ? tile_pass_kernel<KC, true> : tile_pass_kernel<KC, false>;
}
static tile_fn pick_tile_kernel(int kp, bool ll)
{
    switch (kp / 4) {
    case 1: return pick_tile_ll<1>(ll);
    case 2: return pick_tile_ll<2>(ll);
    case 3: return pick_tile_ll<3>(ll);
    case 4: return pick_tile_ll<4>(ll);
    case 5: return pick_tile_ll<5>(ll);
    default: return pick_tile_ll<6>(ll);
    }
}

static int ensure_tiles(plsa_ctx *ctx, int kp)
{
    TileSet &t = ctx->tiles;
    const Corpus &c = ctx->cur();
    const int64_t n = c.n, m = c.m, nnz = c.nnz;
    const int pc = tile_pitch_chunks(kp / 4);
    const int32_t cap = (int32_t)std::max<int64_t>(TILE_MIN_ROWS, ctx->tile_bytes / (pc * 16));
    const int32_t tile_rows = (int32_t)std::min<int64_t>(m, cap);
    const int64_t chunk = choose_chunk(ctx, kp);
    const int align = ctx->vec_entries ? pass_entry_block(kp) : 1;
    if (t.ready && t.kp == kp && t.tile_rows == tile_rows && t.n == n && t.tail_items.ready &&
        t.tail_items.chunk == chunk && t.tail_items.align == align)
        return PLSA_OK;
    cudaStream_t s = ctx->stream;
    const int T = 256;
    t.ready = false;
    t.kp = kp;
    t.tile_rows = tile_rows;
    t.pitch_f = pc * 4;
    t.n = n;
    /* the most frequent columns: count, sort by count (descending, stable: ties by index) */
    CK(t.col_count.ensure((size_t)m * 4));
    CK(t.col_ids.ensure((size_t)m * 4));
    CK(t.col_count_sorted.ensure((size_t)m * 4));
    CK(t.col_sorted.ensure((size_t)m * 4));
    CK(t.slot_of.ensure((size_t)m * 4));
    CK(cudaMemsetAsync(t.col_count.p, 0, (size_t)m * 4, s));
    CK(cudaMemsetAsync(t.slot_of.p, 0xff, (size_t)m * 4, s));
    col_count_kernel<<<(unsigned)cdiv(nnz, T), T, 0, s>>>(c.ent.as<int2>(), nnz, t.col_count.as<int32_t>());
    iota_kernel<<<(unsigned)cdiv(m, T), T, 0, s>>>(t.col_ids.as<int32_t>(), m);
    size_t tmp_a = 0, tmp_b = 0, tmp_c = 0;
    CK(cub::DeviceRadixSort::SortPairsDescending(nullptr, tmp_a, t.col_count.as<int32_t>(),
                                                 t.col_count_sorted.as<int32_t>(), t.col_ids.as<int32_t>(),
                                                 t.col_sorted.as<int32_t>(), (int)m, 0, 32, s));
    CK(t.head_len.ensure((size_t)(n + 1) * 4));
    CK(t.head_mem.ensure((size_t)(n + 1) * 4));
    CK(t.tail_len.ensure((size_t)(n + 1) * 4));
    CK(t.head_indptr.ensure((size_t)(n + 1) * 4));
    CK(t.tail_indptr.ensure((size_t)(n + 1) * 4));
    CK(t.order.ensure((size_t)n * 4));
    CK(t.order_keys.ensure((size_t)n * 4));
    CK(t.row_ids.ensure((size_t)n * 4));
    CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp_b, t.head_len.as<int32_t>(), t.head_indptr.as<int32_t>(),
                                     (int)(n + 1), s));
    CK(cub::DeviceRadixSort::SortPairsDescending(nullptr, tmp_c, t.head_len.as<int32_t>(),
                                                 t.order_keys.as<int32_t>(), t.row_ids.as<int32_t>(),
                                                 t.order.as<int32_t>(), (int)n, 0, 32, s));
    CK(t.cub_tmp.ensure(std::max(tmp_a, std::max(tmp_b, tmp_c))));
    size_t tmp = t.cub_tmp.cap;
    CK(cub::DeviceRadixSort::SortPairsDescending(t.cub_tmp.p, tmp, t.col_count.as<int32_t>(),
                                                 t.col_count_sorted.as<int32_t>(), t.col_ids.as<int32_t>(),
                                                 t.col_sorted.as<int32_t>(), (int)m, 0, 32, s));
    tile_slot_kernel<<<(unsigned)cdiv(tile_rows, T), T, 0, s>>>(t.col_sorted.as<int32_t>(), tile_rows,
                                                              t.slot_of.as<int32_t>());
    /* per row: padded head length and tail length, then the two row-pointer arrays */
    CK(cudaMemsetAsync(t.head_len.as<int32_t>() + n, 0, 4, s));
    CK(cudaMemsetAsync(t.head_mem.as<int32_t>() + n, 0, 4, s));
    CK(cudaMemsetAsync(t.tail_len.as<int32_t>() + n, 0, 4, s));
    const SlotMap map{t.slot_of.as<int32_t>(), 1, 0};
    tile_count_kernel<<<(unsigned)cdiv(n * 32, T), T, 0, s>>>(c.indptr.as<int32_t>(), c.indptr.as<int32_t>() + 1, n,
                                                            c.ent.as<int2>(), map, t.head_len.as<int32_t>(),
                                                            t.head_mem.as<int32_t>(), t.tail_len.as<int32_t>());
    tmp = t.cub_tmp.cap;
    CK(cub::DeviceScan::ExclusiveSum(t.cub_tmp.p, tmp, t.head_mem.as<int32_t>(), t.head_indptr.as<int32_t>(),
                                     (int)(n + 1), s));
    tmp = t.cub_tmp.cap;
    CK(cub::DeviceScan::ExclusiveSum(t.cub_tmp.p, tmp, t.tail_len.as<int32_t>(), t.tail_indptr.as<int32_t>(),
                                     (int)(n + 1), s));
    /* rows by padded head length, longest first: the four items of a warp are equally long and
     * every warp's share of the work is the same to within its last item */
    iota_kernel<<<(unsigned)cdiv(n, T), T, 0, s>>>(t.row_ids.as<int32_t>(), n);
    tmp = t.cub_tmp.cap;
    CK(cub::DeviceRadixSort::SortPairsDescending(t.cub_tmp.p, tmp, t.head_len.as<int32_t>(),
                                                 t.order_keys.as<int32_t>(), t.row_ids.as<int32_t>(),
                                                 t.order.as<int32_t>(), (int)n, 0, 32, s));
    CK(t.headers.ensure((size_t)std::max<int64_t>(n, 1) * sizeof(int4)));
    tile_headers_kernel<<<(unsigned)cdiv(n, T), T, 0, s>>>(t.order.as<int32_t>(), t.head_indptr.as<int32_t>(),
                                                         t.head_len.as<int32_t>(), nullptr, n, t.headers.as<int4>());
    ctx->launches += 7;
    CK(cudaGetLastError());
    int32_t head_total = 0, tail_total = 0;
    if (!ctx->device_plan) { /* the host planner walks the tail's row pointers */
        try {
            t.h_tail_indptr.resize((size_t)n + 1);
        } catch (const std::bad_alloc &) {
            return ctx->fail(PLSA_ENOMEM, "tiles: out of host memory");
        }
        CK(cudaMemcpyAsync(t.h_tail_indptr.data(), t.tail_indptr.p, (size_t)(n + 1) * 4,
                           cudaMemcpyDeviceToHost, s));
    }
    CK(cudaMemcpyAsync(&tail_total, t.tail_indptr.as<int32_t>() + n, 4, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(&head_total, t.head_indptr.as<int32_t>() + n, 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    t.head_slots = head_total;
    t.tail_nnz = tail_total;
    if (t.head_slots < 0 || t.tail_nnz < 0 || t.tail_nnz > nnz)
        return ctx->fail(PLSA_ECUDA, "tiles: inconsistent head / tail split");
    CK(t.head_ent.ensure((size_t)(t.head_slots + 32) * sizeof(int2))); /* + a readable block */
    CK(cudaMemsetAsync(t.head_ent.as<int2>() + t.head_slots, 0, 32 * sizeof(int2), s));
    CK(t.tail_ent.ensure(ent_bytes(t.tail_nnz)));
    CK(cudaMemsetAsync(t.tail_ent.as<int2>() + t.tail_nnz, 0, ENT_PAD * sizeof(int2), s));
    tile_place_kernel<<<(unsigned)cdiv(n * 32, T), T, 0, s>>>(
        c.indptr.as<int32_t>(), c.indptr.as<int32_t>() + 1, n, c.ent.as<int2>(), map, t.head_indptr.as<int32_t>(),
        t.tail_indptr.as<int32_t>(), t.head_ent.as<int2>(), t.tail_ent.as<int2>(), nullptr);
    ctx->launches++;
    CK(cudaGetLastError());
    /* compact images of the tile rows (ping-pong like B), zero beyond the last row */
    const size_t img_bytes = (size_t)std::max<int32_t>(tile_rows, TILE_MIN_ROWS) * t.pitch_f * 4;
    for (int i = 0; i < 2; ++i) {
        CK(t.img[i].ensure(img_bytes));
        CK(cudaMemsetAsync(t.img[i].p, 0, img_bytes, s));
    }
    ctx->b_norm[0] = ctx->b_norm[1] = false; /* the images are empty */
    CK(t.head_sum.ensure((size_t)n * kp * 4));
    CK(t.scale_raw.ensure((size_t)kp * 4 * 2));
    CK(t.ll_part.ensure((size_t)ctx->n_sms * 8));
    CK(t.ll_head.ensure(8));
    CK(t.ll_ticket.ensure(4));
    CK(cudaMemsetAsync(t.ll_ticket.p, 0, 4, s));
    int rc = ctx->device_plan ? build_items_device(ctx, t.tail_indptr.as<int32_t>(), n, t.tail_items, chunk, align)
                              : build_items(ctx, t.h_tail_indptr, n, t.tail_items, chunk, align);
    if (rc) return rc;
    /* opt in to the tile's dynamic shared memory (the doc side and the term side share the
     * kernels and ask for different sizes: allow the largest tile the option permits) */
    for (int ll = 0; ll < 2; ++ll)
        CK(cudaFuncSetAttribute(pick_tile_kernel(kp, ll != 0), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                TILE_MAX_SMEM));
    t.ready = true;
    return PLSA_OK;
}

/* B[which] times its pending column scale, in place, plus its tile image (tiled mode keeps
 * P(w|z) normalised in memory: the flush-to-zero threshold needs gathered values <= 1) */
static int normalise_b(plsa_ctx *ctx, int which, const float *scale, cudaStream_t stream)
{
    TileSet &t = ctx->tiles;
    const int64_t m = ctx->cur().m;
    const int64_t threads = m * (ctx->kp / 4);
    normalise_rows_kernel<<<(unsigned)cdiv(threads, 256), 256, 0, stream>>>(
        ctx->B[which].as<float>(), m, ctx->strideB, ctx->kp, scale, t.slot_of.as<int32_t>(),
        t.img[which].as<float>(), t.pitch_f);
    ctx->launches++;
    CK(cudaGetLastError());
    return PLSA_OK;
}

/* ---- tiled term pass: (term, document block) items of the frequent terms ---------------------- */
static int ensure_term_tiles(plsa_ctx *ctx, int kp, bool weighted)
{
    TermTiles &t = ctx->tterm;
    const Corpus &c = ctx->cur();
    const int64_t n = c.n, m = c.m;
    const int pc = tile_pitch_chunks(kp / 4);
    int32_t block_rows = (int32_t)std::max<int64_t>(TILE_MIN_ROWS, ctx->tile_bytes / (pc * 16) / 8 * 8);
    block_rows = (int32_t)std::min<int64_t>(block_rows, (n + 7) / 8 * 8);
    const int32_t n_blocks = (int32_t)cdiv(n, block_rows);
    const int64_t chunk = choose_chunk(ctx, kp);
    const int align = ctx->vec_entries ? pass_entry_block(kp) : 1;
    if (t.ready && t.kp == kp && t.block_rows == block_rows && t.n == n && t.m == m && t.weighted == weighted &&
        t.grid == ctx->n_sms && t.tail_items.ready && t.tail_items.chunk == chunk && t.tail_items.align == align)
        return PLSA_OK;
    cudaStream_t s = ctx->stream;
    const int T = 256;
    t.ready = false;
    t.kp = kp; t.block_rows = block_rows; t.n_blocks = n_blocks; t.n = n; t.m = m; t.weighted = weighted;
    t.pitch_f = pc * 4;
    t.grid = ctx->n_sms;
    const int2 *t_ent = ctx->t_ent.as<int2>(); /* unweighted: the weights are applied when placing */
    /* which terms are tiled */
    CK(t.flag.ensure((size_t)(m + 1) * 4));
    CK(t.tiled_at.ensure((size_t)(m + 1) * 4));
    const int32_t min_count = (int32_t)std::min<int64_t>(((int64_t)1 << 30), ctx->term_tile_min * n_blocks);
    term_tiled_flag_kernel<<<(unsigned)cdiv(m + 1, T), T, 0, s>>>(ctx->t_indptr.as<int32_t>(), m, min_count,
                                                                t.flag.as<int32_t>());
    size_t tb = 0;
    CK(cub::DeviceScan::ExclusiveSum(nullptr, tb, t.flag.as<int32_t>(), t.tiled_at.as<int32_t>(), (int)(m + 1), s));
    CK(t.cub_tmp.ensure(std::max<size_t>(tb, 1 << 20)));
    tb = t.cub_tmp.cap;
    CK(cub::DeviceScan::ExclusiveSum(t.cub_tmp.p, tb, t.flag.as<int32_t>(), t.tiled_at.as<int32_t>(), (int)(m + 1), s));
    int32_t n_tiled = 0;
    CK(cudaMemcpyAsync(&n_tiled, t.tiled_at.as<int32_t>() + m, 4, cudaMemcpyDeviceToHost, s));
    if (!ctx->device_plan) {
        try {
            t.h_flag.resize((size_t)m + 1);
        } catch (const std::bad_alloc &) {
            return ctx->fail(PLSA_ENOMEM, "term tiles: out of host memory");
        }
        CK(cudaMemcpyAsync(t.h_flag.data(), t.flag.p, (size_t)(m + 1) * 4, cudaMemcpyDeviceToHost, s));
    }
    CK(cudaStreamSynchronize(s));
    if ((int64_t)n_tiled * n_blocks >= ((int64_t)1 << 30)) n_tiled = 0; /* item numbers stay int32 */
    t.n_tiled = n_tiled;
    t.n_items = (int64_t)n_tiled * n_blocks;
    const int64_t V = t.n_items;
    if (V == 0) { /* nothing worth tiling: the whole term pass stays with the group-per-row kernel */
        t.ready = true;
        t.tail_items.ready = false;
        return PLSA_OK;
    }
    /* the items: entry ranges, owned term */
    CK(t.tiled_terms.ensure((size_t)n_tiled * 4));
    CK(t.beg.ensure((size_t)V * 4));
    CK(t.end.ensure((size_t)V * 4));
    CK(t.own_row.ensure((size_t)V * 4));
    CK(t.head_len.ensure((size_t)(V + 1) * 4));
    CK(t.head_mem.ensure((size_t)(V + 1) * 4));
    CK(t.head_indptr.ensure((size_t)(V + 1) * 4));
    term_items_kernel<<<(unsigned)cdiv(m * n_blocks, T), T, 0, s>>>(
        ctx->t_indptr.as<int32_t>(), m, t.flag.as<int32_t>(), t.tiled_at.as<int32_t>(), t_ent, n_blocks, block_rows,
        t.beg.as<int32_t>(), t.end.as<int32_t>(), t.own_row.as<int32_t>(), t.tiled_terms.as<int32_t>());
    const SlotMap map{nullptr, n_blocks, block_rows};
    CK(cudaMemsetAsync(t.head_len.as<int32_t>() + V, 0, 4, s));
    CK(cudaMemsetAsync(t.head_mem.as<int32_t>() + V, 0, 4, s));
    tile_count_kernel<<<(unsigned)cdiv(V * 32, T), T, 0, s>>>(t.beg.as<int32_t>(), t.end.as<int32_t>(), V, t_ent, map,
                                                            t.head_len.as<int32_t>(), t.head_mem.as<int32_t>(),
                                                            nullptr);
    tb = 0;
    CK(cub::DeviceScan::ExclusiveSum(nullptr, tb, t.head_mem.as<int32_t>(), t.head_indptr.as<int32_t>(), (int)(V + 1), s));
    CK(t.cub_tmp.ensure(tb));
    tb = t.cub_tmp.cap;
    CK(cub::DeviceScan::ExclusiveSum(t.cub_tmp.p, tb, t.head_mem.as<int32_t>(), t.head_indptr.as<int32_t>(), (int)(V + 1), s));
    int32_t head_total = 0;
    CK(cudaMemcpyAsync(&head_total, t.head_indptr.as<int32_t>() + V, 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (head_total < 0) return ctx->fail(PLSA_ECUDA, "term tiles: inconsistent item lengths");
    t.head_slots = head_total;
    CK(t.head_ent.ensure((size_t)(t.head_slots + 32) * sizeof(int2))); /* + a readable block */
    CK(cudaMemsetAsync(t.head_ent.as<int2>() + t.head_slots, 0, 32 * sizeof(int2), s));
    tile_place_kernel<<<(unsigned)cdiv(V * 32, T), T, 0, s>>>(
        t.beg.as<int32_t>(), t.end.as<int32_t>(), V, t_ent, map, t.head_indptr.as<int32_t>(), nullptr,
        t.head_ent.as<int2>(), nullptr, weighted ? ctx->sw.as<float>() : nullptr);
    /* launch order: block-major, inside a block longest first; one range of equal work per CTA */
    CK(t.keys.ensure((size_t)V * 4));
    CK(t.keys_sorted.ensure((size_t)V * 4));
    CK(t.ids.ensure((size_t)V * 4));
    CK(t.order.ensure((size_t)V * 4));
    CK(t.work.ensure((size_t)(V + 1) * 4));
    CK(t.work_prefix.ensure((size_t)(V + 1) * 4));
    CK(t.cta_begin.ensure((size_t)(t.grid + 1) * 4));
    CK(t.block_begin.ensure((size_t)(n_blocks + 1) * 4));
    term_item_keys_kernel<<<(unsigned)cdiv(V, T), T, 0, s>>>(t.head_len.as<int32_t>(), V, n_blocks, block_rows,
                                                           t.keys.as<int32_t>());
    iota_kernel<<<(unsigned)cdiv(V, T), T, 0, s>>>(t.ids.as<int32_t>(), V);
    int end_bit = 1;
    while (((int64_t)1 << end_bit) < (int64_t)(n_blocks + 1) * (block_rows / 8 + 1)) ++end_bit;
    tb = 0;
    CK(cub::DeviceRadixSort::SortPairs(nullptr, tb, t.keys.as<int32_t>(), t.keys_sorted.as<int32_t>(),
                                       t.ids.as<int32_t>(), t.order.as<int32_t>(), (int)V, 0, end_bit, s));
    CK(t.cub_tmp.ensure(tb));
    tb = t.cub_tmp.cap;
    CK(cub::DeviceRadixSort::SortPairs(t.cub_tmp.p, tb, t.keys.as<int32_t>(), t.keys_sorted.as<int32_t>(),
                                       t.ids.as<int32_t>(), t.order.as<int32_t>(), (int)V, 0, end_bit, s));
    CK(t.headers.ensure((size_t)V * sizeof(int4)));
    tile_headers_kernel<<<(unsigned)cdiv(V, T), T, 0, s>>>(t.order.as<int32_t>(), t.head_indptr.as<int32_t>(),
                                                         t.head_len.as<int32_t>(), t.own_row.as<int32_t>(), V,
                                                         t.headers.as<int4>());
    term_work_kernel<<<(unsigned)cdiv(V + 1, T), T, 0, s>>>(t.order.as<int32_t>(), t.head_len.as<int32_t>(), V,
                                                          t.work.as<int32_t>());
    tb = t.cub_tmp.cap;
    CK(cub::DeviceScan::ExclusiveSum(t.cub_tmp.p, tb, t.work.as<int32_t>(), t.work_prefix.as<int32_t>(), (int)(V + 1), s));
    term_ranges_kernel<<<(unsigned)cdiv(std::max(t.grid, n_blocks) + 1, T), T, 0, s>>>(
        t.work_prefix.as<int32_t>(), t.keys_sorted.as<int32_t>(), V, t.grid, n_blocks, block_rows,
        t.cta_begin.as<int32_t>(), t.block_begin.as<int32_t>());
    ctx->launches += 9;
    CK(cudaGetLastError());
    /* partial sums: one slot per item; a term's slots are consecutive (fixup_kernel adds them) */
    CK(t.partial.ensure((size_t)V * kp * 4));
    std::vector<int32_t> h_slot_begin;
    try {
        h_slot_begin.resize((size_t)n_tiled + 1);
    } catch (const std::bad_alloc &) {
        return ctx->fail(PLSA_ENOMEM, "term tiles: out of host memory");
    }
    for (int32_t i = 0; i <= n_tiled; ++i) h_slot_begin[(size_t)i] = i * n_blocks;
    CK(t.slot_begin.ensure(h_slot_begin.size() * 4));
    CK(cudaMemcpyAsync(t.slot_begin.p, h_slot_begin.data(), h_slot_begin.size() * 4, cudaMemcpyHostToDevice, s));
    /* compact images of P(z|d) (ping-pong like A), zero behind the last row */
    const size_t img_bytes = ((size_t)n_blocks * block_rows + TILE_MIN_ROWS) * t.pitch_f * 4;
    for (int i = 0; i < 2; ++i) {
        CK(t.acomp[i].ensure(img_bytes));
        CK(cudaMemsetAsync(t.acomp[i].p, 0, img_bytes, s));
    }
    ctx->a_comp[0] = ctx->a_comp[1] = false;
    CK(cudaStreamSynchronize(s)); /* h_slot_begin dies here */
    /* the other terms: group-per-row items that skip the tiled rows */
    int rc = ctx->device_plan
                 ? build_items_device(ctx, ctx->t_indptr.as<int32_t>(), m, t.tail_items, chunk, align, nullptr,
                                      t.flag.as<int32_t>())
                 : build_items(ctx, ctx->h_tindptr, m, t.tail_items, chunk, align, nullptr, t.h_flag.data());
    if (rc) return rc;
    CK(cudaFuncSetAttribute(pick_tile_kernel(kp, false), cudaFuncAttributeMaxDynamicSharedMemorySize,
                            TILE_MAX_SMEM));
    t.ready = true;
    return PLSA_OK;
}

/* A[which] (padded rows) -> its compact image, the TMA source of the term pass's tiles */
static int compact_a(plsa_ctx *ctx, int which, cudaStream_t stream)
{
    TermTiles &t = ctx->tterm;
    const int64_t n = ctx->cur().n;
    const int64_t threads = n * (ctx->kp / 4);
    compact_rows_kernel<<<(unsigned)cdiv(threads, 256), 256, 0, stream>>>(
        ctx->A[which].as<float>(), n, ctx->strideA, ctx->kp, t.acomp[which].as<float>(), t.pitch_f);
    ctx->launches++;
    CK(cudaGetLastError());
    return PLSA_OK;
}

static void corpus_changed(plsa_ctx *ctx)
{
    ctx->presorted = false; /* set again by an upload that started the sort */
    ctx->tiles.ready = false;
    ctx->tiles.tail_items.ready = false;
    ctx->tterm.ready = false;
    ctx->tterm.tail_items.ready = false;
    ctx->t_ready = false;
    ctx->t_weighted_ready = false;
    ctx->doc_items.ready = false;
    ctx->term_items.ready = false;
    ctx->have_factors = false;
    ctx->have_sw = false;
}

/* ============================================================================================
 * C ABI
 * ============================================================================================ */
API int plsa_version(void) { return 200; }

API int plsa_device_count(int *count)
{
    if (!count) return PLSA_EINVAL;
    cudaError_t e = cudaGetDeviceCount(count);
    if (e != cudaSuccess) {
        *count = 0;
        g_err = std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e);
        return PLSA_ECUDA;
    }
    return PLSA_OK;
}

API const char *plsa_last_error(const plsa_ctx *ctx) { return ctx ? ctx->err.c_str() : g_err.c_str(); }

API int plsa_ctx_create(int device, plsa_ctx **out)
{
    if (!out) return PLSA_EINVAL;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        g_err = std::string("no CUDA device: ") + cudaGetErrorString(e);
        return PLSA_ECUDA;
    }
    if (device < 0 || device >= count) {
        g_err = "device index out of range";
        return PLSA_EINVAL;
    }
    plsa_ctx *ctx = new (std::nothrow) plsa_ctx();
    if (!ctx) return PLSA_ENOMEM;
    ctx->device = device;
    if ((e = cudaSetDevice(device)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaEventCreate(&ctx->ev0)) != cudaSuccess ||
        (e = cudaEventCreate(&ctx->ev1)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&ctx->ev_ll, cudaEventDisableTiming)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&ctx->ev_a, cudaEventDisableTiming)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&ctx->ev_b, cudaEventDisableTiming)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&ctx->ev_go, cudaEventDisableTiming)) != cudaSuccess) {
        g_err = std::string("context setup: ") + cudaGetErrorString(e);
        delete ctx;
        return PLSA_ECUDA;
    }
    int max_lin = 0;
    if (cudaDeviceGetAttribute(&max_lin, cudaDevAttrMaxTexture1DLinearWidth, device) == cudaSuccess &&
        max_lin > 0)
        ctx->tex_max_texels = (size_t)max_lin;
    cudaDeviceGetAttribute(&ctx->n_sms, cudaDevAttrMultiProcessorCount, device);
    if (ctx->n_sms <= 0) ctx->n_sms = 148;
    /* every context of the process, the one-shot entry points' included (A/B runs, tests) */
    if (const char *e = getenv("ENSTOP_B200_TILED")) ctx->tiled_opt = std::max(-1, std::min(1, atoi(e)));
    if (const char *e = getenv("ENSTOP_B200_TERM_TILED")) ctx->term_tiled_opt = atoi(e) != 0;
    if (const char *e = getenv("ENSTOP_B200_TERM_TILE_MIN"))
        ctx->term_tile_min = std::max(1, std::min(100000, atoi(e)));
    if (const char *e = getenv("ENSTOP_B200_DEVICE_PLAN")) ctx->device_plan = atoi(e) != 0;
    if (const char *e = getenv("ENSTOP_B200_TWO_SHOT")) ctx->p2p.two_shot = atoi(e) < 0 ? -1 : (atoi(e) != 0);
    if (const char *e = getenv("ENSTOP_B200_TILE_KB"))
        ctx->tile_bytes = (int64_t)std::max(1, std::min(220, atoi(e))) * 1024;
    *out = ctx;
    return PLSA_OK;
}

API int plsa_ctx_destroy(plsa_ctx *ctx)
{
    if (!ctx) return PLSA_OK;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->stream2) cudaStreamSynchronize(ctx->stream2); /* a term sort nobody waited for */
    prof_collect(ctx);
    for (cudaEvent_t e : ctx->ev_pool) cudaEventDestroy(e);
    if (ctx->ev_presort) cudaEventDestroy(ctx->ev_presort);
    if (ctx->ev_cols) cudaEventDestroy(ctx->ev_cols);
    if (ctx->ev_d2h) cudaEventDestroy(ctx->ev_d2h);
    for (Corpus *c : {&ctx->base, &ctx->boot}) {
        c->indptr.release(); c->ent.release();
    }
    p2p_release(ctx);
    ctx->tiles.release();
    ctx->tterm.release();
    for (DevBuf &b : ctx->scratch) b.release();
    ctx->doc_items.release();
    ctx->term_items.release();
    for (DevBuf *b : {&ctx->t_ent, &ctx->t_entw, &ctx->t_indptr, &ctx->up_cols, &ctx->up_vals, &ctx->flag,
                      &ctx->A[0], &ctx->A[1], &ctx->B[0], &ctx->B[1],
                      &ctx->scale, &ctx->ones, &ctx->colnorm, &ctx->colpart, &ctx->partialA,
                      &ctx->partialB, &ctx->sw, &ctx->ll_part, &ctx->ll_out, &ctx->stage, &ctx->tickets,
                      &ctx->topics_dev, &ctx->gathered, &ctx->ll2, &ctx->colpart2})
        b->release();
    for (int i = 0; i < 2; ++i) {
        if (ctx->texA[i]) cudaDestroyTextureObject(ctx->texA[i]);
        if (ctx->texB[i]) cudaDestroyTextureObject(ctx->texB[i]);
    }
    if (ctx->mail) cudaFreeHost(ctx->mail);
    if (ctx->pin_factors) cudaFreeHost(ctx->pin_factors);
    for (int t = 0; t < plsa_ctx::H2D_THREADS; ++t) {
        if (ctx->pin[t]) cudaFreeHost(ctx->pin[t]);
        if (ctx->pin_stream[t]) cudaStreamDestroy(ctx->pin_stream[t]);
        for (int b = 0; b < 2; ++b)
            if (ctx->pin_ev[t][b]) cudaEventDestroy(ctx->pin_ev[t][b]);
    }
    if (ctx->ev_ll) cudaEventDestroy(ctx->ev_ll);
    if (ctx->ev_a) cudaEventDestroy(ctx->ev_a);
    if (ctx->ev_b) cudaEventDestroy(ctx->ev_b);
    if (ctx->ev_go) cudaEventDestroy(ctx->ev_go);
    if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return PLSA_OK;
}

/* ---- corpus ----------------------------------------------------------------------------------- */
static size_t dtype_size(int dtype)
{
    switch (dtype) {
    case PLSA_F32: case PLSA_I32: return 4;
    case PLSA_F64: case PLSA_I64: return 8;
    default: return 0;
    }
}

static int upload_csr_impl(plsa_ctx *ctx, const int32_t *indptr, const int32_t *indices,
                           const void *data, int dtype, int64_t n, int64_t m, int64_t nnz)
{
    const size_t vsz = dtype_size(dtype);
    if (!vsz) return ctx->fail(PLSA_EINVAL, "upload: unknown value dtype");
    if (n < 0 || m < 0 || nnz < 0 || (nnz > 0 && (!indices || !data)) || !indptr)
        return ctx->fail(PLSA_EINVAL, "upload: null pointer or negative size");
    if (nnz >= ((int64_t)1 << 31) || n >= ((int64_t)1 << 31) - 1 || m >= ((int64_t)1 << 31) - 1)
        return ctx->fail(PLSA_EINVAL, "upload: sizes must fit int32 indices");
    if (indptr[0] != 0 || indptr[n] != nnz)
        return ctx->fail(PLSA_EINVAL, "upload: indptr[0] != 0 or indptr[n] != nnz");
    for (int64_t r = 0; r < n; ++r)
        if (indptr[r + 1] < indptr[r]) return ctx->fail(PLSA_EINVAL, "upload: indptr decreases");
    Corpus &c = ctx->base;
    /* a failed upload leaves the context without a corpus, not with half of one */
    ctx->use_boot = false;
    corpus_changed(ctx);
    c.h_indptr.clear();
    c.n = n; c.m = m; c.nnz = nnz;
    std::vector<int32_t> h_indptr;
    try { /* no C++ exception crosses the C ABI */
        h_indptr.assign(indptr, indptr + n + 1);
    } catch (const std::bad_alloc &) {
        return ctx->fail(PLSA_ENOMEM, "upload: out of host memory");
    }
    CK(c.indptr.ensure((size_t)(n + 1) * 4));
    CK(c.ent.ensure(ent_bytes(nnz)));
    CK(ctx->flag.ensure(4));
    CK(cudaMemsetAsync(ctx->flag.p, 0, 4, ctx->stream));
    CK(cudaMemsetAsync(c.ent.as<int2>() + nnz, 0, ENT_PAD * sizeof(int2), ctx->stream));
    CK(cudaMemcpyAsync(c.indptr.p, indptr, (size_t)(n + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
    int bad = 0;
    bool presort_started = false;
    if (nnz > 0) {
        CK(ctx->up_cols.ensure((size_t)nnz * 4));
        CK(ctx->up_vals.ensure((size_t)nnz * vsz));
        if (ctx->ev_presort) /* an earlier sort that nobody waited for still reads up_cols */
            CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_presort, 0));
        CK(h2d_fast(ctx, ctx->up_cols.p, indices, (size_t)nnz * 4));
        if (ctx->presort_opt && ctx->stream2) {
            /* the indices are enough to sort the entries by term: do that on the second stream
             * while the values follow over PCIe (build_term_major picks the result up) */
            if (!ctx->ev_cols) CK(cudaEventCreateWithFlags(&ctx->ev_cols, cudaEventDisableTiming));
            if (!ctx->ev_presort) CK(cudaEventCreateWithFlags(&ctx->ev_presort, cudaEventDisableTiming));
            CK(cudaEventRecord(ctx->ev_cols, ctx->stream));
            CK(cudaStreamWaitEvent(ctx->stream2, ctx->ev_cols, 0));
            const int src = term_major_sort(ctx, c, ctx->up_cols.as<int32_t>(), ctx->stream2);
            if (src != PLSA_OK) return src;
            CK(cudaEventRecord(ctx->ev_presort, ctx->stream2));
            presort_started = true;
        }
        CK(h2d_fast(ctx, ctx->up_vals.p, data, (size_t)nnz * vsz));
        const unsigned grid = (unsigned)cdiv(nnz, 256);
        int2 *ent = c.ent.as<int2>();
        int *flag = ctx->flag.as<int>();
        const int32_t *cols = ctx->up_cols.as<int32_t>();
        switch (dtype) { /* the float32 cast of plsa.py:714 happens on the device */
        case PLSA_F32: interleave_kernel<float><<<grid, 256, 0, ctx->stream>>>(cols, ctx->up_vals.as<float>(), nnz, m, ent, flag); break;
        case PLSA_F64: interleave_kernel<double><<<grid, 256, 0, ctx->stream>>>(cols, ctx->up_vals.as<double>(), nnz, m, ent, flag); break;
        case PLSA_I32: interleave_kernel<int32_t><<<grid, 256, 0, ctx->stream>>>(cols, ctx->up_vals.as<int32_t>(), nnz, m, ent, flag); break;
        default: interleave_kernel<int64_t><<<grid, 256, 0, ctx->stream>>>(cols, ctx->up_vals.as<int64_t>(), nnz, m, ent, flag); break;
        }
        ctx->launches++;
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(&bad, ctx->flag.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream));
    if (bad) return ctx->fail(PLSA_EINVAL, "upload: column index out of range");
    c.h_indptr.swap(h_indptr);
    ctx->presorted = presort_started;
    return PLSA_OK;
}

API int plsa_upload_csr(plsa_ctx *ctx, const int32_t *indptr, const int32_t *indices,
                        const float *data, int64_t n_docs, int64_t n_terms, int64_t nnz)
{
    CHECK_CTX(ctx);
    return upload_csr_impl(ctx, indptr, indices, data, PLSA_F32, n_docs, n_terms, nnz);
}

API int plsa_upload_csr_typed(plsa_ctx *ctx, const int32_t *indptr, const int32_t *indices,
                              const void *data, int32_t dtype, int64_t n_docs, int64_t n_terms,
                              int64_t nnz)
{
    CHECK_CTX(ctx);
    return upload_csr_impl(ctx, indptr, indices, data, dtype, n_docs, n_terms, nnz);
}

API int plsa_upload_coo(plsa_ctx *ctx, const int32_t *rows, const int32_t *cols,
                        const float *vals, int64_t n_docs, int64_t n_terms, int64_t nnz)
{
    CHECK_CTX(ctx);
    if (n_docs < 0 || nnz < 0 || (nnz > 0 && !rows))
        return ctx->fail(PLSA_EINVAL, "upload_coo: null pointer or negative size");
    std::vector<int32_t> indptr;
    try {
        indptr.assign((size_t)n_docs + 1, 0);
    } catch (const std::bad_alloc &) {
        return ctx->fail(PLSA_ENOMEM, "upload_coo: out of host memory");
    }
    for (int64_t i = 0; i < nnz; ++i) {
        if ((uint32_t)rows[i] >= (uint32_t)n_docs)
            return ctx->fail(PLSA_EINVAL, "upload_coo: row index out of range");
        if (i > 0 && rows[i] < rows[i - 1])
            return ctx->fail(PLSA_EINVAL,
                             "upload_coo: triplets must be sorted by row (X.tocoo() of a CSR matrix)");
        indptr[(size_t)rows[i] + 1]++;
    }
    for (int64_t r = 0; r < n_docs; ++r) indptr[(size_t)r + 1] += indptr[(size_t)r];
    return upload_csr_impl(ctx, indptr.data(), cols, vals, PLSA_F32, n_docs, n_terms, nnz);
}

API int plsa_bootstrap(plsa_ctx *ctx, const int32_t *row_idx, int64_t n_rows)
{
    CHECK_CTX(ctx);
    if (ctx->base.h_indptr.empty()) return ctx->fail(PLSA_EINVAL, "bootstrap: no corpus uploaded");
    if (!row_idx) {
        ctx->use_boot = false;
        corpus_changed(ctx);
        return PLSA_OK;
    }
    if (n_rows < 0) return ctx->fail(PLSA_EINVAL, "bootstrap: negative row count");
    const Corpus &b = ctx->base;
    Corpus &c = ctx->boot;
    /* the row pointers are checked and built aside: a rejected call leaves the context on the
     * corpus it had; a call that fails later leaves it on the base corpus */
    std::vector<int32_t> h_indptr;
    try {
        h_indptr.assign((size_t)n_rows + 1, 0);
    } catch (const std::bad_alloc &) {
        return ctx->fail(PLSA_ENOMEM, "bootstrap: out of host memory");
    }
    int64_t run = 0;
    for (int64_t i = 0; i < n_rows; ++i) {
        if ((uint32_t)row_idx[i] >= (uint32_t)b.n)
            return ctx->fail(PLSA_EINVAL, "bootstrap: row index out of range");
        run += b.h_indptr[(size_t)row_idx[i] + 1] - b.h_indptr[(size_t)row_idx[i]];
        if (run >= ((int64_t)1 << 31))
            return ctx->fail(PLSA_EINVAL, "bootstrap: resampled corpus exceeds int32 entries");
        h_indptr[(size_t)i + 1] = (int32_t)run;
    }
    ctx->use_boot = false;
    corpus_changed(ctx);
    c.h_indptr.swap(h_indptr);
    c.n = n_rows; c.m = b.m; c.nnz = run;
    CK(c.indptr.ensure((size_t)(n_rows + 1) * 4));
    CK(c.ent.ensure(ent_bytes(run)));
    CK(cudaMemsetAsync(c.ent.as<int2>() + run, 0, ENT_PAD * sizeof(int2), ctx->stream));
    CK(ctx->stage.ensure((size_t)std::max<int64_t>(n_rows, 1) * 4));
    CK(cudaMemcpyAsync(c.indptr.p, c.h_indptr.data(), (size_t)(n_rows + 1) * 4,
                       cudaMemcpyHostToDevice, ctx->stream));
    if (n_rows > 0) {
        CK(cudaMemcpyAsync(ctx->stage.p, row_idx, (size_t)n_rows * 4, cudaMemcpyHostToDevice,
                           ctx->stream));
        gather_rows_kernel<<<(unsigned)cdiv(n_rows * 32, 256), 256, 0, ctx->stream>>>(
            ctx->stage.as<int32_t>(), n_rows, b.indptr.as<int32_t>(), b.ent.as<int2>(),
            c.indptr.as<int32_t>(), c.ent.as<int2>());
        ctx->launches++;
        CK(cudaGetLastError());
    }
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->use_boot = true;
    corpus_changed(ctx);
    return PLSA_OK;
}

API int plsa_corpus_shape(const plsa_ctx *ctx, int64_t *n_docs, int64_t *n_terms, int64_t *nnz)
{
    if (!ctx) return PLSA_EINVAL;
    const Corpus &c = ctx->cur();
    if (n_docs) *n_docs = c.n;
    if (n_terms) *n_terms = c.m;
    if (nnz) *nnz = c.nnz;
    return PLSA_OK;
}

/* ---- model state -------------------------------------------------------------------------------- */
/* A factor buffer as a linear float4 texture: the row pass gathers through the texture pipe,
 * which leaves the LSU/shared-memory pipe to the shuffles.  0 if the buffer is too large. */
static int make_texture(plsa_ctx *ctx, cudaTextureObject_t *tex, void *ptr, size_t bytes)
{
    /* repeated fits reuse their factor buffers: keep the texture object of an unchanged
     * (pointer, size) — creating four of them cost ~1 ms of every plsa_set_factors */
    auto &bound = ctx->tex_bound[tex];
    if (*tex && bound.first == ptr && bound.second == bytes) return PLSA_OK;
    bound = std::make_pair(ptr, bytes);
    if (*tex) {
        cudaDestroyTextureObject(*tex);
        *tex = 0;
    }
    if (bytes / 16 > ctx->tex_max_texels) return PLSA_OK; /* kernels fall back to LDG */
    cudaResourceDesc rd;
    memset(&rd, 0, sizeof(rd));
    rd.resType = cudaResourceTypeLinear;
    rd.res.linear.devPtr = ptr;
    rd.res.linear.desc = cudaCreateChannelDesc<float4>();
    rd.res.linear.sizeInBytes = bytes;
    cudaTextureDesc td;
    memset(&td, 0, sizeof(td));
    td.readMode = cudaReadModeElementType;
    CK(cudaCreateTextureObject(tex, &rd, &td, nullptr));
    return PLSA_OK;
}

static inline float *cur_scale(plsa_ctx *ctx)
{
    return reinterpret_cast<float *>(ctx->scale.p) + (size_t)ctx->curB * ctx->kp;
}

static int32_t row_stride(int32_t kp)
{
    if (kp * 4 <= 128) { /* rows never straddle a 128-byte line */
        int32_t s = 4;
        while (s < kp) s <<= 1;
        return s;
    }
    return (kp + 31) / 32 * 32;
}

static int fill(plsa_ctx *ctx, float *p, int64_t n, float v)
{
    if (n > 0) {
        fill_kernel<<<(unsigned)cdiv(n, 256), 256, 0, ctx->stream>>>(p, n, v);
        ctx->launches++;
        CK(cudaGetLastError());
    }
    return PLSA_OK;
}

API int plsa_set_factors(plsa_ctx *ctx, const float *p_z_given_d, const float *p_w_given_z,
                         int32_t k)
{
    CHECK_CTX(ctx);
    const Corpus &c = ctx->cur();
    if (c.h_indptr.empty()) return ctx->fail(PLSA_EINVAL, "set_factors: no corpus uploaded");
    if (k < 1 || k > PLSA_MAX_K) return ctx->fail(PLSA_EINVAL, "set_factors: k out of range");
    if (!p_z_given_d || !p_w_given_z) return ctx->fail(PLSA_EINVAL, "set_factors: null factor");
    const int32_t kp = (k + 3) / 4 * 4;
    ctx->k = k;
    ctx->kp = kp;
    ctx->strideA = ctx->strideB = row_stride(kp);
    const int64_t n = c.n, m = c.m;
    const size_t bytesA = (size_t)std::max<int64_t>(n, 1) * ctx->strideA * 4;
    const size_t bytesB = (size_t)std::max<int64_t>(m, 1) * ctx->strideB * 4;
    for (int i = 0; i < 2; ++i) {
        CK(ctx->A[i].ensure(bytesA));
        CK(ctx->B[i].ensure(bytesB));
        CK(cudaMemsetAsync(ctx->A[i].p, 0, bytesA, ctx->stream));
        CK(cudaMemsetAsync(ctx->B[i].p, 0, bytesB, ctx->stream));
        int trc;
        if ((trc = make_texture(ctx, &ctx->texA[i], ctx->A[i].p, bytesA))) return trc;
        if ((trc = make_texture(ctx, &ctx->texB[i], ctx->B[i].p, bytesB))) return trc;
    }
    CK(ctx->scale.ensure((size_t)kp * 4 * 2)); /* one per P(w|z) ping-pong buffer */
    CK(ctx->ones.ensure((size_t)kp * 4));
    CK(ctx->colnorm.ensure((size_t)kp * 8));
    CK(ctx->ll_out.ensure(8));
    CK(ctx->tickets.ensure(16));
    CK(cudaMemsetAsync(ctx->tickets.p, 0, 16, ctx->stream));
    int rc;
    if ((rc = fill(ctx, ctx->scale.as<float>(), 2 * kp, 1.f))) return rc;
    if ((rc = fill(ctx, ctx->ones.as<float>(), kp, 1.f))) return rc;

    const size_t stage_bytes = (size_t)std::max<int64_t>(std::max(n, m), 1) * k * 4;
    CK(ctx->stage.ensure(stage_bytes));
    if (n > 0) {
        CK(cudaMemcpyAsync(ctx->stage.p, p_z_given_d, (size_t)n * k * 4, cudaMemcpyHostToDevice,
                           ctx->stream));
        pack_rows_kernel<<<(unsigned)cdiv(n * kp, 256), 256, 0, ctx->stream>>>(
            ctx->stage.as<float>(), ctx->A[0].as<float>(), n, k, kp, ctx->strideA, 0);
        ctx->launches++;
    }
    if (m > 0) {
        CK(cudaMemcpyAsync(ctx->stage.p, p_w_given_z, (size_t)m * k * 4, cudaMemcpyHostToDevice,
                           ctx->stream));
        pack_rows_kernel<<<(unsigned)cdiv(m * kp, 256), 256, 0, ctx->stream>>>(
            ctx->stage.as<float>(), ctx->B[0].as<float>(), m, k, kp, ctx->strideB, 1);
        ctx->launches++;
    }
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->curA = ctx->curB = 0;
    ctx->b_norm[0] = ctx->b_norm[1] = false;
    ctx->a_comp[0] = ctx->a_comp[1] = false;
    ctx->have_factors = true;
    return PLSA_OK;
}

/* Page-locked host memory for the initial factors, owned by the context and reused by later
 * fits: the seeded initialisation writes straight into it (no first-touch page faults on 12 MB
 * of fresh pages) and plsa_set_factors then copies at full PCIe speed.  May be called while
 * another thread uploads the corpus through the same context: it touches nothing else. */
API int plsa_pinned_factors(plsa_ctx *ctx, int64_t n_docs, int64_t n_terms, int32_t k,
                            float **p_z_given_d, float **p_w_given_z)
{
    CHECK_CTX(ctx);
    if (n_docs < 0 || n_terms < 0 || k < 1 || k > PLSA_MAX_K || !p_z_given_d || !p_w_given_z)
        return ctx->fail(PLSA_EINVAL, "pinned_factors: bad arguments");
    const size_t a = ((size_t)n_docs * k * 4 + 255) / 256 * 256, b = (size_t)n_terms * k * 4;
    if (a + b > ctx->pin_factors_cap) {
        if (ctx->pin_factors) cudaFreeHost(ctx->pin_factors);
        ctx->pin_factors = nullptr;
        ctx->pin_factors_cap = 0;
        const size_t want = (a + b) + (a + b) / 16 + 256;
        CK(cudaHostAlloc((void **)&ctx->pin_factors, want, cudaHostAllocDefault));
        ctx->pin_factors_cap = want;
    }
    *p_z_given_d = reinterpret_cast<float *>(ctx->pin_factors);
    *p_w_given_z = reinterpret_cast<float *>(ctx->pin_factors + a);
    return PLSA_OK;
}

/* Page-locked host memory for the caller's own buffers (e.g. the CSR arrays of a corpus that
 * is fitted repeatedly): uploads from it skip the staging copy. */
API int plsa_host_alloc(int64_t bytes, void **ptr)
{
    if (!ptr || bytes < 0) return PLSA_EINVAL;
    *ptr = nullptr;
    const cudaError_t e = cudaHostAlloc(ptr, (size_t)std::max<int64_t>(bytes, 1), cudaHostAllocPortable);
    if (e != cudaSuccess) {
        g_err = std::string("host_alloc: ") + cudaGetErrorString(e);
        return e == cudaErrorMemoryAllocation ? PLSA_ENOMEM : PLSA_ECUDA;
    }
    return PLSA_OK;
}

API int plsa_host_free(void *ptr)
{
    if (ptr) cudaFreeHost(ptr);
    return PLSA_OK;
}

API int plsa_set_sample_weight(plsa_ctx *ctx, const float *sample_weight)
{
    CHECK_CTX(ctx);
    const Corpus &c = ctx->cur();
    if (c.h_indptr.empty()) return ctx->fail(PLSA_EINVAL, "set_sample_weight: no corpus uploaded");
    CK(ctx->sw.ensure((size_t)std::max<int64_t>(c.n, 1) * 4));
    if (sample_weight) {
        if (c.n > 0)
            CK(cudaMemcpyAsync(ctx->sw.p, sample_weight, (size_t)c.n * 4, cudaMemcpyHostToDevice,
                               ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    } else {
        int rc = fill(ctx, ctx->sw.as<float>(), c.n, 1.f);
        if (rc) return rc;
    }
    ctx->have_sw = true;
    ctx->t_weighted_ready = false;
    if (ctx->tterm.weighted) ctx->tterm.ready = false; /* its entries carry the old weights */
    return PLSA_OK;
}

/* pageable destination: the copy of a large result out of the page-locked bounce buffer is cut
 * into slices, one per thread — most of its cost is the first touch of the destination's fresh
 * pages, which parallelises */
static void host_copy_mt(void *dst, const void *src, size_t bytes)
{
    static const int threads = [] { /* the rank's share of the cores, at most H2D_THREADS */
        cpu_set_t set;
        int cores = 0;
        if (sched_getaffinity(0, sizeof(set), &set) == 0) cores = CPU_COUNT(&set);
        if (cores <= 0) cores = (int)std::thread::hardware_concurrency();
        int ranks = 1;
        if (const char *e = getenv("LOCAL_WORLD_SIZE")) ranks = std::max(1, atoi(e));
        return std::max(1, std::min(plsa_ctx::H2D_THREADS, cores / ranks));
    }();
    const int T = bytes >= ((size_t)2 << 20) ? threads : 1;
    if (T <= 1) {
        memcpy(dst, src, bytes);
        return;
    }
    const size_t slice = (bytes / T + 4095) / 4096 * 4096;
    std::thread th[plsa_ctx::H2D_THREADS];
    int started = 1;
    try { /* no C++ exception crosses the C ABI: what no thread took is copied here */
        for (; started < T; ++started) {
            const size_t lo = std::min(bytes, slice * (size_t)started), hi = std::min(bytes, lo + slice);
            th[started] = std::thread([=]() { memcpy((char *)dst + lo, (const char *)src + lo, hi - lo); });
        }
    } catch (...) {
    }
    memcpy(dst, src, std::min(bytes, slice));
    if (started < T) {
        const size_t lo = std::min(bytes, slice * (size_t)started);
        memcpy((char *)dst + lo, (const char *)src + lo, bytes - lo);
    }
    for (int t = 1; t < started; ++t) th[t].join();
}

static bool is_pinned_host(const void *p)
{
    cudaPointerAttributes attr;
    const bool yes = cudaPointerGetAttributes(&attr, p) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    cudaGetLastError(); /* an ordinary pointer is not an error */
    return yes;
}

API int plsa_get_factors(plsa_ctx *ctx, float *p_z_given_d, float *p_w_given_z)
{
    CHECK_CTX(ctx);
    if (!ctx->have_factors) return ctx->fail(PLSA_EINVAL, "get_factors: no factors set");
    const Corpus &c = ctx->cur();
    const int64_t n = c.n, m = c.m;
    const int k = ctx->k;
    CK(ctx->stage.ensure((size_t)std::max<int64_t>(std::max(n, m), 1) * k * 4));
    const size_t bytes_a = (p_z_given_d && n > 0) ? (size_t)n * k * 4 : 0;
    const size_t bytes_b = (p_w_given_z && m > 0) ? (size_t)m * k * 4 : 0;
    /* A device-to-host copy into pageable memory goes through the driver's own staging at a
     * fraction of the PCIe rate.  Large results land in the context's page-locked factor buffer
     * by DMA instead and are copied out by a few threads; P(z|d) is copied out while P(w|z) is
     * still on its way.  A page-locked destination (plsa_host_alloc) takes the DMA directly. */
    const bool bounce_a = bytes_a >= ((size_t)1 << 20) && !is_pinned_host(p_z_given_d);
    const bool bounce_b = bytes_b >= ((size_t)1 << 20) && !is_pinned_host(p_w_given_z);
    char *pin_a = nullptr, *pin_b = nullptr;
    if (bounce_a || bounce_b) {
        float *fa = nullptr, *fb = nullptr;
        int rc = plsa_pinned_factors(ctx, n, m, k, &fa, &fb);
        if (rc != PLSA_OK) return rc;
        pin_a = reinterpret_cast<char *>(fa);
        pin_b = reinterpret_cast<char *>(fb);
        if (!ctx->ev_d2h) CK(cudaEventCreateWithFlags(&ctx->ev_d2h, cudaEventDisableTiming));
    }
    if (bytes_a) {
        unpack_rows_kernel<<<(unsigned)cdiv(n * k, 256), 256, 0, ctx->stream>>>(
            ctx->A[ctx->curA].as<float>(), nullptr, ctx->stage.as<float>(), n, k, ctx->strideA, 0);
        ctx->launches++;
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(bounce_a ? (void *)pin_a : (void *)p_z_given_d, ctx->stage.p, bytes_a,
                           cudaMemcpyDeviceToHost, ctx->stream));
        if (bounce_a) CK(cudaEventRecord(ctx->ev_d2h, ctx->stream));
        else CK(cudaStreamSynchronize(ctx->stream));
    }
    if (bytes_b) {
        unpack_rows_kernel<<<(unsigned)cdiv(m * k, 256), 256, 0, ctx->stream>>>(
            ctx->B[ctx->curB].as<float>(), cur_scale(ctx), ctx->stage.as<float>(), m, k,
            ctx->strideB, 1);
        ctx->launches++;
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(bounce_b ? (void *)pin_b : (void *)p_w_given_z, ctx->stage.p, bytes_b,
                           cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (bounce_a) {
        CK(cudaEventSynchronize(ctx->ev_d2h));
        host_copy_mt(p_z_given_d, pin_a, bytes_a);
    }
    CK(cudaStreamSynchronize(ctx->stream));
    if (bounce_b) host_copy_mt(p_w_given_z, pin_b, bytes_b);
    return PLSA_OK;
}

/* ---- EM ------------------------------------------------------------------------------------------ */
/* work items of the doc pass (and of the term pass for a full fit), sized for kp */
static int ensure_items(plsa_ctx *ctx, bool refit, int kp, bool want_doc = true)
{
    const int64_t chunk = choose_chunk(ctx, kp);
    const int align = ctx->vec_entries ? pass_entry_block(kp) : 1;
    int rc;
    /* tiled mode: the doc items of the whole corpus serve only the exact log-likelihood pass
     * and are built when that pass first runs */
    const bool need_doc = want_doc &&
                          (!ctx->doc_items.ready || ctx->doc_items.chunk != chunk ||
                           ctx->doc_items.align != align);
    auto make = [&](const std::vector<int32_t> &h_ptr, const int32_t *d_ptr, int64_t rows, ItemSet &set) {
        return ctx->device_plan ? build_items_device(ctx, d_ptr, rows, set, chunk, align)
                                : build_items(ctx, h_ptr, rows, set, chunk, align);
    };
    if (need_doc &&
        (rc = make(ctx->cur().h_indptr, ctx->cur().indptr.as<int32_t>(), ctx->cur().n, ctx->doc_items)))
        return rc;
    if (!refit) {
        if (!ctx->t_ready && (rc = build_term_major(ctx))) return rc;
        if (!ctx->term_items.ready || ctx->term_items.chunk != chunk || ctx->term_items.align != align)
            if ((rc = make(ctx->h_tindptr, ctx->t_indptr.as<int32_t>(), ctx->cur().m, ctx->term_items)))
                return rc;
    }
    return PLSA_OK;
}

static int run_loglik(plsa_ctx *ctx, double *out)
{
    const Corpus &c = ctx->cur();
    int rc = ensure_items(ctx, true, ctx->kp);
    if (rc) return rc;
    if (!ctx->have_sw && (rc = plsa_set_sample_weight(ctx, nullptr))) return rc;
    const int64_t grid = pass_grid(ctx->doc_items.n_items, ctx->kp);
    if (grid == 0) {
        *out = 0.0;
        return PLSA_OK;
    }
    CK(ctx->ll_part.ensure((size_t)grid * 8));
    {
        ProfScope ps(ctx, PLSA_PROF_LOGLIK);
        PassArgs a{};
        a.items = ctx->doc_items.items.as<Item>();
        a.n_items = ctx->doc_items.n_items;
        a.ent = c.ent.as<int2>();
        a.own_old = ctx->A[ctx->curA].as<float>();
        a.gat_old = ctx->B[ctx->curB].as<float>();
        a.own_scale = cur_scale(ctx);
        a.row_weight = ctx->sw.as<float>();
        a.cta_partial = ctx->ll_part.as<double>();
        a.gat_tex = ctx->texB[ctx->curB];
        a.ticket = ctx->tickets.as<unsigned int>();
        a.ll_out = ctx->ll_out.as<double>();
        a.stride_own = ctx->strideA;
        a.stride_gat = ctx->strideB;
        a.kp = ctx->kp;
        if ((rc = launch_pass(ctx, MODE_LOGLIK, a, ctx->doc_items.align > 1))) return rc;
    }
    /* sharded fit: sum of the shards' log-likelihoods, same value on all ranks */
    if (ctx->shard && (rc = shard_allreduce(ctx, ctx->ll_out.p, 1, true, ctx->stream))) return rc;
    CK(cudaMemcpyAsync(out, ctx->ll_out.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return PLSA_OK;
}

/* Build everything a later plsa_em needs that depends only on the corpus (work items, and
 * for a full fit the term-major copy), so that it can overlap host-side initialisation. */
API int plsa_prepare(plsa_ctx *ctx, int32_t refit, int32_t k)
{
    CHECK_CTX(ctx);
    if (ctx->cur().h_indptr.empty()) return ctx->fail(PLSA_EINVAL, "prepare: no corpus uploaded");
    if (k < 1 || k > PLSA_MAX_K) return ctx->fail(PLSA_EINVAL, "prepare: k out of range");
    const int kp = (k + 3) / 4 * 4;
    const bool tiled = tiled_wanted(ctx, kp);
    int rc = ensure_items(ctx, refit != 0, kp, !tiled);
    if (rc == PLSA_OK && tiled) rc = ensure_tiles(ctx, kp);
    if (rc == PLSA_OK && tiled && !refit && ctx->term_tiled_opt != 0)
        rc = ensure_term_tiles(ctx, kp, false); /* rebuilt by plsa_em if sample weights are in use */
    return rc;
}

API int plsa_log_likelihood(plsa_ctx *ctx, double *ll)
{
    CHECK_CTX(ctx);
    if (!ll) return ctx->fail(PLSA_EINVAL, "log_likelihood: null output");
    if (!ctx->have_factors) return ctx->fail(PLSA_EINVAL, "log_likelihood: no factors set");
    return run_loglik(ctx, ll);
}

/* split rows of one factor (which = 0: P(z|d), normalised; 1: P(w|z)^T) on `stream` */
static int run_fixup(plsa_ctx *ctx, int which, float *own_new, cudaStream_t stream,
                     const ItemSet *set = nullptr, bool with_tiled_terms = false)
{
    const ItemSet &is = set ? *set : (which ? ctx->term_items : ctx->doc_items);
    if (is.n_split == 0 && !with_tiled_terms) return PLSA_OK;
    ProfScope ps(ctx, PLSA_PROF_FIXUP, stream);
    FixArgs f{}, none{};
    if (with_tiled_terms) { /* a tiled term's row = the sum of its per-block partials, in block order */
        const TermTiles &t = ctx->tterm;
        none.rows = t.tiled_terms.as<int32_t>();
        none.slot_begin = t.slot_begin.as<int32_t>();
        none.partial = t.partial.as<float>();
        none.own_new = own_new;
        none.n_split = t.n_tiled;
        none.n_heavy = t.n_blocks > 256 ? t.n_tiled : 0; /* a warp adds up to 256 slots per term quickly enough */
        none.kp = ctx->kp;
        none.stride_own = ctx->strideB;
        none.normalise = 0;
    }
    f.rows = is.split_rows.as<int32_t>();
    f.slot_begin = is.slot_begin.as<int32_t>();
    f.partial = which ? ctx->partialB.as<float>() : ctx->partialA.as<float>();
    f.own_new = own_new;
    f.n_split = is.n_split;
    f.n_heavy = std::min(is.n_heavy, is.n_split);
    f.kp = ctx->kp;
    f.stride_own = which ? ctx->strideB : ctx->strideA;
    f.normalise = which ? 0 : 1;
    const int blocks = f.n_heavy + (int)cdiv(f.n_split - f.n_heavy, 8) + none.n_heavy +
                       (int)cdiv(none.n_split - none.n_heavy, 8);
    fixup_kernel<<<(unsigned)blocks, 256, 0, stream>>>(f, none);
    ctx->launches++;
    CK(cudaGetLastError());
    return PLSA_OK;
}

API int plsa_em(plsa_ctx *ctx, int32_t n_iter, int32_t n_iter_per_test, double tolerance,
                float e_step_thresh, int32_t refit, int32_t use_sample_weights,
                int32_t *iters_run, double *ll_trace, int32_t ll_cap, int32_t *n_ll)
{
    CHECK_CTX(ctx);
    if (iters_run) *iters_run = 0;
    if (n_ll) *n_ll = 0;
    if (!ctx->have_factors) return ctx->fail(PLSA_EINVAL, "em: no factors set");
    if (n_iter < 0 || n_iter_per_test < 1)
        return ctx->fail(PLSA_EINVAL, "em: n_iter < 0 or n_iter_per_test < 1");
    Corpus &c = ctx->cur();
    int rc;
    const int kp = ctx->kp;
    /* tiled doc pass (plsa_tile.cuh): head entries from a shared-memory tile, tail entries
     * through the texture path; P(w|z) is then kept normalised in memory */
    const bool tiled = tiled_wanted(ctx, kp);
    if ((rc = ensure_items(ctx, refit != 0, kp, !tiled))) return rc;
    if (tiled && (rc = ensure_tiles(ctx, kp))) return rc;
    if (!ctx->have_sw && (rc = plsa_set_sample_weight(ctx, nullptr))) return rc;
    if (!refit && use_sample_weights && (rc = ensure_weighted_vals(ctx))) return rc;
    /* the term pass tiled as well: (term, document block) items of the frequent terms */
    if (tiled && !refit && ctx->term_tiled_opt != 0 &&
        (rc = ensure_term_tiles(ctx, kp, use_sample_weights != 0)))
        return rc;
    const bool term_tiled = tiled && !refit && ctx->term_tiled_opt != 0 && ctx->tterm.ready &&
                            ctx->tterm.n_items > 0;
    const ItemSet &doc_set = tiled ? ctx->tiles.tail_items : ctx->doc_items;
    const ItemSet &term_set = term_tiled ? ctx->tterm.tail_items : ctx->term_items;
    /* products at or below the threshold are dropped (plsa.py:98-102); subnormal products
     * are dropped as well so that a surviving posterior normaliser is never subnormal */
    e_step_thresh = std::max(e_step_thresh, 1.17549435e-38f);
    CK(ctx->partialA.ensure((size_t)std::max(doc_set.n_slots, 1) * kp * 4));
    if (tiled && !ctx->b_norm[ctx->curB]) {
        /* fold the pending column scale into B and build its tile image; from here on the
         * scale vectors are all ones */
        if ((rc = normalise_b(ctx, ctx->curB, cur_scale(ctx), ctx->stream))) return rc;
        if ((rc = fill(ctx, ctx->scale.as<float>(), 2 * kp, 1.f))) return rc;
        ctx->b_norm[ctx->curB] = true;
    }
    if (term_tiled && !ctx->a_comp[ctx->curA]) {
        if ((rc = compact_a(ctx, ctx->curA, ctx->stream))) return rc;
        ctx->a_comp[ctx->curA] = true;
    }
    /* the flush-to-zero scale of the tiled pass: S = FLT_MIN / thresh (>= thresh is kept) */
    const float ftz_scale = (float)(1.17549435e-38 / (double)e_step_thresh);
    if (!refit) CK(ctx->partialB.ensure((size_t)std::max(term_set.n_slots, 1) * kp * 4));

    /* plsa.py:913 — the refit loop's early stop is guarded by LL > 0, which a
     * log-likelihood never satisfies: LL is evaluated there only when a trace is asked for */
    const bool want_ll = !refit || ll_trace != nullptr;
    /* Fused evaluation: the log-likelihood "after iteration t" is the log-likelihood of the
     * factors iteration t+1 reads, so the doc pass of iteration t+1 returns it for free
     * (MODE_DOC_LL) instead of a separate gather pass.  Iteration t+1 is then launched
     * speculatively; if the test says stop, its output (in the ping-pong buffers) is simply
     * not adopted — the returned model is exactly the reference's (plsa.py:630-638). */
    const bool fuse = want_ll && !refit && ctx->fuse_ll && e_step_thresh <= PLSA_FUSED_LL_MAX_THRESH;
    if (fuse) {
        if (!ctx->mail) CK(cudaHostAlloc((void **)&ctx->mail, 32, cudaHostAllocDefault));
        CK(ctx->flag.ensure(4));
        CK(ctx->ll_part.ensure((size_t)std::max<int64_t>(pass_grid(doc_set.n_items, kp), 1) * 8));
    }
    if (!refit) { /* per-CTA column sums of the term pass and its last-arrival tickets */
        const int64_t tgrid = pass_grid(term_set.n_items, kp);
        CK(ctx->colpart.ensure((size_t)(tgrid + tgrid / 32 + 2) * kp * 8));
        if (ctx->tickets.cap < (size_t)(tgrid / 32 + 8) * 4) {
            CK(ctx->tickets.ensure((size_t)(tgrid / 32 + 8) * 4 * 2));
            CK(cudaMemsetAsync(ctx->tickets.p, 0, ctx->tickets.cap, ctx->stream));
        }
    }
    int32_t nl = 0;
    double prev = 0.0;
    bool stopped = false;
    /* plsa.py:632-637: returns true when the loop must stop */
    auto judge = [&](double cur) {
        if (ll_trace && nl < ll_cap) ll_trace[nl] = cur;
        nl++;
        if (!refit) {
            /* the reference holds the log-likelihood in float32 (plsa.py:322) */
            const float curf = (float)cur, prevf = (float)prev;
            const float change = fabsf(curf - prevf);
            if (change == 0.f || (double)(change / fabsf(curf)) < tolerance) return true;
            prev = cur;
        } else if (cur > 0.0) { /* plsa.py:913 */
            const float change = fabsf((float)cur - (float)prev);
            if ((double)(change / fabsf((float)cur)) < tolerance) return true;
            prev = cur;
        }
        return false;
    };
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    if (want_ll && (!fuse || n_iter == 0)) {
        if ((rc = run_loglik(ctx, &prev))) return rc; /* plsa.py:591 */
        if (ll_trace && nl < ll_cap) ll_trace[nl] = prev;
        nl++;
    }
    int32_t done = 0;
    /* The doc pass and the term pass of an iteration read the same old factors and write
     * different new ones: they run on two streams (with their split-row fixups behind them)
     * and join before the next iteration.  Per-kernel profiling keeps them serial. */
    const bool two = ctx->overlap && !refit && !ctx->profiling;
    cudaStream_t s1 = ctx->stream, s2 = two ? ctx->stream2 : ctx->stream;
    /* Document-sharded fit: the doc pass is local; the term pass leaves this shard's raw
     * P(w|z)^T sums, which are added over the ranks (all-reduce on s2, behind the term pass,
     * while the doc pass still runs on s1), then the column sums are taken from the complete
     * matrix.  Every collective of this context is issued on s2, in the same order on all
     * ranks; the log-likelihood is the sum of the shards' values, so all ranks take the
     * same stop decision. */
    const bool sharded = ctx->shard != nullptr; /* one rank: same path, empty collectives */
    const int colsum_grid = term_tiled ? 592 : 148; /* 4 CTAs per SM / the sharded fit's reduce kernel: 1 */
    if (sharded) CK(ctx->ll2.ensure(16));
    if (sharded || term_tiled) CK(ctx->colpart2.ensure((size_t)colsum_grid * kp * 8));
    /* Peer-memory path (plsa_shard_p2p_*): the term pass writes into this rank's exchange
     * buffer and ONE kernel per rank adds the ranks' buffers over NVLink and takes the column
     * sums (shard_reduce_kernel); otherwise NCCL all-reduce + separate column sums. */
    const int n_ranks = sharded ? shard_n_ranks(ctx) : 1;
    const bool p2p = sharded && !refit && ctx->p2p.enabled && ctx->p2p.block.p &&
                     ctx->p2p.n_attached == n_ranks - 1 && n_ranks <= SHARD_MAX_RANKS &&
                     ctx->p2p.part_bytes == (size_t)std::max<int64_t>(c.m, 1) * ctx->strideB * 4;
    bool p2p_used = false;
    if (p2p && n_iter > 0) {
        /* the ranks reach this point unsynchronised (corpus preparation, lazy module loads):
         * a one-word all-reduce on the exchange stream lines them up before the first peer wait */
        if ((rc = shard_allreduce(ctx, ctx->ll2.p, 1, true, s2))) return rc;
    }
    for (int32_t i = 0; i < n_iter; ++i) {
        const int nA = ctx->curA ^ 1, nB = ctx->curB ^ 1;
        const bool fused_now = fuse && (i == 0 || (i - 1) % n_iter_per_test == 0);
        if (two) { /* s2 starts once everything issued so far on s1 (previous join) is done */
            CK(cudaEventRecord(ctx->ev_go, s1));
            CK(cudaStreamWaitEvent(s2, ctx->ev_go, 0));
        }
        {   /* E-step + M-step of P(z|d): plsa.py:91-105, :189-194 (P(z|d) part), :199-202 */
            ProfScope ps(ctx, PLSA_PROF_DOC_PASS);
            if (tiled) { /* head entries: shared-memory tile, TMA-staged, lane per entry */
                TileSet &t = ctx->tiles;
                TileArgs h{};
                h.items = t.headers.as<int4>();
                h.ent = t.head_ent.as<int2>();
                h.own_old = ctx->A[ctx->curA].as<float>();
                h.tile_src = t.img[ctx->curB].as<float>();
                h.partial_out = t.head_sum.as<float>();
                h.n_items = c.n;
                h.src_rows = t.tile_rows;
                h.block_rows = t.tile_rows;
                h.n_blocks = 1;
                h.stride_own = ctx->strideA;
                h.kp = kp;
                h.ftz_scale = ftz_scale;
                h.inv_ftz_scale = 1.f / ftz_scale;
                h.log2_ftz_corr = std::log2((double)ftz_scale * (double)h.inv_ftz_scale);
                if (fused_now) {
                    h.row_weight = ctx->sw.as<float>();
                    h.cta_partial = t.ll_part.as<double>();
                    h.ticket = t.ll_ticket.as<unsigned int>();
                    h.ll_out = t.ll_head.as<double>();
                    h.flag = ctx->flag.as<int>();
                    CK(cudaMemsetAsync(ctx->flag.p, 0, 4, ctx->stream));
                }
                const size_t smem = (size_t)std::max<int32_t>(t.tile_rows, TILE_MIN_ROWS) * t.pitch_f * 4;
                ProfScope ph(ctx, PLSA_PROF_DOC_HEAD);
                pick_tile_kernel(kp, fused_now)<<<ctx->n_sms, TILE_THREADS, smem, ctx->stream>>>(h);
                ctx->launches++;
                CK(cudaGetLastError());
            }
            PassArgs a{};
            a.items = doc_set.items.as<Item>();
            a.n_items = doc_set.n_items;
            a.ent = tiled ? ctx->tiles.tail_ent.as<int2>() : c.ent.as<int2>();
            a.add_partial = tiled ? ctx->tiles.head_sum.as<float>() : nullptr;
            a.own_old = ctx->A[ctx->curA].as<float>();
            a.gat_old = ctx->B[ctx->curB].as<float>();
            a.own_scale = cur_scale(ctx);
            a.own_new = ctx->A[nA].as<float>();
            a.partial = ctx->partialA.as<float>();
            a.gat_tex = ctx->texB[ctx->curB];
            a.stride_own = ctx->strideA;
            a.stride_gat = ctx->strideB;
            a.kp = kp;
            a.thresh = e_step_thresh;
            if (fused_now) {
                a.row_weight = ctx->sw.as<float>();
                a.cta_partial = ctx->ll_part.as<double>();
                a.ticket = ctx->tickets.as<unsigned int>();
                a.ll_out = ctx->ll_out.as<double>();
                a.flag = ctx->flag.as<int>();
                if (!tiled) CK(cudaMemsetAsync(ctx->flag.p, 0, 4, ctx->stream));
            }
            if ((rc = launch_pass(ctx, fused_now ? MODE_DOC_LL : MODE_DOC, a, doc_set.align > 1)))
                return rc;
            if (fused_now && !sharded) {
                CK(cudaMemcpyAsync(&ctx->mail[0], ctx->ll_out.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
                CK(cudaMemcpyAsync(&ctx->mail[1], ctx->flag.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
                if (tiled)
                    CK(cudaMemcpyAsync(&ctx->mail[2], ctx->tiles.ll_head.p, 8, cudaMemcpyDeviceToHost,
                                       ctx->stream));
                CK(cudaEventRecord(ctx->ev_ll, ctx->stream));
            }
            if (fused_now && sharded) CK(cudaEventRecord(ctx->ev_a, s1)); /* doc pass done */
        }
        if ((rc = run_fixup(ctx, 0, ctx->A[nA].as<float>(), s1, &doc_set))) return rc;
        if (term_tiled && (rc = compact_a(ctx, nA, s1))) return rc; /* next iteration's tiles */
        if (!refit) {
            {   /* E-step + M-step of P(w|z): plsa.py:91-105, :189-193 (P(w|z) part) */
                ProfScope ps(ctx, PLSA_PROF_WORD_PASS, s2);
                if (term_tiled) { /* frequent terms: P(z|d) blocks in shared memory */
                    TermTiles &t = ctx->tterm;
                    TileArgs h{};
                    h.items = t.headers.as<int4>();
                    h.ent = t.head_ent.as<int2>();
                    h.own_old = ctx->B[ctx->curB].as<float>();
                    h.tile_src = t.acomp[ctx->curA].as<float>();
                    h.partial_out = t.partial.as<float>();
                    h.cta_begin = t.cta_begin.as<int32_t>();
                    h.block_begin = t.block_begin.as<int32_t>();
                    h.n_items = t.n_items;
                    h.src_rows = c.n;
                    h.block_rows = t.block_rows;
                    h.n_blocks = t.n_blocks;
                    h.stride_own = ctx->strideB;
                    h.kp = kp;
                    h.ftz_scale = ftz_scale;
                    const size_t smem = (size_t)std::max<int32_t>(t.block_rows, TILE_MIN_ROWS) * t.pitch_f * 4;
                    ProfScope ph(ctx, PLSA_PROF_TERM_HEAD, s2);
                    pick_tile_kernel(kp, false)<<<t.grid, TILE_THREADS, smem, s2>>>(h);
                    ctx->launches++;
                    CK(cudaGetLastError());
                }
                PassArgs a{};
                a.items = term_set.items.as<Item>();
                a.n_items = term_set.n_items;
                a.ent = use_sample_weights ? ctx->t_entw.as<int2>() : ctx->t_ent.as<int2>();
                a.own_old = ctx->B[ctx->curB].as<float>();
                a.gat_old = ctx->A[ctx->curA].as<float>();
                a.own_scale = cur_scale(ctx);
                float *term_out = ctx->B[nB].as<float>();
                if (p2p) /* exchange buffer (seq + 1) & 1, read by the peers */
                    term_out = reinterpret_cast<float *>(ctx->p2p.block.as<char>() +
                                                         ((ctx->p2p.seq + 1) & 1u) * ctx->p2p.part_bytes);
                a.own_new = term_out;
                a.partial = ctx->partialB.as<float>();
                a.gat_tex = ctx->texA[ctx->curA];
                /* plsa.py:196-198: per-topic normaliser of the new P(w|z), applied lazily */
                /* (tiled term pass: the column sums are taken from the finished matrix instead) */
                a.cta_partial = term_tiled ? nullptr : ctx->colpart.as<double>();
                a.ticket = ctx->tickets.as<unsigned int>() + 1; /* [0] is the log-likelihood's */
                a.scale_out = tiled ? ctx->tiles.scale_raw.as<float>() + (size_t)nB * kp
                                    : reinterpret_cast<float *>(ctx->scale.p) + (size_t)nB * kp;
                a.colnorm_out = ctx->colnorm.as<double>();
                a.stride_own = ctx->strideB;
                a.stride_gat = ctx->strideA;
                a.kp = kp;
                a.thresh = e_step_thresh;
                if ((rc = launch_pass(ctx, MODE_TERM, a, term_set.align > 1, s2))) return rc;
            }
            /* split rows: ordered sums of their chunk partials (and, tiled, of the tiled terms'
             * per-block partials) */
            if ((rc = run_fixup(ctx, 1,
                                p2p ? reinterpret_cast<float *>(ctx->p2p.block.as<char>() +
                                                                ((ctx->p2p.seq + 1) & 1u) * ctx->p2p.part_bytes)
                                    : ctx->B[nB].as<float>(), s2, &term_set, term_tiled)))
                return rc;
            if (term_tiled && c.m > 0) { /* plsa.py:196-198: column sums of the finished matrix */
                ProfScope ps(ctx, PLSA_PROF_NORMALIZE, s2);
                colsum_partial_kernel<<<colsum_grid, 256, 0, s2>>>(ctx->B[nB].as<float>(), c.m, ctx->strideB, kp,
                                                                  ctx->colpart2.as<double>());
                colsum_final_kernel<<<1, 256, 0, s2>>>(ctx->colpart2.as<double>(), colsum_grid, kp,
                                                       ctx->tiles.scale_raw.as<float>() + (size_t)nB * kp,
                                                       ctx->colnorm.as<double>());
                ctx->launches += 2;
                CK(cudaGetLastError());
            }
            if (tiled) { /* plsa.py:196-198 applied now, not lazily: B[nB] *= scale, + tile image */
                ProfScope ps(ctx, PLSA_PROF_NORMALIZE, s2);
                if ((rc = normalise_b(ctx, nB, ctx->tiles.scale_raw.as<float>() + (size_t)nB * kp, s2)))
                    return rc;
            }
            if (p2p && c.m > 0) {
                ProfScope ps(ctx, PLSA_PROF_NORMALIZE, s2);
                ctx->p2p.seq += 1;
                p2p_used = true;
                const int me = shard_rank(ctx);
                const size_t sig_off = 2 * ctx->p2p.part_bytes + 2 * ctx->p2p.red_bytes;
                const bool two_shot = ctx->p2p.two_shot > 0 || (ctx->p2p.two_shot < 0 && n_ranks >= 4);
                if (two_shot) { /* each rank adds its slice of the rows, the finished slices are exchanged */
                    ShardTwoShotArgs ta{};
                    for (int p = 0; p < n_ranks; ++p) {
                        char *base = (p == me) ? ctx->p2p.block.as<char>() : (char *)ctx->p2p.peer_base[p];
                        ta.part[p] = reinterpret_cast<const float *>(base + (ctx->p2p.seq & 1u) * ctx->p2p.part_bytes);
                        ta.red[p] = reinterpret_cast<float *>(base + 2 * ctx->p2p.part_bytes +
                                                              (ctx->p2p.seq & 1u) * ctx->p2p.red_bytes);
                        ta.peer_sig[p] = reinterpret_cast<unsigned int *>(base + sig_off);
                        ta.peer_sig2[p] = reinterpret_cast<unsigned int *>(base + sig_off + 64);
                    }
                    ta.my_sig = reinterpret_cast<volatile unsigned int *>(ctx->p2p.block.as<char>() + sig_off);
                    ta.my_sig2 = reinterpret_cast<volatile unsigned int *>(ctx->p2p.block.as<char>() + sig_off + 64);
                    ta.out = ctx->B[nB].as<float>();
                    ta.colpart = ctx->colpart2.as<double>();
                    ta.err = ctx->p2p.err.as<int>();
                    ta.rows = c.m;
                    ta.slice_rows = cdiv(std::max<int64_t>(c.m, 1), n_ranks);
                    ta.stride = ctx->strideB;
                    ta.kp = kp;
                    ta.n_ranks = n_ranks;
                    ta.rank = me;
                    ta.seq = ctx->p2p.seq;
                    ta.timeout_clocks = (long long)ctx->p2p.timeout_ms * 2000000LL;
                    shard_slice_reduce_kernel<<<colsum_grid, 256, 0, s2>>>(ta);
                    shard_slice_gather_kernel<<<colsum_grid, 256, 0, s2>>>(ta);
                    colsum_final_kernel<<<1, 256, 0, s2>>>(
                        ctx->colpart2.as<double>(), colsum_grid, kp,
                        reinterpret_cast<float *>(ctx->scale.p) + (size_t)nB * kp, ctx->colnorm.as<double>());
                    ctx->launches += 3;
                    CK(cudaGetLastError());
                } else {
                ShardReduceArgs ra{};
                char *sig0 = nullptr;
                for (int p = 0; p < n_ranks; ++p) {
                    char *base = (p == me) ? ctx->p2p.block.as<char>() : (char *)ctx->p2p.peer_base[p];
                    ra.part[p] = reinterpret_cast<const float *>(base + (ctx->p2p.seq & 1u) * ctx->p2p.part_bytes);
                    ra.peer_sig[p] = reinterpret_cast<unsigned int *>(base + sig_off);
                    if (p == me) sig0 = base + sig_off;
                }
                ra.my_sig = reinterpret_cast<volatile unsigned int *>(sig0);
                ra.out = ctx->B[nB].as<float>();
                ra.colpart = ctx->colpart2.as<double>();
                ra.err = ctx->p2p.err.as<int>();
                ra.rows = c.m;
                ra.stride = ctx->strideB;
                ra.kp = kp;
                ra.n_ranks = n_ranks;
                ra.rank = me;
                ra.seq = ctx->p2p.seq;
                ra.timeout_clocks = (long long)ctx->p2p.timeout_ms * 2000000LL; /* ~2 GHz SM clock */
                shard_reduce_kernel<<<colsum_grid, 256, 0, s2>>>(ra);
                colsum_final_kernel<<<1, 256, 0, s2>>>(
                    ctx->colpart2.as<double>(), colsum_grid, kp,
                    reinterpret_cast<float *>(ctx->scale.p) + (size_t)nB * kp, ctx->colnorm.as<double>());
                ctx->launches += 2;
                CK(cudaGetLastError());
                }
            } else if (sharded) {
                ProfScope ps(ctx, PLSA_PROF_NORMALIZE, s2);
                if ((rc = shard_allreduce(ctx, ctx->B[nB].p, (size_t)c.m * ctx->strideB, false, s2)))
                    return rc;
                if (c.m > 0) {
                    colsum_partial_kernel<<<colsum_grid, 256, 0, s2>>>(
                        ctx->B[nB].as<float>(), c.m, ctx->strideB, kp, ctx->colpart2.as<double>());
                    colsum_final_kernel<<<1, 256, 0, s2>>>(
                        ctx->colpart2.as<double>(), colsum_grid, kp,
                        reinterpret_cast<float *>(ctx->scale.p) + (size_t)nB * kp,
                        ctx->colnorm.as<double>());
                    ctx->launches += 2;
                    CK(cudaGetLastError());
                }
            }
            if (fused_now && sharded) { /* {log-likelihood, flag} summed over the shards */
                CK(cudaStreamWaitEvent(s2, ctx->ev_a, 0));
                pack_ll_kernel<<<1, 1, 0, s2>>>(ctx->ll_out.as<double>(), ctx->flag.as<int>(),
                                                ctx->ll2.as<double>());
                ctx->launches++;
                CK(cudaGetLastError());
                if ((rc = shard_allreduce(ctx, ctx->ll2.p, 2, true, s2))) return rc;
                CK(cudaMemcpyAsync(ctx->mail, ctx->ll2.p, 16, cudaMemcpyDeviceToHost, s2));
                CK(cudaEventRecord(ctx->ev_ll, s2));
            }
            if (two) {
                CK(cudaEventRecord(ctx->ev_b, s2));
                CK(cudaStreamWaitEvent(s1, ctx->ev_b, 0)); /* join */
            }
        }
        if (fused_now) {
            CK(cudaEventSynchronize(ctx->ev_ll));
            double v;
            int bad;
            memcpy(&v, &ctx->mail[0], 8);
            if (sharded) bad = ctx->mail[1] != 0.0; /* any shard's flag */
            else memcpy(&bad, &ctx->mail[1], 4);
            if (tiled) v += ctx->mail[2]; /* tail entries + head entries */
            if (bad && (rc = run_loglik(ctx, &v))) return rc; /* exact pass on the same factors */
            if (i == 0) { /* plsa.py:591, the value before the loop */
                prev = v;
                if (ll_trace && nl < ll_cap) ll_trace[nl] = v;
                nl++;
            } else if (judge(v)) { /* the test of iteration i-1: iteration i is not adopted */
                stopped = true;
                break;
            }
        }
        if (!refit) {
            ctx->curB = nB;
            ctx->b_norm[nB] = tiled;
        }
        ctx->curA = nA;
        ctx->a_comp[nA] = term_tiled;
        done = i + 1;
        if (want_ll && !fuse && i % n_iter_per_test == 0) { /* plsa.py:630-638 / :909-918 */
            double cur = 0.0;
            if ((rc = run_loglik(ctx, &cur))) return rc;
            if (judge(cur)) break;
        }
    }
    if (fuse && !stopped && n_iter > 0 && (n_iter - 1) % n_iter_per_test == 0) {
        double cur = 0.0; /* the test after the last iteration: no later pass to ride on */
        if ((rc = run_loglik(ctx, &cur))) return rc;
        judge(cur);
    }
    if (p2p_used) {
        /* no rank may return (and perhaps free or rewrite its exchange buffer) while a peer's
         * last reduce kernel still reads it: a one-word all-reduce is the closing barrier */
        if ((rc = shard_allreduce(ctx, ctx->ll2.p, 1, true, ctx->stream))) return rc;
    }
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaEventElapsedTime(&ctx->last_em_ms, ctx->ev0, ctx->ev1));
    prof_collect(ctx);
    if (p2p_used) {
        int bad = 0;
        CK(cudaMemcpy(&bad, ctx->p2p.err.p, 4, cudaMemcpyDeviceToHost));
        if (bad) {
            CK(cudaMemset(ctx->p2p.err.p, 0, 4));
            return ctx->fail(PLSA_ENCCL, "sharded fit: a peer GPU did not signal within the "
                                         "time limit (peer-memory all-reduce)");
        }
    }
    if (iters_run) *iters_run = done;
    if (n_ll) *n_ll = nl;
    return PLSA_OK;
}

/* ---- measurement ------------------------------------------------------------------------------------ */
API int plsa_last_em_ms(const plsa_ctx *ctx, float *ms)
{
    if (!ctx || !ms) return PLSA_EINVAL;
    *ms = ctx->last_em_ms;
    return PLSA_OK;
}

API int plsa_set_profiling(plsa_ctx *ctx, int32_t on)
{
    if (!ctx) return PLSA_EINVAL;
    ctx->profiling = on != 0;
    for (int i = 0; i < PLSA_PROF_SLOTS; ++i) {
        ctx->prof_ms[i] = 0.0;
        ctx->prof_n[i] = 0;
    }
    return PLSA_OK;
}

API int plsa_get_profile(plsa_ctx *ctx, double *ms, int64_t *launches)
{
    if (!ctx) return PLSA_EINVAL;
    for (int i = 0; i < PLSA_PROF_SLOTS; ++i) {
        if (ms) ms[i] = ctx->prof_ms[i];
        if (launches) launches[i] = ctx->prof_n[i];
    }
    return PLSA_OK;
}

API int plsa_launch_count(const plsa_ctx *ctx, int64_t *launches)
{
    if (!ctx || !launches) return PLSA_EINVAL;
    *launches = ctx->launches;
    return PLSA_OK;
}

API int plsa_set_option(plsa_ctx *ctx, const char *name, int64_t value)
{
    if (!ctx || !name) return PLSA_EINVAL;
    if (!strcmp(name, "chunk")) {
        if (value != 0 && (value < 32 || value > ENT_PAD - ENT_SLACK))
            return ctx->fail(PLSA_EINVAL, "chunk out of range (32..4096)");
        ctx->chunk_user = value; /* item sets are rebuilt on the next prepare / em */
        return PLSA_OK;
    }
    if (!strcmp(name, "overlap")) {
        ctx->overlap = value != 0;
        return PLSA_OK;
    }
    if (!strcmp(name, "fuse_ll")) {
        ctx->fuse_ll = value != 0;
        return PLSA_OK;
    }
    if (!strcmp(name, "texture")) {
        ctx->use_texture = value != 0;
        return PLSA_OK;
    }
    if (!strcmp(name, "p2p")) { /* sharded fit: 0 = NCCL all-reduce even when peers are attached */
        ctx->p2p.enabled = value != 0;
        return PLSA_OK;
    }
    if (!strcmp(name, "presort")) { /* 1: plsa_upload_csr sorts the entries by term while the values upload */
        ctx->presort_opt = value != 0;
        return PLSA_OK;
    }
    if (!strcmp(name, "p2p_two_shot")) { /* -1: from 4 ranks up, 0: one-shot exchange, 1: two-shot */
        ctx->p2p.two_shot = value < 0 ? -1 : (value != 0);
        return PLSA_OK;
    }
    if (!strcmp(name, "p2p_timeout_ms")) { /* sharded fit: how long a rank waits for a peer's partial sums */
        if (value < 1 || value > 3600000) return ctx->fail(PLSA_EINVAL, "p2p_timeout_ms: 1..3600000");
        ctx->p2p.timeout_ms = value;
        return PLSA_OK;
    }
    if (!strcmp(name, "device_plan")) { /* 0: plan the work items on the host (the CPU-testable twin) */
        ctx->device_plan = value != 0;
        ctx->doc_items.ready = ctx->term_items.ready = ctx->tiles.tail_items.ready = false;
        return PLSA_OK;
    }
    if (!strcmp(name, "vec_entries")) {
        ctx->vec_entries = value != 0; /* item sets are rebuilt on the next prepare / em */
        return PLSA_OK;
    }
    if (!strcmp(name, "tiled")) { /* -1 auto (by corpus size), 0 never, 1 wherever possible */
        if (value < -1 || value > 1) return ctx->fail(PLSA_EINVAL, "tiled: -1, 0 or 1");
        ctx->tiled_opt = (int)value;
        return PLSA_OK;
    }
    if (!strcmp(name, "term_tiled")) { /* with "tiled": the term pass reads P(z|d) blocks from shared memory too */
        ctx->term_tiled_opt = value != 0;
        return PLSA_OK;
    }
    if (!strcmp(name, "term_tile_min")) { /* average entries per (term, block) item for a term to be tiled */
        if (value < 1 || value > 100000) return ctx->fail(PLSA_EINVAL, "term_tile_min: 1..100000");
        ctx->term_tile_min = value;
        ctx->tterm.ready = false;
        return PLSA_OK;
    }
    if (!strcmp(name, "tile_kb")) { /* shared memory of the tile, per CTA */
        if (value < 1 || value > 220) return ctx->fail(PLSA_EINVAL, "tile_kb: 1..220");
        ctx->tile_bytes = value * 1024;
        ctx->tiles.ready = false;
        ctx->tterm.ready = false;
        return PLSA_OK;
    }
    return ctx->fail(PLSA_EINVAL, std::string("unknown option: ") + name);
}

/* Host-only view of the work-item plan of a pass (no device needed): lets the CPU tests check
 * that the items cover every stored entry exactly once. */
API int plsa_plan_items(const int32_t *indptr, int64_t rows, int64_t chunk, int32_t align,
                        int64_t cap, int64_t *start, int32_t *row, int32_t *len,
                        int32_t *slot, int32_t *skip, int64_t *n_items, int32_t *n_split,
                        int32_t *n_slots)
{
    if (!indptr || rows < 0 || chunk < 1 || align < 1 || (align & (align - 1)) || chunk < align ||
        !n_items) {
        g_err = "plan_items: bad argument";
        return PLSA_EINVAL;
    }
    ItemPlan plan;
    try {
        plan_items(indptr, rows, chunk, align, plan);
    } catch (const std::bad_alloc &) {
        g_err = "plan_items: out of host memory";
        return PLSA_ENOMEM;
    }
    *n_items = (int64_t)plan.sorted.size();
    if (n_split) *n_split = (int32_t)plan.split_rows.size();
    if (n_slots) *n_slots = plan.slots;
    const int64_t w = std::min<int64_t>(cap, *n_items);
    for (int64_t i = 0; i < w; ++i) {
        const Item &it = plan.sorted[(size_t)i];
        if (start) start[i] = it.start;
        if (row) row[i] = it.row;
        if (len) len[i] = it.len;
        if (slot) slot[i] = it.slot;
        if (skip) skip[i] = it.skip & 0xffff;
    }
    return PLSA_OK;
}

/* Read back the work items a context holds on the device (which: 0 doc pass, 1 term pass,
 * 2 tail part of the tiled doc pass) together with the row pointers they were planned from,
 * so that a test can compare the device planner with plsa_plan_items. */
API int plsa_debug_items(plsa_ctx *ctx, int32_t which, int64_t cap, int64_t *start, int32_t *row,
                         int32_t *len, int32_t *slot, int32_t *skip, int64_t *n_items, int32_t *n_split,
                         int32_t *n_slots, int64_t *chunk, int32_t *align, int32_t *indptr_out,
                         int64_t indptr_cap)
{
    CHECK_CTX(ctx);
    const ItemSet &is = which == 0 ? ctx->doc_items : which == 1 ? ctx->term_items : ctx->tiles.tail_items;
    if (!is.ready || !n_items) return ctx->fail(PLSA_EINVAL, "debug_items: item set not built");
    *n_items = is.n_items;
    if (n_split) *n_split = is.n_split;
    if (n_slots) *n_slots = is.n_slots;
    if (chunk) *chunk = is.chunk;
    if (align) *align = is.align;
    const int64_t w = std::min<int64_t>(cap, is.n_items);
    if (w > 0) {
        std::vector<Item> h;
        try {
            h.resize((size_t)w);
        } catch (const std::bad_alloc &) {
            return ctx->fail(PLSA_ENOMEM, "debug_items: out of host memory");
        }
        CK(cudaMemcpy(h.data(), is.items.p, (size_t)w * sizeof(Item), cudaMemcpyDeviceToHost));
        for (int64_t i = 0; i < w; ++i) {
            if (start) start[i] = h[(size_t)i].start;
            if (row) row[i] = h[(size_t)i].row;
            if (len) len[i] = h[(size_t)i].len;
            if (slot) slot[i] = h[(size_t)i].slot;
            if (skip) skip[i] = h[(size_t)i].skip & 0xffff;
        }
    }
    if (indptr_out && indptr_cap > 0) {
        const DevBuf &src = which == 0 ? ctx->cur().indptr : which == 1 ? ctx->t_indptr : ctx->tiles.tail_indptr;
        const int64_t rows = which == 1 ? ctx->cur().m : ctx->cur().n;
        CK(cudaMemcpy(indptr_out, src.p, (size_t)std::min<int64_t>(indptr_cap, rows + 1) * 4,
                      cudaMemcpyDeviceToHost));
    }
    return PLSA_OK;
}

/* ---- one-shot drop-ins ---------------------------------------------------------------------------------- */
static int one_shot(const int32_t *rows, const int32_t *cols, const float *vals, int64_t nnz,
                    float *pwz_inout, const float *topics, float *pzd, const float *sw, int64_t n,
                    int64_t m, int32_t k, int32_t n_iter, int32_t per_test, double tol,
                    float thresh, int32_t use_sw, int32_t refit, int32_t device, int32_t *iters)
{
    plsa_ctx *ctx = nullptr;
    int rc = plsa_ctx_create(device, &ctx);
    if (rc) return rc;
    auto bail = [&](int code) {
        g_err = ctx->err;
        plsa_ctx_destroy(ctx);
        return code;
    };
    if ((rc = plsa_upload_coo(ctx, rows, cols, vals, n, m, nnz))) return bail(rc);
    if ((rc = plsa_set_factors(ctx, pzd, refit ? topics : pwz_inout, k))) return bail(rc);
    if ((rc = plsa_set_sample_weight(ctx, sw))) return bail(rc);
    if ((rc = plsa_em(ctx, n_iter, per_test, tol, thresh, refit, use_sw, iters, nullptr, 0, nullptr)))
        return bail(rc);
    if ((rc = plsa_get_factors(ctx, pzd, refit ? nullptr : pwz_inout))) return bail(rc);
    plsa_ctx_destroy(ctx);
    return PLSA_OK;
}

API int plsa_b200_fit_inner(const int32_t *X_rows, const int32_t *X_cols, const float *X_vals,
                            int64_t nnz, float *p_w_given_z, float *p_z_given_d,
                            const float *sample_weight, int64_t n_docs, int64_t n_terms,
                            int32_t k, int32_t n_iter, int32_t n_iter_per_test, double tolerance,
                            float e_step_thresh, int32_t use_sample_weights, int32_t device,
                            int32_t *iters_run)
{
    return one_shot(X_rows, X_cols, X_vals, nnz, p_w_given_z, nullptr, p_z_given_d, sample_weight,
                    n_docs, n_terms, k, n_iter, n_iter_per_test, tolerance, e_step_thresh,
                    use_sample_weights, 0, device, iters_run);
}

API int plsa_b200_refit_inner(const int32_t *X_rows, const int32_t *X_cols, const float *X_vals,
                              int64_t nnz, const float *topics, float *p_z_given_d,
                              const float *sample_weight, int64_t n_docs, int64_t n_terms,
                              int32_t k, int32_t n_iter, int32_t n_iter_per_test,
                              double tolerance, float e_step_thresh, int32_t device,
                              int32_t *iters_run)
{
    return one_shot(X_rows, X_cols, X_vals, nnz, nullptr, topics, p_z_given_d, sample_weight,
                    n_docs, n_terms, k, n_iter, n_iter_per_test, tolerance, e_step_thresh, 0, 1,
                    device, iters_run);
}

/* ---- all-pairs distances between topic vectors (ensemble clustering input) ------------------------ */
static thread_local float g_distances_kernel_ms = 0.f; /* device time of the last call's kernels */

API int plsa_last_distances_ms(float *kernel_ms)
{
    if (!kernel_ms) return PLSA_EINVAL;
    *kernel_ms = g_distances_kernel_ms;
    return PLSA_OK;
}

/* distances of a [n_topics, n_terms] matrix that already lives on `device` (host_src: uploaded
 * first); out is a host array [n_topics, n_topics] */
static int topic_distances_impl(int32_t device, const float *dev_src, const float *host_src,
                                int64_t n_topics, int64_t n_terms, int32_t kind, double *out)
{
    if ((!dev_src && !host_src) || !out || n_topics < 0 || n_terms < 0 || (kind != 0 && kind != 1) ||
        n_topics >= ((int64_t)1 << 24)) {
        g_err = "topic_distances: bad arguments";
        return PLSA_EINVAL;
    }
    if (n_topics == 0) return PLSA_OK;
    cudaError_t e = cudaSetDevice(device);
    DevBuf P, A, B, l1, D, part;
    cudaStream_t st = nullptr;
    const size_t cells = (size_t)n_topics * (size_t)std::max<int64_t>(n_terms, 1);
    auto done = [&](int rc, const char *what) {
        if (rc != PLSA_OK) g_err = std::string("topic_distances: ") + what + ": " + cudaGetErrorString(e);
        P.release(); A.release(); B.release(); l1.release(); D.release(); part.release();
        if (st) cudaStreamDestroy(st);
        return rc;
    };
    if (e != cudaSuccess) return done(PLSA_ECUDA, "cudaSetDevice");
    if ((e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking)) != cudaSuccess)
        return done(PLSA_ECUDA, "stream");
    /* term slices: enough CTAs to fill the GPU a few times over */
    const int64_t tiles = cdiv(n_topics, 32) * cdiv(n_topics, 32);
    int n_sms = 148;
    cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, device);
    const int slices = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(64, cdiv(n_terms, 1024)),
                                                                   cdiv(4 * (int64_t)n_sms, tiles)));
    if ((!dev_src && (e = P.ensure(cells * 4)) != cudaSuccess) || (e = A.ensure(cells * 4)) != cudaSuccess ||
        (e = B.ensure(kind == 1 ? cells * 4 : 16)) != cudaSuccess ||
        (e = l1.ensure((size_t)n_topics * 8)) != cudaSuccess ||
        (e = D.ensure((size_t)n_topics * n_topics * 8)) != cudaSuccess ||
        (e = part.ensure((size_t)slices * n_topics * n_topics * 8)) != cudaSuccess)
        return done(PLSA_ENOMEM, "allocation");
    const float *src = dev_src;
    if (!dev_src) {
        if (n_terms > 0 &&
            (e = cudaMemcpyAsync(P.p, host_src, cells * 4, cudaMemcpyHostToDevice, st)) != cudaSuccess)
            return done(PLSA_ECUDA, "upload");
        src = P.as<float>();
    }
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEventCreate(&ev0);
    cudaEventCreate(&ev1);
    cudaEventRecord(ev0, st);
    topic_rowsum_kernel<<<(unsigned)n_topics, 256, 0, st>>>(src, n_terms, l1.as<double>());
    if (n_terms > 0)
        topic_prep_kernel<<<(unsigned)cdiv((int64_t)cells, 256), 256, 0, st>>>(
            src, l1.as<double>(), n_topics, n_terms, kind, A.as<float>(), B.as<float>());
    const dim3 grid((unsigned)cdiv(n_topics, 32), (unsigned)cdiv(n_topics, 32), (unsigned)slices);
    if (kind == 0)
        topic_pairs_kernel<0><<<grid, 256, 0, st>>>(A.as<float>(), B.as<float>(), (int)n_topics, n_terms,
                                                    part.as<double>());
    else
        topic_pairs_kernel<1><<<grid, 256, 0, st>>>(A.as<float>(), B.as<float>(), (int)n_topics, n_terms,
                                                    part.as<double>());
    topic_pairs_finish_kernel<<<(unsigned)cdiv(n_topics * n_topics, 256), 256, 0, st>>>(
        part.as<double>(), slices, l1.as<double>(), (int)n_topics, kind, D.as<double>());
    cudaEventRecord(ev1, st);
    if ((e = cudaGetLastError()) == cudaSuccess)
        e = cudaMemcpyAsync(out, D.p, (size_t)n_topics * n_topics * 8, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e == cudaSuccess) cudaEventElapsedTime(&g_distances_kernel_ms, ev0, ev1);
    cudaEventDestroy(ev0);
    cudaEventDestroy(ev1);
    if (e != cudaSuccess) return done(PLSA_ECUDA, "kernels / download");
    return done(PLSA_OK, "");
}

API int plsa_topic_distances(int32_t device, const float *topics, int64_t n_topics, int64_t n_terms,
                             int32_t kind, double *out)
{
    if (!topics) {
        g_err = "topic_distances: bad arguments";
        return PLSA_EINVAL;
    }
    return topic_distances_impl(device, nullptr, topics, n_topics, n_terms, kind, out);
}

/* The same on the stack of topic matrices the last gather left on this context's device
 * (plsa_gather_topics / plsa_comm_gather_topics keep it there): no second trip over PCIe for the
 * 64 MB the clustering stage would otherwise upload again. */
API int plsa_gathered_distances(plsa_ctx *ctx, int32_t kind, double *out, int64_t *n_topics)
{
    CHECK_CTX(ctx);
    if (!ctx->gathered.p || ctx->gathered_n <= 0)
        return ctx->fail(PLSA_EINVAL, "gathered_distances: no gathered topics on this context");
    if (n_topics) *n_topics = ctx->gathered_n;
    if (!out) return PLSA_OK; /* size query */
    int rc = topic_distances_impl(ctx->device, ctx->gathered.as<float>(), nullptr, ctx->gathered_n,
                                  ctx->gathered_m, kind, out);
    if (rc) ctx->err = g_err;
    return rc;
}

/* ---- ensemble topic stash + gather over NCCL ------------------------------------------------------ */
/* Each finished ensemble member leaves its P(w|z) [k, m] (reference layout) in a slot of the
 * context's device-side stash; one gather at the end moves every slot to the root
 * (enstop_.py:231 np.vstack(topics)).  NCCL is resolved at run time (dlopen) so the library
 * loads on hosts without it; a single rank / single context never touches NCCL. */
API int plsa_stash_topics(plsa_ctx *ctx, int32_t slot, int32_t n_slots)
{
    CHECK_CTX(ctx);
    if (!ctx->have_factors) return ctx->fail(PLSA_EINVAL, "stash_topics: no factors set");
    if (n_slots < 1 || slot < 0 || slot >= n_slots)
        return ctx->fail(PLSA_EINVAL, "stash_topics: slot out of range");
    const int64_t m = ctx->cur().m;
    const size_t per = (size_t)std::max<int64_t>(m, 1) * ctx->k;
    if (ctx->stash_slots != n_slots || ctx->stash_per != per) {
        ctx->topics_dev.release();
        CK(ctx->topics_dev.ensure(per * (size_t)n_slots * 4));
        ctx->stash_slots = n_slots;
        ctx->stash_per = per;
    }
    if (m > 0) {
        unpack_rows_kernel<<<(unsigned)cdiv(m * ctx->k, 256), 256, 0, ctx->stream>>>(
            ctx->B[ctx->curB].as<float>(), cur_scale(ctx),
            ctx->topics_dev.as<float>() + per * (size_t)slot, m, ctx->k, ctx->strideB, 1);
        ctx->launches++;
        CK(cudaGetLastError());
    }
    CK(cudaStreamSynchronize(ctx->stream));
    return PLSA_OK;
}

API int plsa_topics_device(plsa_ctx *ctx, void **device_ptr, int64_t *floats_per_slot)
{
    if (!ctx || !device_ptr) return PLSA_EINVAL;
    if (!ctx->topics_dev.p) return ctx->fail(PLSA_EINVAL, "topics_device: nothing stashed");
    *device_ptr = ctx->topics_dev.p;
    if (floats_per_slot) *floats_per_slot = (int64_t)ctx->stash_per;
    return PLSA_OK;
}

namespace {
typedef struct ncclComm *ncclComm_t;
struct NcclId { char bytes[PLSA_NCCL_ID_BYTES]; };
typedef int (*fn_GetUniqueId)(NcclId *);
typedef int (*fn_CommInitRank)(ncclComm_t *, int, NcclId, int);
typedef int (*fn_CommInitAll)(ncclComm_t *, int, const int *);
typedef int (*fn_CommDestroy)(ncclComm_t);
typedef int (*fn_CommAbort)(ncclComm_t);
typedef int (*fn_GroupStart)(void);
typedef int (*fn_GroupEnd)(void);
typedef int (*fn_Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t);
typedef int (*fn_Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t);
typedef int (*fn_AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t);
typedef const char *(*fn_GetErrorString)(int);
struct Nccl {
    void *h = nullptr;
    fn_GetUniqueId GetUniqueId;
    fn_CommInitRank CommInitRank;
    fn_CommInitAll CommInitAll;
    fn_CommDestroy CommDestroy;
    fn_CommAbort CommAbort; /* optional */
    fn_GroupStart GroupStart;
    fn_GroupEnd GroupEnd;
    fn_Send Send;
    fn_Recv Recv;
    fn_AllReduce AllReduce;
    fn_GetErrorString GetErrorString;
};
const int kNcclFloat = 7;  /* ncclFloat32 */
const int kNcclDouble = 8; /* ncclFloat64 */
const int kNcclSum = 0;    /* ncclSum */

/* Resolved once per process; the ensemble's helper thread (plsa_gather_warmup) and its worker
 * threads may ask at the same time, so the one-time work sits behind std::call_once. */
static Nccl *load_nccl()
{
    static Nccl n;
    static std::once_flag once;
    std::call_once(once, []() {
        for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
            n.h = dlopen(name, RTLD_NOW | RTLD_LOCAL);
            if (n.h) break;
        }
        if (!n.h) return;
        n.GetUniqueId = (fn_GetUniqueId)dlsym(n.h, "ncclGetUniqueId");
        n.CommInitRank = (fn_CommInitRank)dlsym(n.h, "ncclCommInitRank");
        n.CommInitAll = (fn_CommInitAll)dlsym(n.h, "ncclCommInitAll");
        n.CommDestroy = (fn_CommDestroy)dlsym(n.h, "ncclCommDestroy");
        n.CommAbort = (fn_CommAbort)dlsym(n.h, "ncclCommAbort");
        n.GroupStart = (fn_GroupStart)dlsym(n.h, "ncclGroupStart");
        n.GroupEnd = (fn_GroupEnd)dlsym(n.h, "ncclGroupEnd");
        n.Send = (fn_Send)dlsym(n.h, "ncclSend");
        n.Recv = (fn_Recv)dlsym(n.h, "ncclRecv");
        n.AllReduce = (fn_AllReduce)dlsym(n.h, "ncclAllReduce");
        n.GetErrorString = (fn_GetErrorString)dlsym(n.h, "ncclGetErrorString");
        if (!n.GetUniqueId || !n.CommInitRank || !n.CommInitAll || !n.CommDestroy ||
            !n.GroupStart || !n.GroupEnd || !n.Send || !n.Recv || !n.AllReduce) {
            dlclose(n.h);
            n.h = nullptr;
        }
    });
    return n.h ? &n : nullptr;
}

static int nccl_fail(Nccl *nc, const char *what, int r)
{
    g_err = std::string(what) + ": NCCL error: " +
            ((nc && nc->GetErrorString) ? nc->GetErrorString(r) : "?");
    return PLSA_ENCCL;
}
} // namespace

/* Communicators of the single-process gather, kept per device list: ncclCommInitAll costs a
 * few hundred milliseconds, more than a whole ensemble member; plsa_gather_warmup lets the
 * host create them in the background while the members are being fitted. */
namespace {
std::mutex g_comm_mutex;
std::map<std::vector<int>, std::vector<ncclComm_t>> g_comm_cache;

static int cached_comms(Nccl *nc, const std::vector<int> &devs, std::vector<ncclComm_t> **out)
{
    std::lock_guard<std::mutex> lock(g_comm_mutex);
    auto it = g_comm_cache.find(devs);
    if (it == g_comm_cache.end()) {
        std::vector<ncclComm_t> comms(devs.size());
        const int r = nc->CommInitAll(comms.data(), (int)devs.size(), devs.data());
        if (r != 0) return r;
        it = g_comm_cache.emplace(devs, std::move(comms)).first;
    }
    *out = &it->second;
    return 0;
}
} // namespace

API int plsa_gather_warmup(const int32_t *devices, int32_t n_devices)
{
    if (!devices || n_devices < 1) {
        g_err = "gather_warmup: bad arguments";
        return PLSA_EINVAL;
    }
    if (n_devices == 1) return PLSA_OK;
    Nccl *nc = load_nccl();
    if (!nc) {
        g_err = "gather_warmup: libnccl.so.2 could not be loaded";
        return PLSA_ENCCL;
    }
    std::vector<int> devs(devices, devices + n_devices);
    std::vector<ncclComm_t> *comms = nullptr;
    const int r = cached_comms(nc, devs, &comms);
    if (r != 0) return nccl_fail(nc, "gather_warmup: ncclCommInitAll", r);
    return PLSA_OK;
}

/* Same device: append the stashed topic matrices of `src` to those of `dst` (two contexts
 * per device keep the GPU busy while one of them prepares its next ensemble member). */
API int plsa_stash_append(plsa_ctx *dst, plsa_ctx *src, int32_t n_dst, int32_t n_src)
{
    plsa_ctx *ctx = dst;
    CHECK_CTX(ctx);
    if (!src || src->device != dst->device) return ctx->fail(PLSA_EINVAL, "stash_append: contexts must share a device");
    if (n_dst < 0 || n_src < 0 || (n_src > 0 && (!src->topics_dev.p || n_src > src->stash_slots)) ||
        (n_dst > 0 && (!dst->topics_dev.p || n_dst > dst->stash_slots)))
        return ctx->fail(PLSA_EINVAL, "stash_append: fewer stashed topic matrices than announced");
    const size_t per = n_src > 0 ? src->stash_per : dst->stash_per;
    if (n_dst > 0 && n_src > 0 && dst->stash_per != src->stash_per)
        return ctx->fail(PLSA_EINVAL, "stash_append: contexts must share k and n_terms");
    if (n_src == 0) return PLSA_OK;
    DevBuf merged;
    CK(merged.ensure(per * (size_t)(n_dst + n_src) * 4));
    CK(cudaStreamSynchronize(src->stream));
    if (n_dst > 0)
        CK(cudaMemcpyAsync(merged.p, dst->topics_dev.p, per * (size_t)n_dst * 4,
                           cudaMemcpyDeviceToDevice, dst->stream));
    CK(cudaMemcpyAsync(merged.as<float>() + per * (size_t)n_dst, src->topics_dev.p,
                       per * (size_t)n_src * 4, cudaMemcpyDeviceToDevice, dst->stream));
    CK(cudaStreamSynchronize(dst->stream));
    dst->topics_dev.release();
    dst->topics_dev = merged;
    dst->stash_slots = n_dst + n_src;
    dst->stash_per = per;
    return PLSA_OK;
}

/* Single process, one context per device (one host thread per GPU during the fits). */
API int plsa_gather_topics(plsa_ctx **ctxs, int32_t n_ctx, const int32_t *n_slots, float *out)
{
    if (!ctxs || n_ctx < 1 || !n_slots || !out) {
        g_err = "gather_topics: bad arguments";
        return PLSA_EINVAL;
    }
    plsa_ctx *root = ctxs[0];
    size_t per = 0, total = 0;
    for (int i = 0; i < n_ctx; ++i) {
        plsa_ctx *c = ctxs[i];
        if (!c || n_slots[i] < 0 || (n_slots[i] > 0 && (!c->topics_dev.p || n_slots[i] > c->stash_slots))) {
            g_err = "gather_topics: a context has fewer stashed topic matrices than requested";
            return PLSA_EINVAL;
        }
        if (n_slots[i] > 0) {
            if (per == 0) per = c->stash_per;
            if (c->stash_per != per) {
                g_err = "gather_topics: contexts must share k and n_terms";
                return PLSA_EINVAL;
            }
        }
        for (int j = 0; j < i; ++j)
            if (ctxs[j]->device == c->device) {
                g_err = "gather_topics: one context per device";
                return PLSA_EINVAL;
            }
        total += (size_t)n_slots[i];
    }
    if (total == 0) return PLSA_OK;
    cudaSetDevice(root->device);
    DevBuf &stack = root->gathered; /* stays on the root device for plsa_gathered_distances */
    root->gathered_n = root->gathered_m = 0;
    if (stack.ensure(per * total * 4) != cudaSuccess) {
        g_err = "gather_topics: staging allocation failed";
        return PLSA_ENOMEM;
    }
    int rc = PLSA_OK;
    cudaError_t ce = cudaSuccess;
    if (n_slots[0] > 0)
        ce = cudaMemcpyAsync(stack.p, root->topics_dev.p, per * (size_t)n_slots[0] * 4,
                             cudaMemcpyDeviceToDevice, root->stream);
    bool any_remote = false;
    for (int i = 1; i < n_ctx; ++i) any_remote |= n_slots[i] > 0;
    if (ce == cudaSuccess && any_remote) {
        Nccl *nc = load_nccl();
        if (!nc) {
            g_err = "gather_topics: libnccl.so.2 could not be loaded";
            return PLSA_ENCCL;
        }
        std::vector<int> devs((size_t)n_ctx);
        for (int i = 0; i < n_ctx; ++i) devs[(size_t)i] = ctxs[i]->device;
        std::vector<ncclComm_t> *cached = nullptr;
        int r = cached_comms(nc, devs, &cached);
        if (r != 0) {
            return nccl_fail(nc, "gather_topics: ncclCommInitAll", r);
        }
        std::vector<ncclComm_t> &comms = *cached;
        std::lock_guard<std::mutex> use(g_comm_mutex); /* one gather at a time per process */
        nc->GroupStart();
        size_t off = (size_t)n_slots[0] * per;
        for (int i = 1; i < n_ctx && r == 0; ++i) {
            const size_t cnt = (size_t)n_slots[i] * per;
            if (cnt == 0) continue;
            cudaSetDevice(ctxs[i]->device);
            r = nc->Send(ctxs[i]->topics_dev.p, cnt, kNcclFloat, 0, comms[(size_t)i],
                         ctxs[i]->stream);
            if (r) break;
            cudaSetDevice(root->device);
            r = nc->Recv(stack.as<float>() + off, cnt, kNcclFloat, i, comms[0], root->stream);
            off += cnt;
        }
        const int r2 = nc->GroupEnd();
        if (r == 0) r = r2;
        for (int i = 0; i < n_ctx; ++i) {
            cudaSetDevice(ctxs[i]->device);
            cudaStreamSynchronize(ctxs[i]->stream);
        }
        if (r != 0) rc = nccl_fail(nc, "gather_topics", r);
    }
    cudaSetDevice(root->device);
    if (rc == PLSA_OK && ce == cudaSuccess)
        ce = cudaMemcpyAsync(out, stack.p, per * total * 4, cudaMemcpyDeviceToHost, root->stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(root->stream);
    if (rc == PLSA_OK && ce == cudaSuccess && root->k > 0) {
        root->gathered_n = (int64_t)total * root->k;
        root->gathered_m = (int64_t)(per / (size_t)root->k);
    }
    if (rc == PLSA_OK && ce != cudaSuccess) {
        g_err = std::string("gather_topics: ") + cudaGetErrorString(ce);
        rc = PLSA_ECUDA;
    }
    if (rc) root->err = g_err;
    return rc;
}

/* One process per GPU (torchrun-style launch): rank 0 makes the id, the launcher's own
 * rendezvous distributes the 128 bytes, every rank creates its communicator. */
struct plsa_comm {
    ncclComm_t comm = nullptr;
    int device = 0, n_ranks = 1, rank = 0;
    cudaStream_t stream = nullptr;
    /* plsa_comm_abort may come from another thread (a peer rank of a sharded fit failed): the
     * handle is only read or dropped under this mutex; collectives are enqueued (not waited
     * for) under it, so an abort is never held up by a rank that is blocked on the device */
    std::mutex mu;
};

API int plsa_nccl_unique_id(char *id)
{
    if (!id) return PLSA_EINVAL;
    Nccl *nc = load_nccl();
    if (!nc) {
        g_err = "nccl_unique_id: libnccl.so.2 could not be loaded";
        return PLSA_ENCCL;
    }
    NcclId u;
    memset(&u, 0, sizeof(u));
    const int r = nc->GetUniqueId(&u);
    if (r != 0) return nccl_fail(nc, "nccl_unique_id", r);
    memcpy(id, u.bytes, PLSA_NCCL_ID_BYTES);
    return PLSA_OK;
}

API int plsa_comm_create(int device, int32_t n_ranks, int32_t rank, const char *id, plsa_comm **out)
{
    if (!out || n_ranks < 1 || rank < 0 || rank >= n_ranks || (n_ranks > 1 && !id)) {
        g_err = "comm_create: bad arguments";
        return PLSA_EINVAL;
    }
    *out = nullptr;
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) {
        g_err = std::string("comm_create: cudaSetDevice: ") + cudaGetErrorString(e);
        return PLSA_ECUDA;
    }
    plsa_comm *c = new (std::nothrow) plsa_comm();
    if (!c) return PLSA_ENOMEM;
    c->device = device;
    c->n_ranks = n_ranks;
    c->rank = rank;
    if ((e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess) {
        g_err = std::string("comm_create: ") + cudaGetErrorString(e);
        delete c;
        return PLSA_ECUDA;
    }
    if (n_ranks > 1) {
        Nccl *nc = load_nccl();
        if (!nc) {
            cudaStreamDestroy(c->stream);
            delete c;
            g_err = "comm_create: libnccl.so.2 could not be loaded";
            return PLSA_ENCCL;
        }
        NcclId u;
        memcpy(u.bytes, id, PLSA_NCCL_ID_BYTES);
        const int r = nc->CommInitRank(&c->comm, n_ranks, u, rank);
        if (r != 0) {
            cudaStreamDestroy(c->stream);
            delete c;
            return nccl_fail(nc, "comm_create: ncclCommInitRank", r);
        }
    }
    *out = c;
    return PLSA_OK;
}

/* ---- document-sharded fit ------------------------------------------------------------------------- */
/* in-place sum over the ranks of the shard communicator (float32 or float64) */
static int shard_allreduce(plsa_ctx *ctx, void *buf, size_t count, bool f64, cudaStream_t stream)
{
    plsa_comm *c = ctx->shard;
    if (!c || c->n_ranks < 2 || count == 0) return PLSA_OK;
    Nccl *nc = load_nccl();
    if (!nc) return ctx->fail(PLSA_ENCCL, "sharded fit: libnccl.so.2 could not be loaded");
    int r;
    {
        std::lock_guard<std::mutex> lock(c->mu);
        if (!c->comm) return ctx->fail(PLSA_ENCCL, "sharded fit: the communicator was aborted (a peer rank failed)");
        r = nc->AllReduce(buf, buf, count, f64 ? kNcclDouble : kNcclFloat, kNcclSum, c->comm, stream);
    }
    ctx->launches++;
    if (r != 0) {
        const int rc = nccl_fail(nc, "sharded fit: ncclAllReduce", r);
        ctx->err = g_err;
        return rc;
    }
    return PLSA_OK;
}

static int shard_n_ranks(const plsa_ctx *ctx) { return ctx->shard ? ctx->shard->n_ranks : 1; }
static int shard_rank(const plsa_ctx *ctx) { return ctx->shard ? ctx->shard->rank : 0; }

static void p2p_detach_peers(plsa_ctx *ctx)
{
    for (int p = 0; p < SHARD_MAX_RANKS; ++p) {
        if (ctx->p2p.peer_base[p] && ctx->p2p.peer_ipc[p]) cudaIpcCloseMemHandle(ctx->p2p.peer_base[p]);
        ctx->p2p.peer_base[p] = nullptr;
        ctx->p2p.peer_ipc[p] = false;
    }
    ctx->p2p.n_attached = 0;
}

static void p2p_release(plsa_ctx *ctx)
{
    p2p_detach_peers(ctx);
    ctx->p2p.block.release();
    ctx->p2p.err.release();
    ctx->p2p.part_bytes = 0;
    ctx->p2p.seq = 0;
}

API int plsa_set_shard(plsa_ctx *ctx, plsa_comm *comm)
{
    CHECK_CTX(ctx);
    if (comm && comm->device != ctx->device)
        return ctx->fail(PLSA_EINVAL, "set_shard: context and communicator devices differ");
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    p2p_release(ctx);
    ctx->shard = comm;
    return PLSA_OK;
}

API int plsa_shard_p2p_prepare(plsa_ctx *ctx, uint64_t *base, int64_t *bytes)
{
    CHECK_CTX(ctx);
    if (!ctx->shard) return ctx->fail(PLSA_EINVAL, "shard_p2p_prepare: no shard communicator attached");
    if (!ctx->have_factors) return ctx->fail(PLSA_EINVAL, "shard_p2p_prepare: set the factors first");
    if (ctx->shard->n_ranks > SHARD_MAX_RANKS)
        return ctx->fail(PLSA_EINVAL, "shard_p2p_prepare: too many ranks for the peer-memory path");
    p2p_release(ctx);
    /* [partial 0 | partial 1 | finished slice 0 | finished slice 1 | signal words (2 x 64 B)] */
    const size_t part = (size_t)std::max<int64_t>(ctx->cur().m, 1) * ctx->strideB * 4;
    const size_t red = (size_t)cdiv(std::max<int64_t>(ctx->cur().m, 1), ctx->shard->n_ranks) * ctx->strideB * 4;
    const size_t total = 2 * part + 2 * red + 256;
    ctx->p2p.red_bytes = red;
    CK(ctx->p2p.block.ensure(total));
    CK(ctx->p2p.err.ensure(4));
    CK(cudaMemsetAsync(ctx->p2p.block.p, 0, total, ctx->stream));
    CK(cudaMemsetAsync(ctx->p2p.err.p, 0, 4, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->p2p.part_bytes = part;
    if (base) *base = (uint64_t)(uintptr_t)ctx->p2p.block.p;
    if (bytes) *bytes = (int64_t)total;
    return PLSA_OK;
}

/* Unmap the peers' exchange blocks (this rank's own block stays).  Between processes an
 * exported block must outlive its importers' mappings: every rank detaches, the caller's
 * barrier follows, and only then plsa_set_shard(ctx, NULL) / plsa_ctx_destroy free the blocks. */
API int plsa_shard_p2p_detach(plsa_ctx *ctx)
{
    CHECK_CTX(ctx);
    if (ctx->stream) CK(cudaStreamSynchronize(ctx->stream));
    p2p_detach_peers(ctx);
    return PLSA_OK;
}

API int plsa_shard_p2p_export(plsa_ctx *ctx, char *handle)
{
    CHECK_CTX(ctx);
    if (!handle || !ctx->p2p.block.p) return ctx->fail(PLSA_EINVAL, "shard_p2p_export: nothing prepared");
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, ctx->p2p.block.p));
    static_assert(sizeof(h) == PLSA_IPC_HANDLE_BYTES, "IPC handle size");
    memcpy(handle, &h, sizeof(h));
    return PLSA_OK;
}

API int plsa_shard_p2p_attach(plsa_ctx *ctx, int32_t peer_rank, int32_t peer_device, uint64_t base,
                              const char *ipc_handle)
{
    CHECK_CTX(ctx);
    if (!ctx->shard || !ctx->p2p.block.p)
        return ctx->fail(PLSA_EINVAL, "shard_p2p_attach: prepare first");
    if (peer_rank < 0 || peer_rank >= ctx->shard->n_ranks || peer_rank == ctx->shard->rank)
        return ctx->fail(PLSA_EINVAL, "shard_p2p_attach: bad peer rank");
    if (ctx->p2p.peer_base[peer_rank]) return ctx->fail(PLSA_EINVAL, "shard_p2p_attach: peer attached twice");
    void *ptr = nullptr;
    if (ipc_handle) { /* another process: map its block (peer access is enabled lazily) */
        cudaIpcMemHandle_t h;
        memcpy(&h, ipc_handle, sizeof(h));
        CK(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        ctx->p2p.peer_ipc[peer_rank] = true;
    } else {          /* same process: the peer's device pointer, after enabling peer access */
        int can = 0;
        CK(cudaDeviceCanAccessPeer(&can, ctx->device, peer_device));
        if (!can) return ctx->fail(PLSA_ECUDA, "shard_p2p_attach: no peer access between the devices");
        /* once per (device, peer) and process: asking again is an error return, harmless but
         * noisy under compute-sanitizer */
        static std::mutex mu;
        static std::set<std::pair<int, int>> enabled;
        std::lock_guard<std::mutex> lock(mu);
        if (!enabled.count({ctx->device, peer_device})) {
            cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError(); /* NCCL was first */
            else if (e != cudaSuccess)
                return ctx->fail(PLSA_ECUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
            enabled.insert({ctx->device, peer_device});
        }
        ptr = (void *)(uintptr_t)base;
    }
    if (!ptr) return ctx->fail(PLSA_EINVAL, "shard_p2p_attach: null peer block");
    ctx->p2p.peer_base[peer_rank] = ptr;
    ctx->p2p.n_attached++;
    return PLSA_OK;
}

/* Tear the communicator down without waiting for outstanding collectives (ncclCommAbort):
 * called for every rank of a sharded fit when one of them failed, so that the others do not
 * wait for it forever.  The handle stays valid for plsa_comm_destroy. */
API int plsa_comm_abort(plsa_comm *c)
{
    if (!c) return PLSA_OK;
    cudaSetDevice(c->device);
    std::lock_guard<std::mutex> lock(c->mu);
    if (c->comm) {
        Nccl *nc = load_nccl();
        if (nc && nc->CommAbort) nc->CommAbort(c->comm);
        else if (nc) nc->CommDestroy(c->comm);
        c->comm = nullptr;
    }
    return PLSA_OK;
}

API int plsa_comm_destroy(plsa_comm *c)
{
    if (!c) return PLSA_OK;
    cudaSetDevice(c->device);
    if (c->comm) {
        Nccl *nc = load_nccl();
        if (nc) nc->CommDestroy(c->comm);
    }
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return PLSA_OK;
}

/* Every rank sends the first n_local slots of ctx's stash; `root` receives them in rank
 * order into `out` (host, [sum(n_per_rank) * k, m]); n_per_rank is read on every rank. */
API int plsa_comm_gather_topics(plsa_comm *c, plsa_ctx *ctx, const int32_t *n_per_rank,
                                int32_t root, float *out)
{
    if (!c || !n_per_rank || root < 0 || root >= c->n_ranks) {
        g_err = "comm_gather_topics: bad arguments";
        return PLSA_EINVAL;
    }
    CHECK_CTX(ctx);
    if (ctx->device != c->device)
        return ctx->fail(PLSA_EINVAL, "comm_gather_topics: context and communicator devices differ");
    const int32_t mine = n_per_rank[c->rank];
    if (mine < 0 || (mine > 0 && (!ctx->topics_dev.p || mine > ctx->stash_slots)))
        return ctx->fail(PLSA_EINVAL, "comm_gather_topics: fewer stashed topic matrices than announced");
    if (!ctx->have_factors) return ctx->fail(PLSA_EINVAL, "comm_gather_topics: no model");
    const size_t per = (size_t)std::max<int64_t>(ctx->cur().m, 1) * ctx->k;
    size_t total = 0;
    for (int r = 0; r < c->n_ranks; ++r) total += (size_t)n_per_rank[r];
    const bool is_root = c->rank == root;
    if (is_root && !out && total > 0) return ctx->fail(PLSA_EINVAL, "comm_gather_topics: null output");
    DevBuf &stack = ctx->gathered; /* stays on the root's device for plsa_gathered_distances */
    ctx->gathered_n = ctx->gathered_m = 0;
    if (is_root && total > 0) CK(stack.ensure(per * total * 4));
    Nccl *nc = c->n_ranks > 1 ? load_nccl() : nullptr;
    int r = 0;
    if (c->n_ranks > 1) {
        nc->GroupStart();
        if (is_root) {
            size_t off = 0;
            for (int p = 0; p < c->n_ranks && r == 0; ++p) {
                const size_t cnt = (size_t)n_per_rank[p] * per;
                if (p != root && cnt > 0)
                    r = nc->Recv(stack.as<float>() + off, cnt, kNcclFloat, p, c->comm, c->stream);
                off += cnt;
            }
        } else if (mine > 0) {
            r = nc->Send(ctx->topics_dev.p, (size_t)mine * per, kNcclFloat, root, c->comm, c->stream);
        }
        const int r2 = nc->GroupEnd();
        if (r == 0) r = r2;
    }
    cudaError_t ce = cudaSuccess;
    if (is_root && total > 0) {
        size_t off = 0;
        for (int p = 0; p < root; ++p) off += (size_t)n_per_rank[p] * per;
        if (mine > 0)
            ce = cudaMemcpyAsync(stack.as<float>() + off, ctx->topics_dev.p, (size_t)mine * per * 4,
                                 cudaMemcpyDeviceToDevice, c->stream);
        if (ce == cudaSuccess && r == 0)
            ce = cudaMemcpyAsync(out, stack.p, per * total * 4, cudaMemcpyDeviceToHost, c->stream);
    }
    cudaError_t ce2 = cudaStreamSynchronize(c->stream);
    if (ce == cudaSuccess) ce = ce2;
    if (is_root && total > 0 && r == 0 && ce == cudaSuccess) {
        ctx->gathered_n = (int64_t)total * ctx->k;
        ctx->gathered_m = (int64_t)(per / (size_t)ctx->k);
    }
    if (r != 0) {
        const int rc = nccl_fail(nc, "comm_gather_topics", r);
        ctx->err = g_err;
        return rc;
    }
    if (ce != cudaSuccess)
        return ctx->fail(PLSA_ECUDA, std::string("comm_gather_topics: ") + cudaGetErrorString(ce));
    return PLSA_OK;
}
