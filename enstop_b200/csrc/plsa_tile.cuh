/*
 * plsa_tile.cuh — the "tiled" row pass: the hot rows of the gathered factor live in shared
 * memory, staged by the TMA unit, and are consumed LANE-PER-ENTRY.
 *
 * Why (profiles/r2_microbench_stage.txt, profiles/r2_kernel_experiments.md): a k-wide factor
 * row per stored entry costs 1.75 cycles/entry/SM through the texture path and every one of
 * them crosses the L2->SM fabric; with a Zipf vocabulary two thirds of the doc pass's gathers
 * hit the ~2.5 k most frequent terms, whose rows (80 B each at k = 20) fit one CTA's shared
 * memory.  Those entries are split off into their own CSR ("head"), whose column index is the
 * tile slot; the rest ("tail") stays with the group-per-row texture kernel of plsa_kernels.cuh.
 *
 * Head kernel (tile_pass_kernel): one persistent CTA per SM copies the tile with cp.async.bulk
 * (TMA, mbarrier complete_tx) and then walks work items OCTET-PER-ITEM: 8 lanes share one row
 * of X, each lane owns one stored entry per step, reads the whole factor row of its entry with
 * KC LDS.128 and keeps the posterior normaliser in-lane — no shuffle per entry; the k
 * accumulators are folded over the octet once per item.  An LDS.128 is served in quarter-warp
 * phases of 8 lanes x 16 B; rows sit at an odd pitch (in 16-byte chunks), so the 8 lanes of a
 * phase are conflict free exactly when their slots are distinct mod 8.  tile_place_kernel
 * therefore lays every head row out in steps of 8 slots, position p of a step holding an entry
 * whose slot is = p (mod 8) or a zero-valued padding entry (slot p): conflict free by
 * construction (measured: an octet with a two-way conflict costs 3x a conflict-free one).
 *
 * E-step threshold (enstop/plsa.py:98-102) without compare/select: the owned row is multiplied
 * by S = FLT_MIN / thresh, so a product at or below the threshold is subnormal and the
 * flush-to-zero multiply drops it; posterior and M-step sums do not see the common factor.
 * This needs gathered values <= 1, which is why the tiled mode keeps P(w|z) normalised in
 * memory (normalise_rows_kernel) instead of folding the column scale into the owned row.
 */
#pragma once
#include "plsa_kernels.cuh"

namespace plsa {

#ifndef PLSA_TILE_ABLATE
#define PLSA_TILE_ABLATE 0                        /* 1-3: timing experiments, results are wrong */
#endif
#ifndef PLSA_TILE_PF_L1
#define PLSA_TILE_PF_L1 0
#endif
#ifndef PLSA_TILE_THREADS
#define PLSA_TILE_THREADS 512
#endif
constexpr int TILE_THREADS = PLSA_TILE_THREADS;   /* 16 warps, one CTA per SM */
constexpr int TILE_MIN_ROWS = 8;                  /* padding entries address slots 0..7 */
constexpr int TILE_MAX_SMEM = 221 * 1024;         /* dynamic shared memory a tile may ask for (option "tile_kb" <= 220) */

__host__ __device__ constexpr int tile_pitch_chunks(int kc) { return kc | 1; } /* odd */

struct TileArgs {
    const int4 *items;        /* [n_items] work items in launch order {first entry, entries (a
                                 multiple of 8), factor row the item owns, its partial-sum
                                 slot}: doc side by padded length, longest first; term side by
                                 tile block, then by length (tile_headers_kernel)            */
    const int2 *ent;          /* entries {slot inside the tile, value bits}                  */
    const float *own_old;     /* [rows, stride_own]                                          */
    const float *tile_src;    /* compact image of the gathered factor, [*, pitch]; block b of
                                 the tile starts at row b * block_rows                        */
    float *partial_out;       /* [n_items, kp] raw M-step sums of each item                  */
    const int32_t *cta_begin; /* [grid + 1] ranges into `order`, one per CTA (equal work);
                                 nullptr: batches of four items dealt round-robin to all warps */
    const int32_t *block_begin; /* [n_blocks + 1] ranges into `order`; nullptr: one block     */
    const float *row_weight;  /* LL: sample_weight[d]                                        */
    double *cta_partial;      /* LL: [grid]                                                  */
    unsigned int *ticket;     /* LL: zeroed counter                                          */
    double *ll_out;           /* LL: log-likelihood of the tiled entries                     */
    int *flag;                /* LL: raised when a normaliser is too small to trust          */
    int64_t n_items;
    int64_t src_rows;         /* rows of the gathered factor (last block may be short)       */
    int32_t block_rows, n_blocks, stride_own, kp;
    float ftz_scale;          /* S = FLT_MIN / thresh                                        */
    float inv_ftz_scale;      /* LL: float(1 / S): the log is taken of the unscaled normaliser */
    double log2_ftz_corr;     /* LL: log2(S * float(1 / S)), the systematic part of that
                                 un-scaling, taken off once per item in double                */
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "TILE_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra TILE_DONE;\n"
        "bra TILE_WAIT;\n"
        "TILE_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
/* TMA bulk copy global -> shared, completion counted in bytes on the mbarrier (SASS UBLKCP) */
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

/* Sum K values over the 8 lanes of an octet by halving: at every level a lane keeps one half of
 * its values and receives its partner's copy of that half.  After the three levels lane li
 * holds the octet totals of `n` consecutive topics starting at `first` (n <= ceil(K/8)). */
template <int N, int K>
__device__ __forceinline__ void octet_fold_level(float (&a)[K], int li, int bit, int &first, int &n)
{
    constexpr int half = (N + 1) / 2;
    const bool up = (li & bit) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
        const float lo = a[i];
        const float hi = (i + half < N) ? a[i + half < K ? i + half : 0] : 0.f;
        const float send = up ? lo : hi;
        const float keep = up ? hi : lo;
        a[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
    }
    first += up ? half : 0;
    n = up ? max(0, n - half) : min(n, half);
}

template <int K>
__device__ __forceinline__ int octet_fold(float (&a)[K], int li, int &first)
{
    constexpr int N0 = K, N1 = (N0 + 1) / 2, N2 = (N1 + 1) / 2;
    int n = K;
    first = 0;
    octet_fold_level<N0, K>(a, li, 4, first, n);
    octet_fold_level<N1, K>(a, li, 2, first, n);
    octet_fold_level<N2, K>(a, li, 1, first, n);
    return n;
}

/* four items (one per octet of the warp) starting at item `first`, items at or past `limit`
 * are idle */
template <int KC, bool LL>
__device__ __forceinline__ void tile_batch(const TileArgs &a, const float4 *tile, const int4 hdr,
                                           bool has, const int4 next, bool has_next, int li,
                                           double &ll_acc, float &min_norm)
{
    constexpr int PC = tile_pitch_chunks(KC);
    constexpr int K = 4 * KC;
    const int start = hdr.x, len = has ? hdr.y : 0, row = hdr.z, item = hdr.w;
    /* Items are short (a dozen steps): what a batch reads first — its owned row and its first
     * entry blocks — is pulled into L2 while the batch before it runs; inside an item the
     * stream is prefetched four blocks ahead (below). */
    if (has_next) {
        const int2 *ne = a.ent + next.x + li * 4;
        const int nmem = (next.y + 31) & ~31;
#if PLSA_TILE_PF_L1   /* experiment: the first two blocks and the owned row into L1 (needs a smaller tile) */
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (q < 2 && q * 32 < nmem) asm volatile("prefetch.global.L1 [%0];" ::"l"(ne + q * 32));
            else if (q * 32 < nmem) asm volatile("prefetch.global.L2 [%0];" ::"l"(ne + q * 32));
        }
        if (li == 0) asm volatile("prefetch.global.L1 [%0];" ::"l"(a.own_old + (int64_t)next.z * a.stride_own));
#else
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (q * 32 < nmem) asm volatile("prefetch.global.L2 [%0];" ::"l"(ne + q * 32));
        if (li == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.own_old + (int64_t)next.z * a.stride_own));
#endif
    }
    int maxlen = len;
    maxlen = max(maxlen, __shfl_xor_sync(0xffffffffu, maxlen, 8));
    maxlen = max(maxlen, __shfl_xor_sync(0xffffffffu, maxlen, 16));
    if (maxlen == 0) { /* items without entries still own a (zero) partial */
        if (has)
            for (int z = li; z < a.kp; z += 8) a.partial_out[(int64_t)item * a.kp + z] = 0.f;
        return;
    }
    f32x2 own2[2 * KC], acc2[2 * KC];
#pragma unroll
    for (int c = 0; c < KC; ++c) {
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        if (has) o = ldg_f4(a.own_old + (int64_t)row * a.stride_own + 4 * c);
        own2[2 * c] = pk2(o.x * a.ftz_scale, o.y * a.ftz_scale);
        own2[2 * c + 1] = pk2(o.z * a.ftz_scale, o.w * a.ftz_scale);
        acc2[2 * c] = 0ull;
        acc2[2 * c + 1] = 0ull;
    }
    double ll_row = 0.0, x_sum = 0.0; /* LL: sum of x * log2(normaliser), sum of x */
    /* Entries come in blocks of 4 steps, lane-major (tile_place_kernel): lane li reads its four
     * entries of a block with ONE 32-byte load.  The block after the current one is already in
     * registers and the lines three blocks further are on their way into L2, so that the entry
     * stream (8 bytes per lane and step, from HBM) never stalls a step — with a look-ahead of a
     * single step every step waited a full memory latency (profiles/r2_kernel_experiments.md). */
    const int2 *ent = a.ent + start + li * 4;
    const int len_mem = (len + 31) & ~31;
    int2 e[4], en[4];
#if PLSA_TILE_ABLATE == 3   /* timing experiment: no entry stream */
#pragma unroll
    for (int q = 0; q < 4; ++q) e[q] = make_int2(li + 8 * q, 0x3f800000);
#else
    load_entries<4, true>(ent, e); /* readable for every item: the array is padded by a block */
#endif
    if (len == 0) {
#pragma unroll
        for (int q = 0; q < 4; ++q) e[q] = make_int2(li, 0);
    }
    for (int t = 0; t < maxlen; t += 32) {
        if (t + 32 < len_mem) {
#if PLSA_TILE_ABLATE == 3
#pragma unroll
            for (int q = 0; q < 4; ++q) en[q] = make_int2(li + 8 * q + (t & 1023), 0x3f800000);
#else
            load_entries<4, true>(ent + t + 32, en);
#endif
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) en[q] = make_int2(li, 0);
        }
        if (t + 128 < len_mem) asm volatile("prefetch.global.L2 [%0];" ::"l"(ent + t + 128));
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            /* past the item's end (the block is padded, or another octet's item is longer): own
             * residue class, value 0 */
            const bool on = t + 8 * q < len;
            const int slot = on ? e[q].x : li;
            const float x = on ? __int_as_float(e[q].y) : 0.f;
            float4 g[KC];
#if PLSA_TILE_ABLATE == 1   /* timing experiment: no shared-memory reads */
#pragma unroll
            for (int c = 0; c < KC; ++c) g[c] = make_float4(__int_as_float(slot + c), 1.f, 2.f, 3.f);
#else
#pragma unroll
            for (int c = 0; c < KC; ++c) g[c] = tile[slot * PC + c];
#endif
            f32x2 v[2 * KC];
#pragma unroll
            for (int c = 0; c < KC; ++c) {
                v[2 * c] = mul2_ftz(pk2(g[c].x, g[c].y), own2[2 * c]);
                v[2 * c + 1] = mul2_ftz(pk2(g[c].z, g[c].w), own2[2 * c + 1]);
            }
            f32x2 s = add2(v[0], v[1]);
#pragma unroll
            for (int c = 1; c < KC; ++c) s = add2(s, add2(v[2 * c], v[2 * c + 1]));
            float s_lo, s_hi;
            upk2(s, s_lo, s_hi);
            const float norm = s_lo + s_hi;
            if constexpr (LL) {
                /* plsa.py:383-384; x == 0 marks padding.  The log is taken of the same number the
                 * group-per-row pass sees (MUFU.LG2's error depends on the magnitude): the sum
                 * carries S, float(1/S) takes it off, and what S * float(1/S) differs from 1 by is
                 * corrected per item in double */
                ll_row += (double)((x != 0.f) ? x * log2_ftz(norm * a.inv_ftz_scale) : 0.f);
                x_sum += (double)x;
                min_norm = fminf(min_norm, (x != 0.f) ? norm : 3.0e38f);
            }
#if PLSA_TILE_ABLATE == 2   /* timing experiment: no reciprocal, no accumulation */
            acc2[0] = add2(acc2[0], pk2(norm, x));
#else
            /* x / norm (see pass_block).  The scaled normaliser can be as small as FLT_MIN, so
             * the quotient is formed 2^-24 too small (no overflow for counts below 6e7; larger
             * ones are clamped) and the item's sums are multiplied by 2^24 at the end: exact */
            const float cf = fminf((x * 5.9604645e-8f) * rcp_fast(norm), 3.0e38f);
            const f32x2 c2 = pk2(cf, cf);
#pragma unroll
            for (int c = 0; c < 2 * KC; ++c) acc2[c] = fma2(c2, v[c], acc2[c]);
#endif
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) e[q] = en[q];
    }
    if constexpr (LL) {
        /* same float ln 2 as __logf / the group-per-row pass, so that the paths agree */
        if (has)
            ll_acc += (ll_row - a.log2_ftz_corr * x_sum) * (double)0.69314718f *
                      (double)a.row_weight[row];
    }
    float acc[K];
#pragma unroll
    for (int q = 0; q < 2 * KC; ++q) {
        upk2(acc2[q], acc[2 * q], acc[2 * q + 1]);
        acc[2 * q] *= 16777216.f;
        acc[2 * q + 1] *= 16777216.f;
    }
    int first_topic;
    const int nv = octet_fold<K>(acc, li, first_topic);
    if (has) {
        float *dst = a.partial_out + (int64_t)item * a.kp + first_topic;
#pragma unroll
        for (int q = 0; q < (K + 7) / 8; ++q)
            if (q < nv && first_topic + q < a.kp) dst[q] = acc[q];
    }
}

template <int KC, bool LL>
__global__ void __launch_bounds__(TILE_THREADS, 1) tile_pass_kernel(const TileArgs a)
{
    constexpr int PC = tile_pitch_chunks(KC);
    constexpr int NW = TILE_THREADS / 32;
    extern __shared__ __align__(128) unsigned char tile_smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ double ll_sm[NW];
    __shared__ bool ll_last;
    const float4 *tile = reinterpret_cast<const float4 *>(tile_smem);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, li = lane & 7, oct = lane >> 3;

    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t phase = 0;
    /* stage block `b` of the gathered factor: one elected thread arms the barrier with the byte
     * count and issues the TMA bulk copies; everybody waits on the barrier's phase */
    auto load_tile = [&](int b) {
        if (threadIdx.x == 0) {
            const int64_t r0 = (int64_t)b * a.block_rows;
            const int rows = (int)min((int64_t)a.block_rows, a.src_rows - r0);
            const uint32_t bytes = (uint32_t)max(rows, TILE_MIN_ROWS) * PC * 16u;
            /* shared memory the generic proxy has read is about to be written by the async proxy */
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(&bar, bytes);
            const char *src = reinterpret_cast<const char *>(a.tile_src) + r0 * (PC * 16);
            for (uint32_t off = 0; off < bytes; off += 32768u)
                bulk_g2s(tile_smem + off, src + off, min(bytes - off, 32768u), &bar);
        }
        mbar_wait(&bar, phase);
        phase ^= 1u;
    };

    double ll_acc = 0.0;
    float min_norm = 3.0e38f;
    if (a.cta_begin == nullptr) {
        /* one block for everybody; batches dealt round-robin over all warps of the grid: with
         * the items sorted by length every warp gets the same work to within its last batch */
        load_tile(0);
        /* consecutive batches (equally long: the items are sorted) go to DIFFERENT CTAs — warp w
         * of CTA c takes batches w * grid + c, + grid * NW, ...: every SM gets the same mix */
        const int64_t gw = (int64_t)warp * gridDim.x + blockIdx.x, GW = (int64_t)gridDim.x * NW;
        /* item headers are fetched two batches ahead, so that a batch can prefetch what the
         * batch after it reads first */
        auto header = [&](int64_t i) {
            return i < a.n_items ? __ldg(a.items + i) : make_int4(0, 0, 0, 0);
        };
        int4 hdr = header(gw * 4 + oct), next = header(gw * 4 + GW * 4 + oct);
        for (int64_t b = gw * 4; b < a.n_items; b += GW * 4) {
            const int4 after = header(b + 2 * GW * 4 + oct);
            tile_batch<KC, LL>(a, tile, hdr, b + oct < a.n_items, next, b + GW * 4 + oct < a.n_items, li,
                               ll_acc, min_norm);
            hdr = next;
            next = after;
        }
    } else {
        /* this CTA's range of items, cut at tile-block boundaries; a new block = a new tile */
        const int lo = a.cta_begin[blockIdx.x], hi = a.cta_begin[blockIdx.x + 1];
        int blk = 0;
        if (a.block_begin) /* first block whose range ends behind lo */
            while (blk + 1 < a.n_blocks && a.block_begin[blk + 1] <= lo) ++blk;
        bool first_tile = true;
        for (int cur = lo; cur < hi;) {
            const int seg_end = a.block_begin ? min(hi, a.block_begin[blk + 1]) : hi;
            if (seg_end > cur) {
                if (!first_tile) __syncthreads(); /* everybody is done with the previous tile */
                first_tile = false;
                load_tile(blk);
                auto header = [&](int64_t i) {
                    return i < seg_end ? __ldg(a.items + i) : make_int4(0, 0, 0, 0);
                };
                int4 hdr = header(cur + warp * 4 + oct), next = header(cur + warp * 4 + NW * 4 + oct);
                for (int64_t b = cur + warp * 4; b < seg_end; b += NW * 4) {
                    const int4 after = header(b + 2 * NW * 4 + oct);
                    tile_batch<KC, LL>(a, tile, hdr, b + oct < seg_end, next, b + NW * 4 + oct < seg_end, li,
                                       ll_acc, min_norm);
                    hdr = next;
                    next = after;
                }
                cur = seg_end;
            }
            ++blk;
        }
    }
    if constexpr (LL) {
        if (min_norm < PLSA_FUSED_LL_MIN_NORM * a.ftz_scale) *a.flag = 1;
        /* deterministic: lanes (butterfly), warps in order, CTAs in index order by the last
         * CTA to arrive */
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) ll_acc += __shfl_xor_sync(0xffffffffu, ll_acc, off);
        if (lane == 0) ll_sm[warp] = ll_acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int w = 0; w < NW; ++w) t += ll_sm[w];
            a.cta_partial[blockIdx.x] = t;
            __threadfence();
            ll_last = atomicAdd(a.ticket, 1u) == gridDim.x - 1;
        }
        __syncthreads();
        if (ll_last && threadIdx.x == 0) {
            __threadfence();
            double t = 0.0;
            for (unsigned c = 0; c < gridDim.x; ++c) t += a.cta_partial[c];
            *a.ll_out = t;
            *a.ticket = 0u;
        }
    }
}

/* ---- corpus preparation for the tiled pass -------------------------------------------------- */
/* how often every column occurs (the head of the vocabulary = the most frequent columns) */
__global__ void col_count_kernel(const int2 *__restrict__ ent, int64_t nnz, int32_t *__restrict__ count)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nnz) atomicAdd(count + ent[i].x, 1);
}

/* slot_of[col] = s for the tile_rows most frequent columns (sorted_cols[0..tile_rows)) */
__global__ void tile_slot_kernel(const int32_t *__restrict__ sorted_cols, int32_t tile_rows,
                                 int32_t *__restrict__ slot_of)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < tile_rows) slot_of[sorted_cols[s]] = s;
}

/* Slot of an entry inside its item's tile.  Doc side: the tile holds the most frequent terms,
 * slot_of[term] (< 0: not in the tile, the entry goes to the tail).  Term side: an item is one
 * (term, block of documents) pair and the tile holds that block's rows of P(z|d): slot =
 * document - first document of the block. */
struct SlotMap {
    const int32_t *slot_of; /* doc side */
    int32_t n_blocks, block_rows; /* term side (slot_of == nullptr): item v covers block v % n_blocks */
    __device__ __forceinline__ int operator()(int64_t item, int col) const
    {
        if (slot_of) return slot_of[col];
        return col - (int)(item % n_blocks) * block_rows;
    }
};

/* per item: padded length of its tiled part (8 x the largest residue class) and its tail
 * length; one warp per item; item r covers entries [beg[r], end[r]) */
__global__ void tile_count_kernel(const int32_t *__restrict__ beg, const int32_t *__restrict__ end,
                                  int64_t n_items, const int2 *__restrict__ ent, const SlotMap map,
                                  int32_t *__restrict__ head_len, int32_t *__restrict__ head_mem,
                                  int32_t *__restrict__ tail_len)
{
    const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= n_items) return;
    const int lane = threadIdx.x & 31;
    const int32_t p0 = beg[r], p1 = end[r];
    int cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int head = 0;
    for (int32_t p = p0; p < p1; p += 32) {
        const bool ok = p + lane < p1;
        const int s = ok ? map(r, ent[p + lane].x) : -1;
#pragma unroll
        for (int q = 0; q < 8; ++q) cnt[q] += __popc(__ballot_sync(0xffffffffu, s >= 0 && (s & 7) == q));
        head += __popc(__ballot_sync(0xffffffffu, s >= 0));
    }
    int mx = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) mx = max(mx, cnt[q]);
    if (lane == 0) {
        head_len[r] = 8 * mx;                 /* slots the kernel walks */
        head_mem[r] = (8 * mx + 31) & ~31;    /* slots in memory: whole blocks of 4 steps */
        if (tail_len) tail_len[r] = (p1 - p0) - head;
    }
}

/* write the tiled rows (steps of 8 slots, position p of a step = an entry with slot = p mod 8 or
 * padding {p, 0}) and the tail rows (original order); one warp per item.  `weight` (term side,
 * sample weights): values are multiplied by weight[column] (plsa.py:293-297). */
__global__ void tile_place_kernel(const int32_t *__restrict__ beg, const int32_t *__restrict__ end,
                                  int64_t n_items, const int2 *__restrict__ ent, const SlotMap map,
                                  const int32_t *__restrict__ head_indptr,
                                  const int32_t *__restrict__ tail_indptr, int2 *__restrict__ head_ent,
                                  int2 *__restrict__ tail_ent, const float *__restrict__ weight)
{
    const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= n_items) return;
    const int lane = threadIdx.x & 31;
    const int32_t p0 = beg[r], p1 = end[r];
    const int32_t h0 = head_indptr[r], h1 = head_indptr[r + 1], t0 = tail_indptr ? tail_indptr[r] : 0;
    /* memory order inside a block of 4 steps (32 slots): lane-major, [lane p][step q] at
     * 4 p + q — a lane's four entries are one 32-byte load */
    for (int32_t h = h0 + lane; h < h1; h += 32) head_ent[h] = make_int2(((h - h0) & 31) >> 2, 0);
    __syncwarp();
    int base[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int tbase = 0;
    const unsigned lt = (1u << lane) - 1u;
    for (int32_t p = p0; p < p1; p += 32) {
        const bool ok = p + lane < p1;
        int2 e = make_int2(0, 0);
        int s = -1;
        if (ok) {
            e = ent[p + lane];
            s = map(r, e.x);
            if (weight) e.y = __float_as_int(__int_as_float(e.y) * weight[e.x]);
        }
        int rank = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const unsigned m = __ballot_sync(0xffffffffu, s >= 0 && (s & 7) == q);
            if (s >= 0 && (s & 7) == q) rank = base[q] + __popc(m & lt);
            base[q] += __popc(m);
        }
        const unsigned mt = __ballot_sync(0xffffffffu, ok && s < 0);
        if (ok) {
            if (s >= 0) head_ent[h0 + (rank >> 2) * 32 + (s & 7) * 4 + (rank & 3)] = make_int2(s, e.y);
            else if (tail_ent) tail_ent[t0 + tbase + __popc(mt & lt)] = e;
        }
        tbase += __popc(mt);
    }
}

/* ---- term side: which terms are tiled, and their (term, document block) items ---------------- */
/* a term is tiled when it has at least `min_count` entries (an average item then has enough
 * entries to amortise its owned row and its partial sum) */
__global__ void term_tiled_flag_kernel(const int32_t *__restrict__ tindptr, int64_t m, int32_t min_count,
                                       int32_t *__restrict__ flag)
{
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w > m) return;
    flag[w] = (w < m && tindptr[w + 1] - tindptr[w] >= min_count) ? 1 : 0;
}

/* item v = (tiled term ti, block b): its range of term-major entries (documents ascending inside
 * a term: binary search for the block's first document), the term it owns, its sort key
 * (block-major, inside a block longest first) */
__global__ void term_items_kernel(const int32_t *__restrict__ tindptr, int64_t m,
                                  const int32_t *__restrict__ flag, const int32_t *__restrict__ tiled_at,
                                  const int2 *__restrict__ t_ent, int32_t n_blocks, int32_t block_rows,
                                  int32_t *__restrict__ beg, int32_t *__restrict__ end,
                                  int32_t *__restrict__ own_row, int32_t *__restrict__ tiled_terms)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m * n_blocks) return;
    const int64_t w = i / n_blocks;
    const int b = (int)(i - w * n_blocks);
    if (!flag[w]) return;
    const int32_t ti = tiled_at[w];
    auto lower = [&](int32_t doc) { /* first entry of term w with document >= doc */
        int32_t lo = tindptr[w], hi = tindptr[w + 1];
        while (lo < hi) {
            const int32_t mid = (lo + hi) >> 1;
            if (t_ent[mid].x < doc) lo = mid + 1; else hi = mid;
        }
        return lo;
    };
    const int64_t v = (int64_t)ti * n_blocks + b;
    beg[v] = lower(b * block_rows);
    end[v] = (b + 1 == n_blocks) ? tindptr[w + 1] : lower((b + 1) * block_rows);
    own_row[v] = (int32_t)w;
    if (b == 0) tiled_terms[ti] = (int32_t)w;
}

/* sort key of an item: block-major, inside a block longest first (len is a multiple of 8,
 * at most block_rows) */
__global__ void term_item_keys_kernel(const int32_t *__restrict__ head_len, int64_t n_items,
                                      int32_t n_blocks, int32_t block_rows, int32_t *__restrict__ keys)
{
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_items) return;
    const int per = block_rows / 8 + 1;
    keys[v] = (int32_t)(v % n_blocks) * per + (block_rows / 8 - head_len[v] / 8);
}

/* the items in launch order as self-contained headers {first entry, entries, owned row, slot} */
__global__ void tile_headers_kernel(const int32_t *__restrict__ order, const int32_t *__restrict__ indptr,
                                    const int32_t *__restrict__ head_len,
                                    const int32_t *__restrict__ own_row, int64_t n_items,
                                    int4 *__restrict__ items)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_items) return;
    const int32_t v = order[i];
    items[i] = make_int4(indptr[v], head_len[v], own_row ? own_row[v] : v, v);
}

/* work of the items in launch order (padded entries + a per-item constant), input of the scan
 * that cuts the launch order into one range per CTA */
__global__ void term_work_kernel(const int32_t *__restrict__ order, const int32_t *__restrict__ head_len,
                                 int64_t n_items, int32_t *__restrict__ work)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n_items) return;
    work[i] = (i < n_items) ? head_len[order[i]] + 24 : 0;
}

/* cta_begin[c] = first item whose work prefix reaches c / grid of the total; block_begin[b] =
 * first item (in launch order) of block b */
__global__ void term_ranges_kernel(const int32_t *__restrict__ work_prefix, const int32_t *__restrict__ sorted_keys,
                                   int64_t n_items, int32_t grid, int32_t n_blocks, int32_t block_rows,
                                   int32_t *__restrict__ cta_begin, int32_t *__restrict__ block_begin)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t <= grid) {
        const int64_t total = work_prefix[n_items];
        const int64_t target = total * t / grid;
        int64_t lo = 0, hi = n_items;
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (work_prefix[mid] < target) lo = mid + 1; else hi = mid;
        }
        cta_begin[t] = (t == grid) ? (int32_t)n_items : (int32_t)lo;
    }
    if (t <= n_blocks) {
        const int per = block_rows / 8 + 1;
        const int32_t key = t * per;
        int64_t lo = 0, hi = n_items;
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (sorted_keys[mid] < key) lo = mid + 1; else hi = mid;
        }
        block_begin[t] = (int32_t)lo;
    }
}

/* padded [rows, stride] -> compact [rows, pitch] (the TMA source of the term side's tiles) */
__global__ void compact_rows_kernel(const float *__restrict__ src, int64_t rows, int stride, int kp,
                                    float *__restrict__ dst, int pitch_f)
{
    const int nv = kp >> 2;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * nv) return;
    const int64_t r = i / nv;
    const int c = (int)(i - r * nv);
    reinterpret_cast<float4 *>(dst + r * pitch_f)[c] = reinterpret_cast<const float4 *>(src + r * stride)[c];
}

/* P(w|z)^T rows times the per-topic scale, in place (plsa.py:196-198), and the compact image
 * of the tile rows for the TMA copy; kp/4 threads per row */
__global__ void normalise_rows_kernel(float *__restrict__ mat, int64_t rows, int stride, int kp,
                                      const float *__restrict__ scale,
                                      const int32_t *__restrict__ slot_of, float *__restrict__ tile_img,
                                      int pitch_f)
{
    const int nv = kp >> 2;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * nv) return;
    const int64_t r = i / nv;
    const int c = (int)(i - r * nv);
    float4 *p = reinterpret_cast<float4 *>(mat + r * stride) + c;
    float4 v = *p;
    if (scale) {
        const float4 s = *reinterpret_cast<const float4 *>(scale + 4 * c);
        v.x *= s.x; v.y *= s.y; v.z *= s.z; v.w *= s.w;
        *p = v;
    }
    const int sl = slot_of[r];
    if (sl >= 0) reinterpret_cast<float4 *>(tile_img + (int64_t)sl * pitch_f)[c] = v;
}

} // namespace plsa
