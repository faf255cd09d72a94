/*
 * plsa_kernels.cuh — sm_100a device code of the pLSA EM engine.
 *
 * One generic "row pass" kernel carries the whole EM iteration.  It walks a sparse matrix
 * whose rows OWN one factor row (k floats) and whose entries GATHER a row of the other
 * factor:
 *
 *   doc pass   rows = documents (CSR of X)   own = P(z|d)[d,:]   gather = P(w|z)^T[w,:]
 *   term pass  rows = terms     (CSR of X^T) own = P(w|z)^T[w,:] gather = P(z|d)[d,:]
 *
 * For every stored entry x of the row it forms the thresholded products
 * v_z = own_z * gather_z (enstop/plsa.py:95-102), the normaliser sum_z v_z (plsa.py:100),
 * and adds x * v_z / norm to the row's k accumulators — that is the E-step posterior
 * (plsa.py:104-105) consumed immediately by the M-step sums (plsa.py:182-194) without the
 * nnz x k P(z|d,w) array ever existing in memory.  The doc pass row-normalises its result
 * (plsa.py:199-202); the term pass leaves raw sums and a column-sum kernel produces the
 * per-topic normalisers (plsa.py:196-198) that the NEXT pass folds into the owned row.
 *
 * Lane mapping: a row of k floats is KV*G float4 vectors; G lanes cooperate on one stored
 * entry (lane j of the group holds topics 4j..4j+3), 32/G entries are in flight per warp
 * step, the normaliser is a G-lane shuffle reduction.  k = 20 -> G = 5, six entries per
 * step; k = 128 -> G = 32.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace plsa {

struct Item {          /* one unit of warp work: a row, or a chunk of a long row        */
    int64_t start;     /* first stored entry                                            */
    int32_t row;       /* row whose factor this item owns                               */
    int32_t len;       /* stored entries in this item                                   */
    int32_t slot;      /* < 0: whole row, result goes to own_new; else partial-sum slot */
    int32_t pad;
};

enum { MODE_DOC = 0, MODE_TERM = 1, MODE_LOGLIK = 2 };

struct PassArgs {
    const Item *items;
    int64_t n_items;
    const int2 *ent;         /* stored entries {gather-row index, float bits of the value};
                                the array carries ENT_SLACK readable entries past its end   */
    const float *own_old;    /* [rows, stride_own]                                      */
    const float *gat_old;    /* [cols, stride_gat]                                      */
    const float *own_scale;  /* [kp] folded into the owned row (1/column-sum of P(w|z)) */
    float *own_new;          /* [rows, stride_own]                                      */
    float *partial;          /* [slots, kp] raw sums of split rows                      */
    const float *row_weight; /* MODE_LOGLIK: sample_weight[d]                           */
    double *ll_partial;      /* MODE_LOGLIK: one double per CTA                         */
    int32_t stride_own, stride_gat, kp;
    float thresh;
    cudaTextureObject_t gat_tex; /* gat_old as a linear float4 texture (VAR_TEX kernels)   */
    unsigned int *ticket;        /* MODE_LOGLIK: zeroed counter, last CTA does the final sum */
    double *ll_out;              /* MODE_LOGLIK: the log-likelihood                          */
};

/* kernel variants (template parameter VAR) */
enum { VAR_TEX = 1,  /* gather through the texture pipe instead of LDG                       */
       VAR_GROW = 2, /* group-per-row kernel: each G-lane group walks its own row             */
       VAR_REGS = 4, /* allow ~85 registers (3 CTAs/SM) instead of 64 (4 CTAs/SM)            */
       VAR_X_NOTHRESH = 8,  /* TIMING EXPERIMENT ONLY (wrong results): skip the threshold   */
       VAR_X_NOSHFL = 16    /* TIMING EXPERIMENT ONLY (wrong results): skip the group sum   */ };

constexpr int ENT_SLACK = 256;

__device__ __forceinline__ float4 ldg_f4(const float *p)
{
    return __ldg(reinterpret_cast<const float4 *>(p));
}
__device__ __forceinline__ float4 ldg_f4_bytes(const char *p)
{
    return __ldg(reinterpret_cast<const float4 *>(p));
}
/* keep a per-lane 64-bit base in ordinary registers so that base + idx * stride is one
 * IMAD.WIDE (ptxas otherwise splits it into a uniform base plus two carry adds per load) */
__device__ __forceinline__ const char *opaque_ptr(const char *p)
{
    asm volatile("" : "+l"(p));
    return p;
}
__device__ __forceinline__ float rcp_fast(float x)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

/* Sum over the G lanes of a group; every lane of the group receives the total.
 * Power-of-two G: xor butterfly.  Other G: cyclic windows W_s(j) = v_j + .. + v_{j+s-1}
 * (indices mod G) built by doubling, then one window per set bit of G is combined. */
template <int G>
__device__ __forceinline__ float group_sum(float v, int gbase, int j)
{
    if constexpr (G == 1) {
        return v;
    } else if constexpr ((G & (G - 1)) == 0) {
#pragma unroll
        for (int off = G / 2; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        return v;
    } else {
        float w = v, total = 0.f;
#pragma unroll
        for (int b = 0; (1 << b) <= G; ++b) {
            const int s = 1 << b;
            if (G & s) {
                const int off = G & ~((s << 1) - 1); /* set bits above b */
                if (off == 0)
                    total += w;
                else
                    total += __shfl_sync(0xffffffffu, w, gbase + (j + off) % G);
            }
            if ((s << 1) <= G) w += __shfl_sync(0xffffffffu, w, gbase + (j + s) % G);
        }
        return total;
    }
}

#ifndef PLSA_U_OVERRIDE
#define PLSA_U_OVERRIDE 0
#endif
/* Log-likelihood reduction, deterministic: lanes (butterfly), warps in order, one double per
 * CTA; the CTA that arrives last adds the per-CTA values in index order (fixed tree). */
__device__ __forceinline__ void finish_loglik(const PassArgs &a, double ll_acc)
{
    __shared__ double sm[256];
    __shared__ bool last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) ll_acc += __shfl_xor_sync(0xffffffffu, ll_acc, off);
    if (lane == 0) sm[warp] = ll_acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sm[i];
        a.ll_partial[blockIdx.x] = t;
        __threadfence();
        last = atomicAdd(a.ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    double s = 0.0;
    for (unsigned i = threadIdx.x; i < gridDim.x; i += blockDim.x) s += a.ll_partial[i];
    sm[threadIdx.x] = s;
    __syncthreads();
    for (int off = (int)blockDim.x >> 1; off > 0; off >>= 1) {
        if ((int)threadIdx.x < off) sm[threadIdx.x] += sm[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        *a.ll_out = sm[0];
        *a.ticket = 0u;
    }
}

template <int G, int KV> struct PassShape {
    static constexpr int NG = 32 / G;                       /* entries per warp step      */
    static constexpr int U0 = (PLSA_U_OVERRIDE && KV == 1 && NG < 8) ? PLSA_U_OVERRIDE
                              : (KV >= 4) ? 1 : (KV == 2) ? 2 : (NG >= 8) ? 2 : (NG >= 4) ? 3 : 4;
    static constexpr int U = (NG * U0 > 32) ? (32 / NG) : U0; /* a chunk is <= one entry per lane */
    static constexpr int CH = NG * U;                       /* entries per loop iteration */
};

/* One loop iteration: U steps of NG entries.  TAIL: entries at or past `len` are read (the
 * next row's, or the slack) but their value is forced to 0 so they add nothing. */
template <int G, int KV>
__device__ __forceinline__ void load_entries(const int2 *__restrict__ ent, int base, int grp,
                                             int2 (&e)[PassShape<G, KV>::U])
{
#pragma unroll
    for (int u = 0; u < PassShape<G, KV>::U; ++u)
        e[u] = __ldg(ent + base + u * PassShape<G, KV>::NG + grp);
}

/* issue the gathers of one iteration: U steps x KV float4 per lane */
template <int G, int KV, int VAR>
__device__ __forceinline__ void issue_gathers(const PassArgs &a,
                                              const int2 (&e)[PassShape<G, KV>::U],
                                              const char *gat_base, const uint32_t (&lane_off)[KV],
                                              uint32_t stride_bytes,
                                              float4 (&g)[PassShape<G, KV>::U][KV])
{
    constexpr int U = PassShape<G, KV>::U;
#pragma unroll
    for (int u = 0; u < U; ++u) {
        if constexpr (VAR & VAR_TEX) {
            const int t0 = e[u].x * (int)(stride_bytes >> 4);
#pragma unroll
            for (int q = 0; q < KV; ++q)
                g[u][q] = tex1Dfetch<float4>(a.gat_tex, t0 + (int)(lane_off[q] >> 4));
        } else if constexpr (KV == 1) { /* lane offset is folded into gat_base */
            g[u][0] = ldg_f4_bytes(gat_base + (uint64_t)(uint32_t)e[u].x * stride_bytes);
        } else {
            const char *row = gat_base + (uint64_t)(uint32_t)e[u].x * stride_bytes;
#pragma unroll
            for (int q = 0; q < KV; ++q) g[u][q] = ldg_f4_bytes(row + lane_off[q]);
        }
    }
}

/* E-step + M-step sums of one iteration's entries.  TAIL: entries at or past `len` were
 * read (the next row's, or the slack) but their value is forced to 0 so they add nothing. */
template <int G, int KV, int MODE, bool TAIL, int VAR = 0>
__device__ __forceinline__ void consume_iteration(const int2 (&e)[PassShape<G, KV>::U],
                                                  float4 (&g)[PassShape<G, KV>::U][KV], int base,
                                                  int len, const float4 (&own)[KV],
                                                  float4 (&acc)[KV], double &ll_acc, float rw,
                                                  float thresh, int grp, int j, int gbase)
{
    constexpr int NG = PassShape<G, KV>::NG;
    constexpr int U = PassShape<G, KV>::U;
#pragma unroll
    for (int u = 0; u < U; ++u) {
        float x = __int_as_float(e[u].y);
        if constexpr (TAIL) x = (base + u * NG + grp < len) ? x : 0.f;
        float part;
#pragma unroll
        for (int q = 0; q < KV; ++q) {
            float4 v;
            v.x = g[u][q].x * own[q].x;
            v.y = g[u][q].y * own[q].y;
            v.z = g[u][q].z * own[q].z;
            v.w = g[u][q].w * own[q].w;
            if constexpr (MODE != MODE_LOGLIK && !(VAR & VAR_X_NOTHRESH)) { /* plsa.py:98-102 */
                v.x = v.x > thresh ? v.x : 0.f;
                v.y = v.y > thresh ? v.y : 0.f;
                v.z = v.z > thresh ? v.z : 0.f;
                v.w = v.w > thresh ? v.w : 0.f;
            }
            g[u][q] = v;
            const float s4 = (v.x + v.y) + (v.z + v.w);
            part = (q == 0) ? s4 : part + s4;
        }
        const float norm = (VAR & VAR_X_NOSHFL) ? part : group_sum<G>(part, gbase, j);
        if constexpr (MODE == MODE_LOGLIK) {
            /* plsa.py:383-384; one lane per entry contributes, x == 0 marks a non-entry */
            if (j == 0 && grp < NG && x != 0.f) ll_acc += (double)(x * __logf(norm) * rw);
        } else {
            /* plsa.py:104: posterior = v / norm if norm > 0.  Products that survive the
             * threshold are normal floats (the host passes thresh >= FLT_MIN), so norm is
             * 0 or normal; norm == 0 gives x * inf (or NaN), clamped to a finite c that
             * multiplies v == 0. */
            const float c = fminf(x * rcp_fast(norm), 3.0e38f);
#pragma unroll
            for (int q = 0; q < KV; ++q) {
                acc[q].x = fmaf(c, g[u][q].x, acc[q].x);
                acc[q].y = fmaf(c, g[u][q].y, acc[q].y);
                acc[q].z = fmaf(c, g[u][q].z, acc[q].z);
                acc[q].w = fmaf(c, g[u][q].w, acc[q].w);
            }
        }
    }
}

template <int G, int KV, int MODE, int VAR>
__global__ void __launch_bounds__(256, (KV == 1) ? ((VAR & VAR_REGS) ? 3 : 4) : (KV == 2) ? 2 : 1)
    row_pass_kernel(const PassArgs a)
{
    constexpr int NG = PassShape<G, KV>::NG;
    constexpr int CH = PassShape<G, KV>::CH;

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int64_t item_id = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;

    const int grp = lane / G;      /* == NG for the 32 % G idle lanes: they read valid memory,
                                      own zeros, and are never folded in                    */
    const int j = lane - grp * G;
    const int gbase = grp * G;

    double ll_acc = 0.0;
    if (item_id < a.n_items) {
        const Item it = a.items[item_id];

        /* the owned row, with the lazily applied P(w|z) normaliser folded in */
        float4 own[KV];
#pragma unroll
        for (int q = 0; q < KV; ++q) {
            const int c = 4 * (j + G * q);
            if (grp < NG && c < a.kp) {
                const float4 o = ldg_f4(a.own_old + (int64_t)it.row * a.stride_own + c);
                const float4 s = ldg_f4(a.own_scale + c);
                own[q] = make_float4(o.x * s.x, o.y * s.y, o.z * s.z, o.w * s.w);
            } else {
                own[q] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        float4 acc[KV];
#pragma unroll
        for (int q = 0; q < KV; ++q) acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        float rw = 1.f;
        if constexpr (MODE == MODE_LOGLIK) rw = a.row_weight[it.row];

        const int2 *ent = a.ent + it.start;
        const int len = it.len;
        const float thresh = a.thresh;
        const uint32_t stride_bytes = (uint32_t)a.stride_gat * 4u;
        /* lanes whose 4 topics lie in the padding (c >= kp) re-read the row's first vector:
         * always in bounds, and their `own` is zero */
        uint32_t lane_off[KV];
#pragma unroll
        for (int q = 0; q < KV; ++q) {
            const int c = 4 * (j + G * q);
            lane_off[q] = (c < a.kp) ? (uint32_t)c * 4u : 0u;
        }
        const char *gat_base = opaque_ptr(reinterpret_cast<const char *>(a.gat_old) +
                                          (KV == 1 ? lane_off[0] : 0u));

        constexpr int U = PassShape<G, KV>::U;
        /* Entries are fetched by one broadcast 8-byte load per group per step; reads past
         * the row end land in the next row or in the array's slack and are ignored. */
        {
            int2 e[U];
            load_entries<G, KV>(ent, 0, grp, e);
            int base = 0;
            for (; base + CH <= len; base += CH) {
                int2 en[U];
                float4 g[U][KV];
                load_entries<G, KV>(ent, base + CH, grp, en);
                issue_gathers<G, KV, VAR>(a, e, gat_base, lane_off, stride_bytes, g);
                consume_iteration<G, KV, MODE, false, VAR>(e, g, base, len, own, acc, ll_acc, rw,
                                                      thresh, grp, j, gbase);
#pragma unroll
                for (int u = 0; u < U; ++u) e[u] = en[u];
            }
            if (base < len) {
                float4 g[U][KV];
                issue_gathers<G, KV, VAR>(a, e, gat_base, lane_off, stride_bytes, g);
                consume_iteration<G, KV, MODE, true, VAR>(e, g, base, len, own, acc, ll_acc, rw, thresh,
                                                     grp, j, gbase);
            }
        }

        if constexpr (MODE != MODE_LOGLIK) {
            /* fold the NG groups: group 0 ends up with the row's sums */
#pragma unroll
            for (int off = G; off < 32; off <<= 1) {
#pragma unroll
                for (int q = 0; q < KV; ++q) {
                    const float tx = __shfl_down_sync(0xffffffffu, acc[q].x, off & 31);
                    const float ty = __shfl_down_sync(0xffffffffu, acc[q].y, off & 31);
                    const float tz = __shfl_down_sync(0xffffffffu, acc[q].z, off & 31);
                    const float tw = __shfl_down_sync(0xffffffffu, acc[q].w, off & 31);
                    if (lane + off < NG * G) {
                        acc[q].x += tx; acc[q].y += ty; acc[q].z += tz; acc[q].w += tw;
                    }
                }
            }
            float inv = 1.f;
            if constexpr (MODE == MODE_DOC) {
                if (it.slot < 0) { /* plsa.py:199-202: divide by the row's total if > 0 */
                    float part = 0.f;
#pragma unroll
                    for (int q = 0; q < KV; ++q)
                        part += (acc[q].x + acc[q].y) + (acc[q].z + acc[q].w);
                    float tot = (grp == 0) ? part : 0.f;
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1)
                        tot += __shfl_xor_sync(0xffffffffu, tot, off);
                    inv = tot > 0.f ? 1.f / tot : 1.f;
                }
            }
            if (grp == 0) {
                float *dst = (it.slot < 0)
                                 ? a.own_new + (int64_t)it.row * a.stride_own
                                 : a.partial + (int64_t)it.slot * a.kp;
#pragma unroll
                for (int q = 0; q < KV; ++q) {
                    const int c = 4 * (j + G * q);
                    if (c < a.kp)
                        *reinterpret_cast<float4 *>(dst + c) = make_float4(
                            acc[q].x * inv, acc[q].y * inv, acc[q].z * inv, acc[q].w * inv);
                }
            }
        }
    }

    if constexpr (MODE == MODE_LOGLIK) finish_loglik(a, ll_acc);
}

/* ==========================================================================================
 * Group-per-row variant.  Instead of one row per warp (32/G entries of the SAME row per step,
 * folded at the row end), every G-lane group owns its OWN row and walks it one entry per
 * step: a warp carries 32/G rows at once.  The serial latency chain at the start of a row
 * (work item -> entries + owned row -> first gathered rows) is then paid by 32/G rows
 * concurrently, and the end-of-row fold across groups disappears.  Items are sorted by
 * length, so the rows of a warp are of (nearly) equal length.
 * ========================================================================================== */
template <int G, int KV, int MODE, int VAR>
__global__ void __launch_bounds__(256, (KV == 1) ? 4 : (KV == 2) ? 2 : 1)
    row_group_kernel(const PassArgs a)
{
    constexpr int NG = 32 / G;
    constexpr int U = (KV >= 4) ? 1 : (KV == 2) ? 2 : 4;   /* entries of a row in flight */

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int grp_raw = lane / G;
    const int grp = grp_raw < NG ? grp_raw : NG - 1;        /* idle lanes shadow the last group */
    const int j = lane - grp_raw * G;
    const int gbase = grp_raw * G;
    const bool lane_on = grp_raw < NG;

    const int64_t item_id = ((int64_t)blockIdx.x * (blockDim.x >> 5) + warp) * NG + grp;
    const bool has = item_id < a.n_items;
    Item it;
    it.start = 0; it.row = 0; it.len = 0; it.slot = -1; it.pad = 0;
    if (has) it = a.items[item_id];

    int maxlen = it.len;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1)
        maxlen = max(maxlen, __shfl_xor_sync(0xffffffffu, maxlen, off));

    float4 own[KV], acc[KV];
    uint32_t lane_off[KV];
#pragma unroll
    for (int q = 0; q < KV; ++q) {
        const int c = 4 * (j + G * q);
        lane_off[q] = (c < a.kp) ? (uint32_t)c * 4u : 0u;
        acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (has && lane_on && c < a.kp) {
            const float4 o = ldg_f4(a.own_old + (int64_t)it.row * a.stride_own + c);
            const float4 s = ldg_f4(a.own_scale + c);
            own[q] = make_float4(o.x * s.x, o.y * s.y, o.z * s.z, o.w * s.w);
        } else {
            own[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    float rw = 1.f;
    if constexpr (MODE == MODE_LOGLIK) rw = has ? a.row_weight[it.row] : 0.f;

    const int2 *ent = a.ent + it.start;
    const int len = it.len;
    const float thresh = a.thresh;
    const uint32_t stride_bytes = (uint32_t)a.stride_gat * 4u;
    const char *gat_base = opaque_ptr(reinterpret_cast<const char *>(a.gat_old) +
                                      (KV == 1 ? lane_off[0] : 0u));
    double ll_acc = 0.0;

    /* entries one iteration ahead; reads past a row's end land in following rows or in the
     * array's slack and get value 0 */
    int2 e[U];
#pragma unroll
    for (int u = 0; u < U; ++u) e[u] = __ldg(ent + u);
    for (int base = 0; base < maxlen; base += U) {
        int2 en[U];
#pragma unroll
        for (int u = 0; u < U; ++u) en[u] = __ldg(ent + base + U + u);
        float4 g[U][KV];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if constexpr (VAR & VAR_TEX) {
                const int t0 = e[u].x * (int)(stride_bytes >> 4);
#pragma unroll
                for (int q = 0; q < KV; ++q)
                    g[u][q] = tex1Dfetch<float4>(a.gat_tex, t0 + (int)(lane_off[q] >> 4));
            } else if constexpr (KV == 1) {
                g[u][0] = ldg_f4_bytes(gat_base + (uint64_t)(uint32_t)e[u].x * stride_bytes);
            } else {
                const char *row = gat_base + (uint64_t)(uint32_t)e[u].x * stride_bytes;
#pragma unroll
                for (int q = 0; q < KV; ++q) g[u][q] = ldg_f4_bytes(row + lane_off[q]);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const float x = (base + u < len) ? __int_as_float(e[u].y) : 0.f;
            float part;
#pragma unroll
            for (int q = 0; q < KV; ++q) {
                float4 v;
                v.x = g[u][q].x * own[q].x;
                v.y = g[u][q].y * own[q].y;
                v.z = g[u][q].z * own[q].z;
                v.w = g[u][q].w * own[q].w;
                if constexpr (MODE != MODE_LOGLIK) { /* plsa.py:98-102 */
                    v.x = v.x > thresh ? v.x : 0.f;
                    v.y = v.y > thresh ? v.y : 0.f;
                    v.z = v.z > thresh ? v.z : 0.f;
                    v.w = v.w > thresh ? v.w : 0.f;
                }
                g[u][q] = v;
                const float s4 = (v.x + v.y) + (v.z + v.w);
                part = (q == 0) ? s4 : part + s4;
            }
            const float norm = group_sum<G>(part, gbase, j);
            if constexpr (MODE == MODE_LOGLIK) {
                if (j == 0 && lane_on && x != 0.f) ll_acc += (double)(x * __logf(norm) * rw);
            } else {
                const float c = fminf(x * rcp_fast(norm), 3.0e38f); /* see consume_iteration */
#pragma unroll
                for (int q = 0; q < KV; ++q) {
                    acc[q].x = fmaf(c, g[u][q].x, acc[q].x);
                    acc[q].y = fmaf(c, g[u][q].y, acc[q].y);
                    acc[q].z = fmaf(c, g[u][q].z, acc[q].z);
                    acc[q].w = fmaf(c, g[u][q].w, acc[q].w);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) e[u] = en[u];
    }

    if constexpr (MODE != MODE_LOGLIK) {
        float inv = 1.f;
        if constexpr (MODE == MODE_DOC) { /* plsa.py:199-202: divide by the row's total if > 0 */
            float part = 0.f;
#pragma unroll
            for (int q = 0; q < KV; ++q) part += (acc[q].x + acc[q].y) + (acc[q].z + acc[q].w);
            const float tot = group_sum<G>(part, gbase, j);
            inv = (it.slot < 0 && tot > 0.f) ? 1.f / tot : 1.f;
        }
        if (has && lane_on) {
            float *dst = (it.slot < 0) ? a.own_new + (int64_t)it.row * a.stride_own
                                       : a.partial + (int64_t)it.slot * a.kp;
#pragma unroll
            for (int q = 0; q < KV; ++q) {
                const int c = 4 * (j + G * q);
                if (c < a.kp)
                    *reinterpret_cast<float4 *>(dst + c) = make_float4(
                        acc[q].x * inv, acc[q].y * inv, acc[q].z * inv, acc[q].w * inv);
            }
        }
    } else {
        finish_loglik(a, ll_acc);
    }
}

/* Split rows: add the partial sums of each split row in slot order (float64), write the
 * row; document rows are normalised here (plsa.py:199-202). */
struct FixArgs {
    const int32_t *rows;       /* [n_split]   */
    const int32_t *slot_begin; /* [n_split+1] */
    const float *partial;      /* [slots, kp] */
    float *own_new;
    int32_t n_split, kp, stride_own, normalise;
};

/* one warp per split row: lane l adds slots l, l+32, ... (four loads in flight, combined in a
 * fixed order), then a fixed butterfly across lanes */
__global__ void __launch_bounds__(256) fixup_kernel(const FixArgs a)
{
    const int lane = threadIdx.x & 31;
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= a.n_split) return;
    const int s0 = a.slot_begin[r], s1 = a.slot_begin[r + 1];
    const float *p = a.partial;
    const int kp = a.kp;
    double inv = 1.0;
    if (a.normalise) {
        double t0 = 0.0, t1 = 0.0;
        for (int sl = s0 + lane; sl < s1; sl += 32) {
            const float4 *row = reinterpret_cast<const float4 *>(p + (int64_t)sl * kp);
            int c = 0;
            for (; c + 1 < kp / 4; c += 2) {
                const float4 v = row[c], w = row[c + 1];
                t0 += ((double)v.x + (double)v.y) + ((double)v.z + (double)v.w);
                t1 += ((double)w.x + (double)w.y) + ((double)w.z + (double)w.w);
            }
            if (c < kp / 4) {
                const float4 v = row[c];
                t0 += ((double)v.x + (double)v.y) + ((double)v.z + (double)v.w);
            }
        }
        double t = t0 + t1;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
        inv = t > 0.0 ? 1.0 / t : 1.0;
    }
    float *dst = a.own_new + (int64_t)a.rows[r] * a.stride_own;
    for (int zb = 0; zb < kp; zb += 8) {
        const bool two = zb + 8 <= kp; /* kp is a multiple of 4: a block is 8 or 4 floats */
        double acc[4][8];
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[u][i] = 0.0;
        for (int sl = s0 + lane; sl < s1; sl += 128) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int s = sl + 32 * u;
                if (s < s1) {
                    const float4 *row = reinterpret_cast<const float4 *>(p + (int64_t)s * kp + zb);
                    const float4 v = row[0];
                    acc[u][0] += (double)v.x; acc[u][1] += (double)v.y;
                    acc[u][2] += (double)v.z; acc[u][3] += (double)v.w;
                    if (two) {
                        const float4 w = row[1];
                        acc[u][4] += (double)w.x; acc[u][5] += (double)w.y;
                        acc[u][6] += (double)w.z; acc[u][7] += (double)w.w;
                    }
                }
            }
        }
        double mine = 0.0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            double t = (acc[0][i] + acc[1][i]) + (acc[2][i] + acc[3][i]);
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
            if (lane == i) mine = t;
        }
        if (lane < (two ? 8 : 4)) dst[zb + lane] = (float)(mine * inv);
    }
}

/* Column sums of the raw P(w|z)^T accumulators -> per-topic scale 1/sum (plsa.py:196-198).
 * Deterministic: per-CTA float64 partials over a slab of rows; the CTA that arrives last
 * adds the per-CTA values in index order. */
constexpr int COLSUM_CTAS = 592;

__global__ void __launch_bounds__(256) colsum_kernel(const float *__restrict__ B, int64_t n_rows,
                                                     int stride, int kp,
                                                     double *__restrict__ partial,
                                                     unsigned int *ticket, float *__restrict__ scale,
                                                     double *__restrict__ colnorm)
{
    __shared__ double sm[256];
    __shared__ bool last;
    const int64_t per = (n_rows + gridDim.x - 1) / gridDim.x;
    const int64_t r0 = per * blockIdx.x;
    const int64_t r1 = min(n_rows, r0 + per);
    for (int zb = 0; zb < kp; zb += 256) {
        const int width = min(256, kp - zb);       /* columns handled in this sweep    */
        const int rl = 256 / width;                /* row lanes                        */
        const int z = threadIdx.x % width, rr = threadIdx.x / width;
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        if (rr < rl) {
            const float *col = B + zb + z;
            int64_t r = r0 + rr;
            for (; r + 3 * rl < r1; r += 4 * rl) { /* four independent loads in flight */
                const float v0 = col[r * stride], v1 = col[(r + rl) * stride],
                            v2 = col[(r + 2 * rl) * stride], v3 = col[(r + 3 * rl) * stride];
                s0 += (double)v0; s1 += (double)v1; s2 += (double)v2; s3 += (double)v3;
            }
            for (; r < r1; r += rl) s0 += (double)col[r * stride];
        }
        sm[threadIdx.x] = (s0 + s1) + (s2 + s3);
        __syncthreads();
        if (threadIdx.x < width) {
            double t = 0.0;
            for (int i = 0; i < rl; ++i) t += sm[i * width + threadIdx.x];
            partial[(int64_t)blockIdx.x * kp + zb + threadIdx.x] = t;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        __threadfence();
        last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int z = warp; z < kp; z += nw) { /* one warp per topic, fixed-order butterfly */
        double t = 0.0;
        for (int i = lane; i < (int)gridDim.x; i += 32) t += partial[(int64_t)i * kp + z];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
        if (lane == 0) {
            colnorm[z] = t;
            scale[z] = t > 0.0 ? (float)(1.0 / t) : 1.f;
        }
    }
    if (threadIdx.x == 0) *ticket = 0u;
}

/* ---- layout conversion between the reference's arrays and the device layout ---------- */
/* dense [rows, k] -> padded [rows, stride] (P(z|d); also P(w|z)^T when src is [k, rows]) */
__global__ void pack_rows_kernel(const float *__restrict__ src, float *__restrict__ dst,
                                 int64_t rows, int k, int kp, int stride, int transposed)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * kp) return;
    const int64_t r = i / kp;
    const int z = (int)(i - r * kp);
    float v = 0.f;
    if (z < k) v = transposed ? src[(int64_t)z * rows + r] : src[r * k + z];
    dst[r * stride + z] = v;
}

/* padded [rows, stride] (* scale[z]) -> dense [rows, k] or its transpose [k, rows] */
__global__ void unpack_rows_kernel(const float *__restrict__ src, const float *__restrict__ scale,
                                   float *__restrict__ dst, int64_t rows, int k, int stride,
                                   int transposed)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * k) return;
    if (transposed) { /* consecutive threads -> consecutive rows of one topic: coalesced writes */
        const int z = (int)(i / rows);
        const int64_t r = i - (int64_t)z * rows;
        dst[i] = src[r * stride + z] * (scale ? scale[z] : 1.f);
    } else {
        const int64_t r = i / k;
        const int z = (int)(i - r * k);
        dst[i] = src[r * stride + z] * (scale ? scale[z] : 1.f);
    }
}

/* ---- corpus preparation ------------------------------------------------------------------ */
/* separate index / value arrays (the caller's CSR) -> interleaved entries */
__global__ void interleave_kernel(const int32_t *__restrict__ cols, const float *__restrict__ vals,
                                  int64_t n, int2 *__restrict__ ent)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) ent[i] = make_int2(cols[i], __float_as_int(vals[i]));
}

/* per entry: its row (expanded indptr) and its column as a sort key; one warp per row */
__global__ void expand_rows_kernel(const int32_t *__restrict__ indptr, int64_t n_rows,
                                   const int2 *__restrict__ ent, int32_t *__restrict__ rows_out,
                                   int32_t *__restrict__ keys_out)
{
    const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= n_rows) return;
    const int lane = threadIdx.x & 31;
    for (int32_t p = indptr[r] + lane; p < indptr[r + 1]; p += 32) {
        rows_out[p] = (int32_t)r;
        keys_out[p] = ent[p].x;
    }
}

/* term-major entries from the stable sort permutation: {document, value} */
__global__ void permute_kernel(const int32_t *__restrict__ perm, int64_t n,
                               const int32_t *__restrict__ rows_in, const int2 *__restrict__ ent_in,
                               int2 *__restrict__ ent_out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t p = perm[i];
    ent_out[i] = make_int2(rows_in[p], ent_in[p].y);
}

/* column pointers of the sorted keys: indptr[w] = first position with key >= w */
__global__ void lower_bound_kernel(const int32_t *__restrict__ sorted_keys, int64_t n,
                                   int64_t n_cols, int32_t *__restrict__ indptr)
{
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w > n_cols) return;
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (sorted_keys[mid] < (int32_t)w) lo = mid + 1; else hi = mid;
    }
    indptr[w] = (int32_t)lo;
}

__global__ void iota_kernel(int32_t *p, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = (int32_t)i;
}

/* P(w|z) receives s * sample_weight[d] (plsa.py:293-297): pre-weighted term-major values */
__global__ void weight_vals_kernel(const int2 *__restrict__ ent, const float *__restrict__ w,
                                   int2 *__restrict__ out, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const int2 e = ent[i];
        out[i] = make_int2(e.x, __float_as_int(__int_as_float(e.y) * w[e.x]));
    }
}

/* bootstrap: new row i <- base row src[i]; one warp per new row */
__global__ void gather_rows_kernel(const int32_t *__restrict__ src, int64_t n_new,
                                   const int32_t *__restrict__ base_indptr,
                                   const int2 *__restrict__ base_ent,
                                   const int32_t *__restrict__ new_indptr, int2 *__restrict__ ent)
{
    const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= n_new) return;
    const int lane = threadIdx.x & 31;
    const int32_t b0 = base_indptr[src[r]];
    const int32_t len = base_indptr[src[r] + 1] - b0;
    const int32_t o0 = new_indptr[r];
    for (int32_t p = lane; p < len; p += 32) ent[o0 + p] = base_ent[b0 + p];
}

__global__ void fill_kernel(float *p, int64_t n, float v)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

} // namespace plsa
