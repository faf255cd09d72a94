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
 * (plsa.py:199-202); the term pass leaves raw sums and also produces the per-topic
 * normalisers (plsa.py:196-198) that the NEXT pass folds into the owned row.
 *
 * Lane mapping ("group per row"): a factor row of k floats is KV*G float4 vectors; a group of
 * G lanes owns ONE work item (a row, or a chunk of a long row) and walks it one entry per
 * step — lane j of the group holds topics 4j..4j+3 of the owned row, of the accumulators and
 * of every gathered row, so a gather is one 16-byte fetch per lane and the G lanes cover one
 * contiguous row segment (80 B at k=20).  A warp carries 32/G items at once (k=20: G=5, six
 * items; k=128: G=32, one item), U entries of each in flight; the posterior normaliser is a
 * G-lane shuffle reduction.  The serial latency chain at the start of an item (item header
 * -> entries + owned row -> first gathered rows) is thereby shared by 32/G items, and there
 * is no cross-group fold at the end.  Items are sorted by length so the items of a warp are
 * (nearly) equally long; rows longer than `chunk` entries are split so that no group walks
 * a long serial chain.
 *
 * Gathers go through the texture pipe (tex1Dfetch on a linear float4 view of the factor)
 * when the factor fits a 1-D linear texture: measured on B200 the LSU data path and the
 * shuffles compete for the same pipe, the texture path does not (scripts/microbench_gather.cu,
 * profiles/).  Otherwise plain read-only 16-byte loads are used.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace plsa {

struct Item {          /* one unit of group work: a row, or a chunk of a long row       */
    int64_t start;     /* first stored entry                                            */
    int32_t row;       /* row whose factor this item owns                               */
    int32_t len;       /* stored entries in this item                                   */
    int32_t slot;      /* < 0: whole row, result goes to own_new; else partial-sum slot */
    int32_t skip;      /* aligned items: `start` is rounded down to a multiple of the entry
                          block, the first `skip` entries (< block) belong to the row before
                          and count as value 0; `len` includes them.  ITEM_FIRST is or-ed in
                          for the first chunk of a split row                                */
};
constexpr int32_t ITEM_FIRST = 0x10000;

enum { MODE_DOC = 0, MODE_TERM = 1, MODE_LOGLIK = 2,
       MODE_DOC_LL = 3 /* doc pass that also returns the log-likelihood of the factors it reads */ };

/* MODE_DOC_LL takes log(sum of the THRESHOLDED products).  That equals the reference's
 * log(sum of all products) (plsa.py:381-384) to float precision as long as the dropped
 * products (each <= thresh <= 1e-30) are negligible next to the sum; a sum below this bound
 * raises PassArgs::flag and the host recomputes with the exact MODE_LOGLIK pass. */
#define PLSA_FUSED_LL_MIN_NORM 1e-20f
#define PLSA_FUSED_LL_MAX_THRESH 1e-30f

struct PassArgs {
    const Item *items;
    int64_t n_items;
    const int2 *ent;         /* stored entries {gather-row index, float bits of the value};
                                readable (zeroed) padding follows the last entry          */
    const float *own_old;    /* [rows, stride_own]                                      */
    const float *gat_old;    /* [cols, stride_gat]                                      */
    const float *own_scale;  /* [kp] folded into the owned row (1/column-sum of P(w|z)) */
    float *own_new;          /* [rows, stride_own]                                      */
    float *partial;          /* [slots, kp] raw sums of split rows                      */
    const float *row_weight; /* MODE_LOGLIK / MODE_DOC_LL: sample_weight[d]                */
    double *cta_partial;     /* MODE_LOGLIK: [grid]; MODE_TERM: [grid, kp] per-CTA sums  */
    unsigned int *ticket;    /* zeroed counters (1 + grid/32): last-arrival reductions       */
    double *ll_out;          /* MODE_LOGLIK / MODE_DOC_LL: the log-likelihood              */
    int *flag;               /* MODE_DOC_LL: set when the fused value cannot be trusted    */
    float *scale_out;        /* MODE_TERM: [kp] 1 / column sum of the new P(w|z)         */
    double *colnorm_out;     /* MODE_TERM: [kp] the column sums                          */
    cudaTextureObject_t gat_tex; /* gat_old as a linear float4 texture (TEX kernels)     */
    const float *add_partial;    /* tiled doc pass: [rows, kp] sums of the row's head entries
                                    (plsa_tile.cuh), added to the row's first item           */
    int32_t stride_own, stride_gat, kp;
    float thresh;
};

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
__device__ __forceinline__ float log2_ftz(float x)
{
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float rcp_fast(float x)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

/* Packed float32 pairs (sm_100 FMUL2 / FADD2 / FFMA2: one issue slot for two lanes' worth of
 * a float4's arithmetic).  The row pass is co-limited by issue slots; the products, the
 * in-lane sum and the accumulation are 13 % of its instructions fewer this way. */
#ifndef PLSA_PACKED_MATH
#define PLSA_PACKED_MATH 1
#endif
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi)
{
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk2(f32x2 v, float &lo, float &hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b)
{
    f32x2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b)
{
    f32x2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
/* flush-to-zero product: the tiled pass (plsa_tile.cuh) applies the E-step threshold with it */
__device__ __forceinline__ f32x2 mul2_ftz(f32x2 a, f32x2 b)
{
    f32x2 r;
    asm("mul.rn.ftz.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c)
{
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

/* U consecutive entries (8 bytes each) of an item.  VEC: one 8*U-byte load (U = 4: the
 * 256-bit LDG.E.ENL2.256 of sm_100; the address is a multiple of 8*U bytes because aligned
 * items start on an entry-block boundary).  The six groups of a warp read six different
 * places, so every entry load costs one L1 wavefront per group: one wide load per block
 * instead of one 8-byte load per entry cuts the LSU data-pipe work of the pass by U. */
template <int U, bool VEC>
__device__ __forceinline__ void load_entries(const int2 *p, int2 (&e)[U])
{
    if constexpr (VEC && U == 4) {
        asm("ld.global.nc.v8.s32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
            : "=r"(e[0].x), "=r"(e[0].y), "=r"(e[1].x), "=r"(e[1].y), "=r"(e[2].x), "=r"(e[2].y),
              "=r"(e[3].x), "=r"(e[3].y)
            : "l"(p));
    } else if constexpr (VEC && U == 2) {
        const int4 v = __ldg(reinterpret_cast<const int4 *>(p));
        e[0] = make_int2(v.x, v.y);
        e[1] = make_int2(v.z, v.w);
    } else {
#pragma unroll
        for (int u = 0; u < U; ++u) e[u] = __ldg(p + u);
    }
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
        bool first = true;
#pragma unroll
        for (int b = 0; (1 << b) <= G; ++b) {
            const int s = 1 << b;
            if (G & s) {
                const int off = G & ~((s << 1) - 1); /* set bits above b */
                const float t = (off == 0) ? w : __shfl_sync(0xffffffffu, w, gbase + (j + off) % G);
                total = first ? t : total + t;
                first = false;
            }
            if ((s << 1) <= G) w += __shfl_sync(0xffffffffu, w, gbase + (j + s) % G);
        }
        return total;
    }
}

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
        a.cta_partial[blockIdx.x] = t;
        __threadfence();
        last = atomicAdd(a.ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    double s = 0.0;
    for (unsigned i = threadIdx.x; i < gridDim.x; i += blockDim.x) s += a.cta_partial[i];
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

/* Column sums of the new raw P(w|z)^T (plsa.py:196-198), fused into the term pass: the sum
 * over rows of P(w|z)^T[w,z] is the sum over ALL work items (whole rows and chunks alike) of
 * their accumulators.  Deterministic: groups of a warp are folded in a fixed shuffle tree
 * (float), the 8 warps of a CTA are added in order (float64) into cta_partial[cta, :], and
 * the CTA that arrives last adds the per-CTA values in index order and writes the scale. */
template <int G, int KV>
__device__ __forceinline__ void finish_colsum(const PassArgs &a, float4 (&acc)[KV], int lane,
                                              int warp, int j, bool lane_on)
{
    constexpr int NG = 32 / G;
    __shared__ float wsum[8][128];
    __shared__ bool last;
    /* fold the NG groups of the warp: group 0 ends up with the warp's sums */
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
    (void)lane_on;
    const int kp = a.kp;
    double *mine = a.cta_partial + (int64_t)blockIdx.x * kp;
#pragma unroll
    for (int q = 0; q < KV; ++q) {   /* 128 topics per sweep (G == 32 when KV > 1) */
        const int c = 4 * (j + G * q);
        __syncthreads();
        if (lane < G && c < kp) {
            const int cc = c - 128 * q * (G == 32 ? 1 : 0);
            wsum[warp][cc] = acc[q].x; wsum[warp][cc + 1] = acc[q].y;
            wsum[warp][cc + 2] = acc[q].z; wsum[warp][cc + 3] = acc[q].w;
        }
        __syncthreads();
        const int zb = (G == 32) ? 128 * q : 0;
        const int width = min(128, kp - zb);
        if ((int)threadIdx.x < width) {
            double t = 0.0;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += (double)wsum[w][threadIdx.x];
            mine[zb + threadIdx.x] = t;
        }
    }
    /* two-level "last arrival" reduction, summation order fixed by index:
     *   level 1: CTAs in groups of 32; the last of a group adds the group's partials
     *   level 2: the last group adds the group sums and writes the scale */
    const unsigned n_groups = (gridDim.x + 31u) >> 5;
    const unsigned grp_id = blockIdx.x >> 5;
    const unsigned grp_size = min(32u, gridDim.x - (grp_id << 5));
    unsigned int *ticket1 = a.ticket + 1 + grp_id;
    double *gpartial = a.cta_partial + (int64_t)gridDim.x * kp; /* [n_groups, kp] */
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        last = atomicAdd(ticket1, 1u) == grp_size - 1;
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    const int nw = blockDim.x >> 5;
    for (int z = warp; z < kp; z += nw) { /* one warp per topic, fixed-order butterfly */
        double t = (lane < (int)grp_size)
                       ? a.cta_partial[(int64_t)((grp_id << 5) + lane) * kp + z] : 0.0;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
        if (lane == 0) gpartial[(int64_t)grp_id * kp + z] = t;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        *ticket1 = 0u;
        __threadfence();
        last = atomicAdd(a.ticket, 1u) == n_groups - 1;
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    for (int z = warp; z < kp; z += nw) {
        double t = 0.0;
        for (unsigned i = lane; i < n_groups; i += 32) t += gpartial[(int64_t)i * kp + z];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
        if (lane == 0) {
            a.colnorm_out[z] = t;
            a.scale_out[z] = t > 0.0 ? (float)(1.0 / t) : 1.f;
        }
    }
    if (threadIdx.x == 0) *a.ticket = 0u;
}

/* U consecutive entries of every group's item: gather the factor rows, E-step, M-step sums.
 * TAIL: entries at or past the item's end (read from following rows or the padding) count
 * as value 0. */
template <int G, int KV, int U, int MODE, bool TEX, bool TAIL>
__device__ __forceinline__ void pass_block(const PassArgs &a, const int2 (&e)[U], int base, int len,
                                           const char *gat_base, const uint32_t (&lane_off)[KV],
                                           uint32_t stride_bytes, const float4 (&own)[KV],
                                           float4 (&acc)[KV], double &ll_acc, float &min_norm,
                                           float rw, float thresh, int j, int gbase, bool lane_on)
{
    float4 g[U][KV];
    float llt[U]; /* log-likelihood terms of the block (MODE_LOGLIK / MODE_DOC_LL) */
#pragma unroll
    for (int u = 0; u < U; ++u) {
        if constexpr (TEX) {
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
#if PLSA_PACKED_MATH
    f32x2 own2[KV][2], acc2[KV][2];
#pragma unroll
    for (int q = 0; q < KV; ++q) {
        own2[q][0] = pk2(own[q].x, own[q].y);
        own2[q][1] = pk2(own[q].z, own[q].w);
        acc2[q][0] = pk2(acc[q].x, acc[q].y);
        acc2[q][1] = pk2(acc[q].z, acc[q].w);
    }
#endif
#pragma unroll
    for (int u = 0; u < U; ++u) {
        float x = __int_as_float(e[u].y);
        if constexpr (TAIL) x = (base + u < len) ? x : 0.f;
        float part;
#pragma unroll
        for (int q = 0; q < KV; ++q) {
            float4 v;
#if PLSA_PACKED_MATH
            upk2(mul2(pk2(g[u][q].x, g[u][q].y), own2[q][0]), v.x, v.y);
            upk2(mul2(pk2(g[u][q].z, g[u][q].w), own2[q][1]), v.z, v.w);
#else
            v.x = g[u][q].x * own[q].x;
            v.y = g[u][q].y * own[q].y;
            v.z = g[u][q].z * own[q].z;
            v.w = g[u][q].w * own[q].w;
#endif
            if constexpr (MODE != MODE_LOGLIK) { /* plsa.py:98-102 */
                v.x = v.x > thresh ? v.x : 0.f;
                v.y = v.y > thresh ? v.y : 0.f;
                v.z = v.z > thresh ? v.z : 0.f;
                v.w = v.w > thresh ? v.w : 0.f;
            }
            g[u][q] = v;
#if PLSA_PACKED_MATH
            float s_lo, s_hi;
            upk2(add2(pk2(v.x, v.y), pk2(v.z, v.w)), s_lo, s_hi);
            const float s4 = s_lo + s_hi;
#else
            const float s4 = (v.x + v.y) + (v.z + v.w);
#endif
            part = (q == 0) ? s4 : part + s4;
        }
        const float norm = group_sum<G>(part, gbase, j);
        if constexpr (MODE == MODE_LOGLIK || MODE == MODE_DOC_LL) {
            /* plsa.py:383-384: x * log(sum) * sample_weight[d]; x == 0 marks a non-entry.
             * Branch-free, every lane of the group carries the same term (lane 0's is used);
             * the block's terms are added in float32, blocks in float64. */
            float lg; /* fused pass: the sum is 0 or a normal float (thresholded products), so the
                         flush-to-zero MUFU.LG2 needs no subnormal pre-scaling branch */
            if constexpr (MODE == MODE_DOC_LL) lg = log2_ftz(norm) * 0.69314718f;
            else lg = __logf(norm);
            llt[u] = (x != 0.f) ? x * rw * lg : 0.f;
            /* idle lanes (32 % G of them) shadow the last group with a zero owned row: their
             * sums are not sums of an entry */
            if constexpr (MODE == MODE_DOC_LL)
                min_norm = fminf(min_norm, (lane_on && x != 0.f) ? norm : 1.f);
        }
        if constexpr (MODE != MODE_LOGLIK) {
            /* plsa.py:104: posterior = v / norm if norm > 0.  Products that survive the
             * threshold are normal floats (the host passes thresh >= FLT_MIN), so norm is 0
             * or normal; norm == 0 gives x * inf (or NaN), clamped to a finite c that
             * multiplies v == 0. */
            const float c = fminf(x * rcp_fast(norm), 3.0e38f);
#if PLSA_PACKED_MATH
            const f32x2 c2 = pk2(c, c);
#pragma unroll
            for (int q = 0; q < KV; ++q) {
                acc2[q][0] = fma2(c2, pk2(g[u][q].x, g[u][q].y), acc2[q][0]);
                acc2[q][1] = fma2(c2, pk2(g[u][q].z, g[u][q].w), acc2[q][1]);
            }
#else
#pragma unroll
            for (int q = 0; q < KV; ++q) {
                acc[q].x = fmaf(c, g[u][q].x, acc[q].x);
                acc[q].y = fmaf(c, g[u][q].y, acc[q].y);
                acc[q].z = fmaf(c, g[u][q].z, acc[q].z);
                acc[q].w = fmaf(c, g[u][q].w, acc[q].w);
            }
#endif
        }
    }
    if constexpr (MODE == MODE_LOGLIK || MODE == MODE_DOC_LL) {
        float t = llt[0];
#pragma unroll
        for (int u = 1; u < U; ++u) t += llt[u];
        ll_acc += (double)t;
    }
#if PLSA_PACKED_MATH
#pragma unroll
    for (int q = 0; q < KV; ++q) {
        upk2(acc2[q][0], acc[q].x, acc[q].y);
        upk2(acc2[q][1], acc[q].z, acc[q].w);
    }
#endif
}

/* how far ahead (in entries) the row pass pulls an item's entry stream into L2; 0 = off */
#ifndef PLSA_ENT_PREFETCH
#define PLSA_ENT_PREFETCH 32
#endif
#ifndef PLSA_ENT_PREFETCH_L1
#define PLSA_ENT_PREFETCH_L1 0
#endif

/* entries of an item in flight = the entry block aligned items start on */
__host__ __device__ constexpr int pass_block_entries(int KV) { return (KV >= 4) ? 1 : (KV == 2) ? 2 : 4; }

/* CTA shape of the row pass: 8 warps, 4 CTAs per SM at 64 registers (KV == 1).  128-thread
 * CTAs at 9 per SM (36 resident warps) were measured slower (profiles/r2a_ab_c2.txt). */
constexpr int PLSA_PASS_THREADS = 256;

template <int G, int KV, int MODE, bool TEX, bool VEC>
__global__ void __launch_bounds__(PLSA_PASS_THREADS, (KV == 1) ? 4 : (KV == 2) ? 2 : 1)
    row_pass_kernel(const PassArgs a)
{
    constexpr int NG = 32 / G;
    constexpr int U = pass_block_entries(KV);

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int grp_raw = lane / G;
    const int grp = grp_raw < NG ? grp_raw : NG - 1;        /* 32 % G idle lanes shadow the
                                                               last group and never write  */
    const int j = lane - grp_raw * G;
    const int gbase = grp_raw * G;
    const bool lane_on = grp_raw < NG;

    const int64_t item_id = ((int64_t)blockIdx.x * (blockDim.x >> 5) + warp) * NG + grp;
    const bool has = item_id < a.n_items;
    Item it;
    it.start = 0; it.row = 0; it.len = 0; it.slot = -1; it.skip = 0;
    if (has) it = a.items[item_id];

    int maxlen = it.len;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1)
        maxlen = max(maxlen, __shfl_xor_sync(0xffffffffu, maxlen, off));

    /* the owned row, with the lazily applied P(w|z) normaliser folded in; lanes whose four
     * topics lie in the padding (c >= kp) own zeros and re-read the gathered row's first
     * vector, which is always in bounds */
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
    if constexpr (MODE == MODE_LOGLIK || MODE == MODE_DOC_LL) rw = has ? a.row_weight[it.row] : 0.f;

    const int2 *ent = a.ent + it.start;
    const int len = it.len;
    const float thresh = a.thresh;
    const uint32_t stride_bytes = (uint32_t)a.stride_gat * 4u;
    const char *gat_base = opaque_ptr(reinterpret_cast<const char *>(a.gat_old) +
                                      (KV == 1 ? lane_off[0] : 0u));
    double ll_acc = 0.0;
    float min_norm = 1.f; /* MODE_DOC_LL: smallest posterior normaliser of a real entry */

    /* Entries are fetched one block ahead (one broadcast 8-byte load per group per entry);
     * reads past an item's end land in following rows or in the array's padding and count
     * as value 0.  (Measured alternatives that were slower on B200: alternating entry sets
     * without the register rotation, an unpredicated main loop, two-deep software pipelining
     * of the gathers, L1 prefetch of the next rows — profiles/r1_kernel_experiments.md.) */
    int2 e[U];
    load_entries<U, VEC>(ent, e);
    if constexpr (VEC && U > 1) { /* entries before the row's first one: value 0 */
#pragma unroll
        for (int u = 0; u < U - 1; ++u) e[u].y = (u < (it.skip & 0xffff)) ? 0 : e[u].y;
    }
    for (int base = 0; base < maxlen; base += U) {
        int2 en[U];
        load_entries<U, VEC>(ent + base + U, en);
        /* the entry stream (8 bytes per entry, from HBM) is pulled into L2 ahead of its use:
         * with only the next block in registers, every block waited a memory latency for its
         * entries and then another one for the rows they point at */
        if constexpr (PLSA_ENT_PREFETCH > 0) {
            if (j == 0 && base + PLSA_ENT_PREFETCH < maxlen + U) {
#if PLSA_ENT_PREFETCH_L1
                asm volatile("prefetch.global.L1 [%0];" ::"l"(ent + base + PLSA_ENT_PREFETCH));
#else
                asm volatile("prefetch.global.L2 [%0];" ::"l"(ent + base + PLSA_ENT_PREFETCH));
#endif
            }
        }
        pass_block<G, KV, U, MODE, TEX, true>(a, e, base, len, gat_base, lane_off, stride_bytes,
                                              own, acc, ll_acc, min_norm, rw, thresh, j, gbase,
                                              lane_on);
#pragma unroll
        for (int u = 0; u < U; ++u) e[u] = en[u];
    }

    if constexpr (MODE == MODE_LOGLIK || MODE == MODE_DOC_LL) {
        if (!(j == 0 && lane_on)) ll_acc = 0.0; /* one lane per item contributes */
        if constexpr (MODE == MODE_DOC_LL)
            if (min_norm < PLSA_FUSED_LL_MIN_NORM) *a.flag = 1;
    }
    if constexpr (MODE == MODE_LOGLIK) {
        finish_loglik(a, ll_acc);
    } else {
        float inv = 1.f;
        if constexpr (MODE == MODE_DOC || MODE == MODE_DOC_LL) {
            /* tiled doc pass: the row's head entries were summed by tile_pass_kernel */
            if (a.add_partial != nullptr && has && lane_on && (it.slot < 0 || (it.skip & ITEM_FIRST))) {
#pragma unroll
                for (int q = 0; q < KV; ++q) {
                    const int c = 4 * (j + G * q);
                    if (c < a.kp) {
                        const float4 h = ldg_f4(a.add_partial + (int64_t)it.row * a.kp + c);
                        acc[q].x += h.x; acc[q].y += h.y; acc[q].z += h.z; acc[q].w += h.w;
                    }
                }
            }
        }
        if constexpr (MODE == MODE_DOC || MODE == MODE_DOC_LL) { /* plsa.py:199-202 */
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
        if constexpr (MODE == MODE_TERM) { /* nullptr: the column sums are taken elsewhere (tiled term pass) */
            if (a.cta_partial != nullptr) finish_colsum<G, KV>(a, acc, lane, warp, j, lane_on);
        }
        if constexpr (MODE == MODE_DOC_LL) finish_loglik(a, ll_acc);
    }
}

/* Split rows: add the partial sums of each split row (float64, fixed order), write the row;
 * document rows are normalised here (plsa.py:199-202).  One launch serves both factors.
 * Rows are listed heavy first: a row with more than 32 chunks gets a whole CTA (its chunks
 * spread over 256 threads), the others one warp each — so that no thread walks a long
 * serial chain of dependent loads. */
struct FixArgs {
    const int32_t *rows;       /* [n_split], rows with > 32 slots first                  */
    const int32_t *slot_begin; /* [n_split+1] */
    const float *partial;      /* [slots, kp] */
    float *own_new;
    int32_t n_split, n_heavy, kp, stride_own, normalise;
};

/* sum over the workers (a warp, or the 8 warps of a CTA) of one double per thread; every
 * thread receives the total; fixed order */
template <bool CTA>
__device__ __forceinline__ double fix_reduce(double t, double *sm /* [8] */)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
    if constexpr (CTA) {
        __syncthreads();
        if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = t;
        __syncthreads();
        t = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += sm[w];
    }
    return t;
}

template <bool CTA>
__device__ __forceinline__ void fixup_row(const FixArgs &a, int r, double *sm)
{
    const int nthr = CTA ? 256 : 32;
    const int tid = CTA ? (int)threadIdx.x : (int)(threadIdx.x & 31);
    const int s0 = a.slot_begin[r], s1 = a.slot_begin[r + 1];
    const float *p = a.partial;
    const int kp = a.kp;
    double inv = 1.0;
    if (a.normalise) {
        double t0 = 0.0, t1 = 0.0;
        for (int sl = s0 + tid; sl < s1; sl += nthr) {
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
        const double t = fix_reduce<CTA>(t0 + t1, sm);
        inv = t > 0.0 ? 1.0 / t : 1.0;
    }
    float *dst = a.own_new + (int64_t)a.rows[r] * a.stride_own;
    for (int zb = 0; zb < kp; zb += 8) {
        const bool two = zb + 8 <= kp; /* kp is a multiple of 4: a block is 8 or 4 floats */
        double acc[2][8];
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[u][i] = 0.0;
        for (int sl = s0 + tid; sl < s1; sl += 2 * nthr) {
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int s = sl + nthr * u;
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
            const double t = fix_reduce<CTA>(acc[0][i] + acc[1][i], sm);
            if (tid == i) mine = t;
        }
        if (tid < (two ? 8 : 4)) dst[zb + tid] = (float)(mine * inv);
    }
}

__device__ __forceinline__ int fix_blocks(const FixArgs &a)
{
    return a.n_heavy + (a.n_split - a.n_heavy + 7) / 8;
}

__global__ void __launch_bounds__(256) fixup_kernel(const FixArgs fa, const FixArgs fb)
{
    __shared__ double sm[8];
    int b = blockIdx.x;
    const FixArgs &a = (b < fix_blocks(fa)) ? fa : fb;
    if (b >= fix_blocks(fa)) b -= fix_blocks(fa);
    if (b < a.n_heavy) {
        fixup_row<true>(a, b, sm);
    } else {
        const int r = a.n_heavy + (b - a.n_heavy) * 8 + (int)(threadIdx.x >> 5);
        if (r < a.n_split) fixup_row<false>(a, r, sm);
    }
}

/* ---- column sums of a whole factor (document-sharded fits) ----------------------------- */
/* After the ranks' raw P(w|z)^T partial sums have been added (all-reduce), the per-topic
 * normaliser (plsa.py:196-198) is the column sum of the full matrix.  Two launches, fixed
 * summation order: each CTA adds a contiguous slice of rows in float64, one CTA adds the
 * per-CTA values in index order.  Every rank runs it on identical input -> identical scale. */
__global__ void __launch_bounds__(256) colsum_partial_kernel(const float *__restrict__ mat,
                                                             int64_t rows, int stride, int kp,
                                                             double *__restrict__ part /*[grid, kp]*/)
{
    __shared__ double sm[256];
    const int64_t per = (rows + gridDim.x - 1) / gridDim.x;
    const int64_t r0 = (int64_t)blockIdx.x * per, r1 = min(rows, r0 + per);
    if (kp <= 256) { /* 256 / kp rows at a time, kp consecutive threads along a row */
        const int nsub = 256 / kp, sub = (int)threadIdx.x / kp, z = (int)threadIdx.x - sub * kp;
        double t = 0.0;
        if (sub < nsub) { /* four independent loads in flight; the order of the sum stays fixed */
            double t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0;
            int64_t r = r0 + sub;
            for (; r + 3 * nsub < r1; r += 4 * nsub) {
                const float a0 = mat[r * stride + z], a1 = mat[(r + nsub) * stride + z],
                            a2 = mat[(r + 2 * nsub) * stride + z], a3 = mat[(r + 3 * nsub) * stride + z];
                t0 += (double)a0; t1 += (double)a1; t2 += (double)a2; t3 += (double)a3;
            }
            for (; r < r1; r += nsub) t0 += (double)mat[r * stride + z];
            t = (t0 + t1) + (t2 + t3);
        }
        sm[threadIdx.x] = t;
        __syncthreads();
        if ((int)threadIdx.x < kp) {
            double acc = 0.0;
            for (int q = 0; q < nsub; ++q) acc += sm[q * kp + threadIdx.x];
            part[(int64_t)blockIdx.x * kp + threadIdx.x] = acc;
        }
    } else {
        for (int z = threadIdx.x; z < kp; z += 256) {
            double t = 0.0;
            for (int64_t r = r0; r < r1; ++r) t += (double)mat[r * stride + z];
            part[(int64_t)blockIdx.x * kp + z] = t;
        }
    }
}

__global__ void __launch_bounds__(256) colsum_final_kernel(const double *__restrict__ part, int n_part,
                                                           int kp, float *__restrict__ scale_out,
                                                           double *__restrict__ colnorm_out)
{
    /* one warp per topic: lanes stride over the partials, fixed-order butterfly */
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int z = warp; z < kp; z += nw) {
        double t = 0.0;
        for (int i = lane; i < n_part; i += 32) t += part[(int64_t)i * kp + z];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
        if (lane == 0) {
            colnorm_out[z] = t;
            scale_out[z] = t > 0.0 ? (float)(1.0 / t) : 1.f;
        }
    }
}

/* ---- sharded fit: all-reduce of the raw P(w|z)^T sums over NVLink peer memory, fused with
 * the column sums -------------------------------------------------------------------------
 * Every rank's term pass leaves its partial sums in an exchange buffer that the other GPUs
 * can read (peer access / CUDA IPC).  This kernel, launched behind the term pass on every
 * rank, (1) tells the peers "my partial number `seq` is complete" by writing `seq` into the
 * signal word it owns in each peer's memory, (2) waits until every peer has said the same,
 * (3) reads each row from all ranks (own copy from local HBM, the others over NVLink), adds
 * them in rank order — the same order on every rank, so all ranks hold bit-identical sums —
 * writes the complete row into the local P(w|z)^T, and (4) accumulates the per-topic column
 * sums of the slice in float64 (finished by colsum_final_kernel).  One pass over the data:
 * no staging copy, no second read for the normaliser, 4*kp bytes per row and peer on the
 * wire.  Buffer reuse: a rank rewrites exchange buffer `parity` two iterations later, after
 * the barrier of the iteration in between, which no rank passes before all ranks have left
 * this kernel.  The wait is bounded (option "p2p_timeout_ms", 30 s by default; the ranks enter the
 * loop together behind a one-word all-reduce); a rank that gives up raises *err, and every later
 * launch of the fit returns at once. */
constexpr int SHARD_MAX_RANKS = 16;
struct ShardReduceArgs {
    const float *part[SHARD_MAX_RANKS];      /* every rank's partial, [rows, stride]; [rank] is local */
    unsigned int *peer_sig[SHARD_MAX_RANKS]; /* rank p's signal array (word [rank] is ours to write) */
    volatile unsigned int *my_sig;           /* our signal array: word [p] is written by rank p */
    float *out;                              /* local complete P(w|z)^T [rows, stride] */
    double *colpart;                         /* [grid, kp] */
    int *err;
    int64_t rows;
    int32_t stride, kp, n_ranks, rank;
    unsigned int seq;
    long long timeout_clocks;                /* bound of the wait for the peers' signals */
};

__device__ __forceinline__ float4 ld_peer_f4(const float *p)
{
    float4 v; /* .cv: never served from this SM's L1 (peer lines are cached there only) */
    asm volatile("ld.global.cv.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

/* Signal the peers (CTA 0 only, so a peer sees `seq` once per round) and wait for theirs.
 * false: a wait timed out, now or earlier in this fit (*err): do not wait or add stale sums. */
__device__ __forceinline__ bool shard_handshake(unsigned int *const *peer_sig, volatile unsigned int *my_sig,
                                                int n_ranks, int rank, unsigned int seq,
                                                long long timeout_clocks, int *err, int *give_up)
{
    if (threadIdx.x == 0) *give_up = *reinterpret_cast<volatile int *>(err);
    __syncthreads();
    const bool dead = *give_up != 0; /* one read per CTA: the whole CTA takes the same branch */
    __syncthreads();
    if (dead) return false;
    if (blockIdx.x == 0 && (int)threadIdx.x < n_ranks && (int)threadIdx.x != rank) {
        __threadfence_system();
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(peer_sig[threadIdx.x] + rank), "r"(seq) : "memory");
    }
    if ((int)threadIdx.x < n_ranks && (int)threadIdx.x != rank) {
        const long long t0 = clock64();
        unsigned int v;
        for (;;) {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v)
                         : "l"(const_cast<unsigned int *>(my_sig) + threadIdx.x) : "memory");
            if ((int)(v - seq) >= 0) break;
            if (clock64() - t0 > timeout_clocks) {
                *give_up = 1;
                break;
            }
            __nanosleep(200);
        }
    }
    __syncthreads();
    if (*give_up) {
        if (threadIdx.x == 0) *err = 1;
        return false;
    }
    return true;
}

__global__ void __launch_bounds__(256) shard_reduce_kernel(const ShardReduceArgs a)
{
    __shared__ double sm[256][4];
    __shared__ int give_up;
    /* (1) + (2) */
    if (!shard_handshake(a.peer_sig, a.my_sig, a.n_ranks, a.rank, a.seq, a.timeout_clocks, a.err, &give_up))
        return;
    /* (3) + (4): kp/4 consecutive threads along a row, 256/(kp/4) rows at a time */
    const int nv = a.kp >> 2;
    const int64_t per = (a.rows + gridDim.x - 1) / gridDim.x;
    const int64_t r0 = (int64_t)blockIdx.x * per, r1 = min(a.rows, r0 + per);
    double c0 = 0.0, c1 = 0.0, c2 = 0.0, c3 = 0.0;
    { /* kp <= 1024: nv <= 256 */
        const int nsub = 256 / nv, sub = (int)threadIdx.x / nv, q = (int)threadIdx.x - sub * nv;
        if (sub < nsub) {
            for (int64_t r = r0 + sub; r < r1; r += nsub) {
                const int64_t off = r * a.stride + 4 * q;
                float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int p = 0; p < a.n_ranks; ++p) {
                    const float4 v = (p == a.rank)
                                         ? *reinterpret_cast<const float4 *>(a.part[p] + off)
                                         : ld_peer_f4(a.part[p] + off);
                    t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
                }
                *reinterpret_cast<float4 *>(a.out + off) = t;
                c0 += (double)t.x; c1 += (double)t.y; c2 += (double)t.z; c3 += (double)t.w;
            }
        }
        sm[threadIdx.x][0] = c0; sm[threadIdx.x][1] = c1;
        sm[threadIdx.x][2] = c2; sm[threadIdx.x][3] = c3;
        __syncthreads();
        for (int z = threadIdx.x; z < a.kp; z += 256) { /* topic z lives in sm[sub*nv + z/4][z%4] */
            double acc = 0.0;
            for (int s2 = 0; s2 < nsub; ++s2) acc += sm[s2 * nv + (z >> 2)][z & 3];
            a.colpart[(int64_t)blockIdx.x * a.kp + z] = acc;
        }
    }
}

/* ---- two-shot variant (4 ranks and up) ----------------------------------------------------------
 * The one-shot kernel above reads (G-1) x the whole matrix per rank over NVLink.  Here rank r
 * first adds only ITS slice of the rows (m / G of them) from all ranks, in rank order, and
 * publishes the finished slice in a second exchange buffer; a second kernel then fetches the
 * other ranks' finished slices and takes the column sums of the complete matrix.  2 (G-1)/G of
 * the matrix per rank instead of (G-1): a quarter of the wire traffic at G = 8, the same sums
 * (every element is still the sum of the ranks' partials in rank order), bit-identical on all
 * ranks.  Two signal rounds per iteration: "my partial #seq is complete" (sig) and "my slice
 * #seq is complete" (sig2). */
struct ShardTwoShotArgs {
    const float *part[SHARD_MAX_RANKS];       /* every rank's partial, [rows, stride] */
    float *red[SHARD_MAX_RANKS];              /* every rank's finished slice, [slice_rows, stride]; [rank] is local */
    unsigned int *peer_sig[SHARD_MAX_RANKS];  /* first round: word [rank] of rank p's array is ours */
    unsigned int *peer_sig2[SHARD_MAX_RANKS]; /* second round */
    volatile unsigned int *my_sig, *my_sig2;
    float *out;                               /* local complete P(w|z)^T [rows, stride] */
    double *colpart;                          /* [grid, kp] (second kernel) */
    int *err;
    int64_t rows, slice_rows;                 /* rank p owns rows [p * slice_rows, min(rows, (p+1) * slice_rows)) */
    int32_t stride, kp, n_ranks, rank;
    unsigned int seq;
    long long timeout_clocks;
};

/* shot 1: this rank's slice = sum of all ranks' partials, into `out` and into the exchange buffer */
__global__ void __launch_bounds__(256) shard_slice_reduce_kernel(const ShardTwoShotArgs a)
{
    __shared__ int give_up;
    if (!shard_handshake(a.peer_sig, a.my_sig, a.n_ranks, a.rank, a.seq, a.timeout_clocks, a.err, &give_up))
        return;
    const int nv = a.kp >> 2;
    const int64_t s0 = (int64_t)a.rank * a.slice_rows, s1 = min(a.rows, s0 + a.slice_rows);
    const int64_t per = (max((int64_t)0, s1 - s0) + gridDim.x - 1) / gridDim.x;
    const int64_t r0 = s0 + (int64_t)blockIdx.x * per, r1 = min(s1, r0 + per);
    const int nsub = 256 / nv, sub = (int)threadIdx.x / nv, q = (int)threadIdx.x - sub * nv;
    if (sub >= nsub) return;
    for (int64_t r = r0 + sub; r < r1; r += nsub) {
        const int64_t off = r * a.stride + 4 * q;
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int p = 0; p < a.n_ranks; ++p) {
            const float4 v = (p == a.rank) ? *reinterpret_cast<const float4 *>(a.part[p] + off)
                                           : ld_peer_f4(a.part[p] + off);
            t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
        }
        *reinterpret_cast<float4 *>(a.out + off) = t;
        *reinterpret_cast<float4 *>(a.red[a.rank] + (r - s0) * a.stride + 4 * q) = t;
    }
}

/* shot 2: fetch the other ranks' finished slices, column sums of the whole matrix */
__global__ void __launch_bounds__(256) shard_slice_gather_kernel(const ShardTwoShotArgs a)
{
    __shared__ double sm[256][4];
    __shared__ int give_up;
    if (!shard_handshake(a.peer_sig2, a.my_sig2, a.n_ranks, a.rank, a.seq, a.timeout_clocks, a.err, &give_up))
        return;
    const int nv = a.kp >> 2;
    const int64_t per = (a.rows + gridDim.x - 1) / gridDim.x;
    const int64_t r0 = (int64_t)blockIdx.x * per, r1 = min(a.rows, r0 + per);
    const int nsub = 256 / nv, sub = (int)threadIdx.x / nv, q = (int)threadIdx.x - sub * nv;
    double c0 = 0.0, c1 = 0.0, c2 = 0.0, c3 = 0.0;
    if (sub < nsub) {
        for (int64_t r = r0 + sub; r < r1; r += nsub) {
            const int64_t off = r * a.stride + 4 * q;
            const int owner = (int)(r / a.slice_rows);
            float4 t;
            if (owner == a.rank) {
                t = *reinterpret_cast<const float4 *>(a.out + off);
            } else {
                t = ld_peer_f4(a.red[owner] + (r - (int64_t)owner * a.slice_rows) * a.stride + 4 * q);
                *reinterpret_cast<float4 *>(a.out + off) = t;
            }
            c0 += (double)t.x; c1 += (double)t.y; c2 += (double)t.z; c3 += (double)t.w;
        }
    }
    sm[threadIdx.x][0] = c0; sm[threadIdx.x][1] = c1;
    sm[threadIdx.x][2] = c2; sm[threadIdx.x][3] = c3;
    __syncthreads();
    for (int z = threadIdx.x; z < a.kp; z += 256) {
        double acc = 0.0;
        for (int s2 = 0; s2 < nsub; ++s2) acc += sm[s2 * nv + (z >> 2)][z & 3];
        a.colpart[(int64_t)blockIdx.x * a.kp + z] = acc;
    }
}

/* {log-likelihood, flag} -> two doubles, so that one sum all-reduce carries both */
__global__ void pack_ll_kernel(const double *ll, const int *flag, double *out2)
{
    out2[0] = *ll;
    out2[1] = flag ? (double)*flag : 0.0;
}

/* ---- all-pairs distances between topics (enstop_.py:234-263) -------------------------------
 * Hellinger (umap.distances.hellinger, as called at enstop_.py:253-263):
 *     sqrt(1 - sum_w sqrt(a_w b_w) / sqrt(|a|_1 |b|_1))  =  sqrt(1/2 sum_w (r_w - s_w)^2)
 * with r = sqrt(a / |a|_1), s = sqrt(b / |b|_1): the right-hand form has only non-negative
 * terms, so float32 products with float64 block sums give the distance of two nearly equal
 * topics to full relative accuracy (the 1 - inner product form cancels).
 * KL (enstop_.py:234-250): sum over w with a_w > 0 and b_w > 0 of a_w (log2 a_w - log2 b_w).
 * prep_kernel turns P [N, m] into the per-row operands; pairs_kernel gives one 32 x 32 tile
 * of pairs to a CTA (each thread 2 x 2 pairs), 32 terms at a time through shared memory. */
__global__ void topic_rowsum_kernel(const float *__restrict__ P, int64_t m, double *__restrict__ l1)
{
    __shared__ double sm[256];
    const float *row = P + (int64_t)blockIdx.x * m;
    double t = 0.0;
    for (int64_t w = threadIdx.x; w < m; w += 256) t += (double)row[w];
    sm[threadIdx.x] = t;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if ((int)threadIdx.x < off) sm[threadIdx.x] += sm[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) l1[blockIdx.x] = sm[0];
}

/* kind 0: A = sqrt(P / l1).  kind 1: A = P, B = log2(P) where P > 0, else A = 0 and B = 0
 * (a zero A switches the term off on the a side; the b side is tested in the pair loop) */
__global__ void topic_prep_kernel(const float *__restrict__ P, const double *__restrict__ l1,
                                  int64_t n, int64_t m, int kind, float *__restrict__ A,
                                  float *__restrict__ B)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * m) return;
    const float p = P[i];
    if (kind == 0) {
        const double s = l1[i / m];
        A[i] = s > 0.0 ? (float)sqrt((double)p / s) : 0.f;
    } else {
        A[i] = p > 0.f ? p : 0.f;
        B[i] = p > 0.f ? log2f(p) : 0.f;
    }
}

/* blockIdx.z cuts the terms into gridDim.z slices (a 320 x 320 problem has only 100 pair tiles:
 * without the cut two thirds of the SMs idle); every slice writes raw float64 sums into its own
 * [n, n] plane of `part`, topic_pairs_finish_kernel adds the planes in order. */
template <int KIND>
__global__ void __launch_bounds__(256) topic_pairs_kernel(const float *__restrict__ A,
                                                          const float *__restrict__ B, int n,
                                                          int64_t m, double *__restrict__ part)
{
    constexpr int T = 32, WC = 32;
    __shared__ float ai[T][WC + 1], aj[T][WC + 1], bi[KIND ? T : 1][WC + 1], bj[KIND ? T : 1][WC + 1];
    const int i0 = blockIdx.y * T, j0 = blockIdx.x * T;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4; /* pairs (i0+ty+16a, j0+tx+16b) */
    double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
    const int64_t per = ((m + gridDim.z - 1) / gridDim.z + WC - 1) / WC * WC;
    const int64_t w_lo = (int64_t)blockIdx.z * per, w_hi = min(m, w_lo + per);
    double *out = part + (int64_t)blockIdx.z * n * n;
    for (int64_t w0 = w_lo; w0 < w_hi; w0 += WC) {
        for (int e = threadIdx.x; e < T * WC; e += 256) {
            const int r = e / WC, c = e % WC;
            const int64_t w = w0 + c;
            const bool okw = w < w_hi;
            const int gi = i0 + r, gj = j0 + r;
            ai[r][c] = (okw && gi < n) ? A[(int64_t)gi * m + w] : 0.f;
            aj[r][c] = (okw && gj < n) ? A[(int64_t)gj * m + w] : 0.f;
            if constexpr (KIND == 1) {
                bi[r][c] = (okw && gi < n) ? B[(int64_t)gi * m + w] : 0.f;
                bj[r][c] = (okw && gj < n) ? B[(int64_t)gj * m + w] : 0.f;
            }
        }
        __syncthreads();
        float part[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
#pragma unroll 8
        for (int c = 0; c < WC; ++c) {
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                    const float x = ai[ty + 16 * a][c], y = aj[tx + 16 * b][c];
                    if constexpr (KIND == 0) {
                        const float d = x - y;
                        part[a][b] = fmaf(d, d, part[a][b]);
                    } else { /* a_w (log2 a_w - log2 b_w) where both are positive */
                        const float t = x * (bi[ty + 16 * a][c] - bj[tx + 16 * b][c]);
                        part[a][b] += (y > 0.f) ? t : 0.f;
                    }
                }
        }
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 2; ++b) acc[a][b] += (double)part[a][b];
        __syncthreads();
    }
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            const int gi = i0 + ty + 16 * a, gj = j0 + tx + 16 * b;
            if (gi >= n || gj >= n) continue;
            out[(int64_t)gi * n + gj] = acc[a][b];
        }
}

/* sum of the term slices (fixed order) and, for Hellinger, the final form of the distance */
__global__ void topic_pairs_finish_kernel(const double *__restrict__ part, int n_slices,
                                          const double *__restrict__ l1, int n, int kind,
                                          double *__restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)n * n) return;
    double v = 0.0;
    for (int s = 0; s < n_slices; ++s) v += part[(int64_t)s * n * n + i];
    if (kind == 0) {
        const int gi = (int)(i / n), gj = (int)(i - (int64_t)gi * n);
        const bool zi = !(l1[gi] > 0.0), zj = !(l1[gj] > 0.0);
        v = (gi == gj || (zi && zj)) ? 0.0 : (zi || zj) ? 1.0 : sqrt(0.5 * v);
    }
    out[i] = v;
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
/* separate index / value arrays (the caller's CSR, values of any numeric type) ->
 * interleaved entries with float32 values (plsa.py:714 `.astype(np.float32)`); an index
 * outside [0, n_cols) raises *bad */
template <typename T>
__global__ void interleave_kernel(const int32_t *__restrict__ cols, const T *__restrict__ vals,
                                  int64_t n, int64_t n_cols, int2 *__restrict__ ent, int *bad)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int32_t c = cols[i];
    if ((uint32_t)c >= (uint32_t)n_cols) {
        *bad = 1;
        c = 0;
    }
    ent[i] = make_int2(c, __float_as_int((float)vals[i]));
}

/* per entry: its row (expanded indptr) and its column as a sort key; one warp per row */
__global__ void expand_rows_kernel(const int32_t *__restrict__ indptr, int64_t n_rows,
                                   const int2 *__restrict__ ent, const int32_t *__restrict__ cols,
                                   int32_t *__restrict__ rows_out, int32_t *__restrict__ keys_out)
{
    const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= n_rows) return;
    const int lane = threadIdx.x & 31;
    for (int32_t p = indptr[r] + lane; p < indptr[r + 1]; p += 32) {
        rows_out[p] = (int32_t)r;
        keys_out[p] = cols ? cols[p] : ent[p].x; /* cols: the upload's staging copy of the indices */
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

/* ---- work items, planned on the device (same plan as the host's plan_items) ------------------
 * A row of `span` entries (its stored entries plus the lead-in up to the previous multiple of
 * `align`) is one item, or — longer than `chunk` — nc equal pieces of `per` entries.  Split rows
 * with more than 32 pieces are "heavy" (fixup_kernel gives them a CTA); partial-sum slots are
 * numbered heavy rows first, in row order. */
struct PlanRow { int32_t pieces, per, lead, span; };
__device__ __forceinline__ PlanRow plan_row(const int32_t *indptr, int64_t r, int32_t chunk, int32_t align,
                                            const int32_t *skip_row)
{
    PlanRow p;
    p.lead = indptr[r] & (align - 1);
    p.span = indptr[r + 1] - indptr[r] + p.lead;
    if (skip_row && skip_row[r]) { /* the row is served elsewhere (tiled term pass): no item */
        p.pieces = 0;
        p.per = 0;
        p.span = 0;
    } else if (p.span <= chunk) {
        p.pieces = 1;
        p.per = p.span;
    } else {
        const int32_t nc = (p.span + chunk - 1) / chunk;
        const int32_t eq = (p.span + nc - 1) / nc;
        p.per = min(chunk, (eq + align - 1) / align * align);
        p.pieces = (p.span + p.per - 1) / p.per;
    }
    return p;
}

/* per row: pieces, heavy / light flags and their piece counts (inputs of five exclusive scans) */
__global__ void plan_count_kernel(const int32_t *__restrict__ indptr, int64_t rows, int32_t chunk,
                                  int32_t align, const int32_t *__restrict__ skip_row,
                                  int32_t *__restrict__ pieces,
                                  int32_t *__restrict__ heavy, int32_t *__restrict__ light,
                                  int32_t *__restrict__ heavy_pieces, int32_t *__restrict__ light_pieces)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > rows) return;
    if (r == rows) { /* the scans run over rows + 1 elements: the last one yields the totals */
        pieces[r] = heavy[r] = light[r] = heavy_pieces[r] = light_pieces[r] = 0;
        return;
    }
    const PlanRow p = plan_row(indptr, r, chunk, align, skip_row);
    const bool split = p.span > chunk, hv = split && p.pieces > 32;
    pieces[r] = p.pieces;
    heavy[r] = hv;
    light[r] = split && !hv;
    heavy_pieces[r] = hv ? p.pieces : 0;
    light_pieces[r] = (split && !hv) ? p.pieces : 0;
}

/* the items in row order (pieces of a row adjacent) with their sort keys, and the split-row list */
__global__ void plan_emit_kernel(const int32_t *__restrict__ indptr, int64_t rows, int32_t chunk,
                                 int32_t align, const int32_t *__restrict__ skip_row,
                                 const int32_t *__restrict__ item_at,
                                 const int32_t *__restrict__ heavy_at, const int32_t *__restrict__ light_at,
                                 const int32_t *__restrict__ heavy_slot, const int32_t *__restrict__ light_slot,
                                 Item *__restrict__ items, int32_t *__restrict__ keys,
                                 int32_t *__restrict__ split_rows, int32_t *__restrict__ slot_begin)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const PlanRow p = plan_row(indptr, r, chunk, align, skip_row);
    const int32_t n_heavy = heavy_at[rows], heavy_slots = heavy_slot[rows];
    const bool split = p.span > chunk, hv = split && p.pieces > 32;
    const int64_t s = (int64_t)indptr[r] - p.lead;
    int32_t first_slot = -1;
    if (split) {
        first_slot = hv ? heavy_slot[r] : heavy_slots + light_slot[r];
        const int32_t pos = hv ? heavy_at[r] : n_heavy + light_at[r];
        split_rows[pos] = (int32_t)r;
        slot_begin[pos] = first_slot;
    }
    if (r == rows - 1) /* closing entry: total number of slots */
        slot_begin[n_heavy + light_at[rows]] = heavy_slots + light_slot[rows];
    const int32_t at = item_at[r];
    for (int32_t c = 0; c < p.pieces; ++c) {
        Item it;
        it.start = s + (int64_t)c * p.per;
        it.row = (int32_t)r;
        it.len = min(p.per, p.span - c * p.per);
        it.slot = split ? first_slot + c : -1;
        it.skip = (c == 0) ? (p.lead | (split ? ITEM_FIRST : 0)) : 0;
        items[at + c] = it;
        keys[at + c] = chunk - it.len; /* ascending key = longest first */
    }
}

__global__ void plan_gather_kernel(const Item *__restrict__ in, const int32_t *__restrict__ perm,
                                   int64_t n, Item *__restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[perm[i]];
}

__global__ void fill_kernel(float *p, int64_t n, float v)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

} // namespace plsa
