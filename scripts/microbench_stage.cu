// Microbenchmark (round 2): how should the k-wide factor row of every stored entry reach the
// registers of the lane that multiplies it?  Rows are 20 floats (k = 20, 80 bytes).
//
//   tile   : the hot rows live in a shared-memory tile (one TMA bulk copy per CTA); consumer is
//            LANE-PER-ENTRY: an octet (8 lanes) walks one work item, each lane reads the whole
//            80-byte row of ITS entry with 5 LDS.128 and keeps the posterior normaliser in-lane
//            (no shuffle per entry; the k accumulators are folded across the octet once per item).
//            "random"  : entries in CSR order (bank group of a row = 5*row mod 8, random)
//            "ordered" : entries of an item permuted so that the 8 lanes of an octet hit 8
//                        different 16-byte bank groups (row index mod 8 distinct) where possible
//   ldgsts : rows come from global/L2 through cp.async (LDGSTS, 5 lanes x 16 B per row) into a
//            per-warp ring of contiguous 80-byte slots, same lane-per-entry consumer
//   bulk   : rows come through the TMA unit, one cp.async.bulk of 80 bytes per entry completing
//            on a per-warp mbarrier (complete_tx), same consumer
//   red    : M-step scatter alternative: red.global.add.v4.f32 of an 80-byte row per entry
//
// Every variant does the real arithmetic of the row pass per entry (20 flush-to-zero products,
// in-lane sum, reciprocal, 20 FMAs).  Output: microseconds and cycles per entry per SM.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo scripts/microbench_stage.cu -o build/microbench_stage
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <random>
#include <algorithm>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1);} } while (0)

constexpr int K = 20, KC = 5;            // floats per row, float4 chunks per row
constexpr int ROW_B = 80;                // compact pitch
constexpr int THREADS = 768;             // 24 warps, one CTA per SM
constexpr int ITEM_LEN = 64;             // entries per work item (multiple of 8)

typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk2(f32x2 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 mul2_ftz(f32x2 a, f32x2 b) { f32x2 r; asm("mul.rn.ftz.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float rcp_fast(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cp_async16(void *dst, const void *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// the arithmetic of one entry, lane-local: row g (5 float4), owned row own2 (carries the
// flush-to-zero scale), value x; accumulates x * v / sum(v) into acc2
__device__ __forceinline__ void entry_math(const float4 (&g)[KC], const f32x2 (&own2)[2 * KC], float x,
                                           f32x2 (&acc2)[2 * KC])
{
    f32x2 v[2 * KC];
#pragma unroll
    for (int c = 0; c < KC; ++c) {
        v[2 * c] = mul2_ftz(pk2(g[c].x, g[c].y), own2[2 * c]);
        v[2 * c + 1] = mul2_ftz(pk2(g[c].z, g[c].w), own2[2 * c + 1]);
    }
    f32x2 s01 = add2(v[0], v[1]), s23 = add2(v[2], v[3]), s45 = add2(v[4], v[5]), s67 = add2(v[6], v[7]),
          s89 = add2(v[8], v[9]);
    f32x2 s = add2(add2(add2(s01, s23), add2(s45, s67)), s89);
    float lo, hi;
    upk2(s, lo, hi);
    const float c1 = fminf(x * rcp_fast(lo + hi), 3.0e38f);
    const f32x2 c2 = pk2(c1, c1);
#pragma unroll
    for (int i = 0; i < 2 * KC; ++i) acc2[i] = fma2(c2, v[i], acc2[i]);
}

// fold the 20 accumulators over the 8 lanes of an octet (halving transpose-reduce) and store
__device__ __forceinline__ void octet_fold_store(f32x2 (&acc2)[2 * KC], float *dst, int li)
{
    float a[K];
#pragma unroll
    for (int i = 0; i < 2 * KC; ++i) upk2(acc2[i], a[2 * i], a[2 * i + 1]);
    // plain butterfly over 3 levels (the transposed variant saves shuffles; this is per item)
#pragma unroll
    for (int off = 4; off > 0; off >>= 1)
#pragma unroll
        for (int i = 0; i < K; ++i) a[i] += __shfl_xor_sync(0xffffffffu, a[i], off);
    if (li == 0) {
#pragma unroll
        for (int c = 0; c < KC; ++c) reinterpret_cast<float4 *>(dst)[c] = make_float4(a[4 * c], a[4 * c + 1], a[4 * c + 2], a[4 * c + 3]);
    }
}

struct Args {
    const int2 *ent;      // [n_items * ITEM_LEN] {row index, value bits}
    const float *own;     // [n_items, K]
    const float *table;   // gathered rows
    float *out;           // [n_items, K]
    long n_items;
    int tile_rows;        // tile variant: rows staged per CTA
    int pitch_b;          // bytes between table rows (ldgsts / bulk variants)
};

// ---- tile: rows resident in shared memory -----------------------------------------------------
__global__ void __launch_bounds__(THREADS, 1) tile_kernel(const Args a)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    float4 *tile = reinterpret_cast<float4 *>(smem);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, li = lane & 7, oct = lane >> 3;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t bytes = (uint32_t)a.tile_rows * ROW_B;
        mbar_expect_tx(&bar, bytes);
        uint32_t off = 0;
        while (off < bytes) {                      // TMA bulk copies of up to 32 KB
            const uint32_t n = min(bytes - off, 32768u);
            bulk_g2s(smem + off, reinterpret_cast<const char *>(a.table) + off, n, &bar);
            off += n;
        }
    }
    mbar_wait(&bar, 0);
    const long per = (a.n_items + gridDim.x - 1) / gridDim.x;
    const long i0 = blockIdx.x * per, i1 = min(a.n_items, i0 + per);
    const int nwarp = THREADS / 32;
    for (long b = i0 + warp * 4; b < i1; b += nwarp * 4) {
        const long item = min(b + oct, i1 - 1);
        f32x2 own2[2 * KC], acc2[2 * KC];
#pragma unroll
        for (int c = 0; c < KC; ++c) {
            const float4 o = __ldg(reinterpret_cast<const float4 *>(a.own + item * K) + c);
            own2[2 * c] = pk2(o.x, o.y); own2[2 * c + 1] = pk2(o.z, o.w);
            acc2[2 * c] = 0ull; acc2[2 * c + 1] = 0ull;
        }
        const int2 *ent = a.ent + item * ITEM_LEN + li;
        int2 e = __ldg(ent);
        for (int t = 0; t < ITEM_LEN; t += 8) {
            const int2 en = __ldg(ent + t + 8);    // one entry block ahead (padding follows the array)
            float4 g[KC];
#pragma unroll
            for (int c = 0; c < KC; ++c) g[c] = tile[e.x * KC + c];
            entry_math(g, own2, __int_as_float(e.y), acc2);
            e = en;
        }
        octet_fold_store(acc2, a.out + item * K, li);
    }
}

// ---- ldgsts / bulk: rows staged from L2 into a per-warp ring ---------------------------------
constexpr int STAGES = 3;
constexpr int RING_THREADS = 512;          // 16 warps: 128 registers, no spills
template <bool BULK>
__global__ void __launch_bounds__(RING_THREADS, 1) ring_kernel(const Args a)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bars[RING_THREADS / 32][STAGES];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, li = lane & 7, oct = lane >> 3;
    unsigned char *ring = smem + (size_t)warp * STAGES * 32 * ROW_B;
    if (BULK && lane == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) mbar_init(&bars[warp][s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const char *table = reinterpret_cast<const char *>(a.table);
    const long per = (a.n_items + gridDim.x - 1) / gridDim.x;
    const long i0 = blockIdx.x * per, i1 = min(a.n_items, i0 + per);
    const int nwarp = RING_THREADS / 32;
    uint32_t phase_bits = 0;   // parity of each stage's barrier
    for (long b = i0 + warp * 4; b < i1; b += nwarp * 4) {
        const long item = min(b + oct, i1 - 1);
        f32x2 own2[2 * KC], acc2[2 * KC];
#pragma unroll
        for (int c = 0; c < KC; ++c) {
            const float4 o = __ldg(reinterpret_cast<const float4 *>(a.own + item * K) + c);
            own2[2 * c] = pk2(o.x, o.y); own2[2 * c + 1] = pk2(o.z, o.w);
            acc2[2 * c] = 0ull; acc2[2 * c + 1] = 0ull;
        }
        const int2 *ent = a.ent + item * ITEM_LEN + li;
        constexpr int NIT = ITEM_LEN / 8;
        // issue the row copies of iteration `it` into stage it % STAGES
        auto issue = [&](int it, int2 e) {
            unsigned char *st = ring + (size_t)(it % STAGES) * 32 * ROW_B;
            if constexpr (BULK) {
                uint64_t *bar = &bars[warp][it % STAGES];
                if (lane == 0) mbar_expect_tx(bar, 32 * ROW_B);
                __syncwarp();
                bulk_g2s(st + lane * ROW_B, table + (size_t)(uint32_t)e.x * a.pitch_b, ROW_B, bar);
            } else {
#pragma unroll
                for (int i = 0; i < KC; ++i) {     // 160 16-byte pieces over 5 instructions
                    const int q = i * 32 + lane, slot = q / KC, ch = q - slot * KC;
                    const uint32_t w = (uint32_t)__shfl_sync(0xffffffffu, e.x, slot);
                    cp_async16(st + slot * ROW_B + ch * 16, table + (size_t)w * a.pitch_b + ch * 16);
                }
                cp_async_commit();
            }
        };
        int2 eq[STAGES];   // entries of the iterations in flight
#pragma unroll
        for (int s = 0; s < STAGES - 1; ++s) {
            eq[s] = __ldg(ent + s * 8);
            issue(s, eq[s]);
        }
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            if (it + STAGES - 1 < NIT) {
                eq[(it + STAGES - 1) % STAGES] = __ldg(ent + (it + STAGES - 1) * 8);
                issue(it + STAGES - 1, eq[(it + STAGES - 1) % STAGES]);
            } else if (!BULK) {
                cp_async_commit();                 // empty group keeps the wait count uniform
            }
            if constexpr (BULK) {
                mbar_wait(&bars[warp][it % STAGES], (phase_bits >> (it % STAGES)) & 1u);
                phase_bits ^= 1u << (it % STAGES);
            } else {
                cp_async_wait<STAGES - 1>();
                __syncwarp();
            }
            const float4 *row = reinterpret_cast<const float4 *>(ring + (size_t)(it % STAGES) * 32 * ROW_B + lane * ROW_B);
            float4 g[KC];
#pragma unroll
            for (int c = 0; c < KC; ++c) g[c] = row[c];
            entry_math(g, own2, __int_as_float(eq[it % STAGES].y), acc2);
            __syncwarp();                          // the stage is rewritten by the next issue
        }
        if (!BULK) cp_async_wait<0>();
        octet_fold_store(acc2, a.out + item * K, li);
    }
}

// ---- red: M-step scatter with vector reductions ------------------------------------------------
__global__ void __launch_bounds__(256, 4) red_kernel(const int2 *__restrict__ ent, long n, float *table, int pitch_f)
{
    const int lane = threadIdx.x & 31;
    const long warp_id = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long n_warps = ((long)gridDim.x * blockDim.x) >> 5;
    const int grp = min(lane / KC, 5), j = lane - (lane / KC) * KC;
    for (long base = warp_id * 6; base < n; base += n_warps * 6) {
        const long i = base + grp;
        if (lane < 30 && i < n) {
            const int2 e = __ldg(ent + i);
            const float x = __int_as_float(e.y);
            float *p = table + (size_t)(uint32_t)e.x * pitch_f + 4 * j;
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(x), "f"(x * 0.5f), "f"(x * 0.25f), "f"(x * 0.125f) : "memory");
        }
    }
}

static std::vector<int> zipf_indices(long n, int rows, std::mt19937 &rng)
{
    std::vector<double> cdf(rows);
    double s = 0;
    for (int r = 0; r < rows; ++r) { s += 1.0 / (r + 1); cdf[r] = s; }
    std::uniform_real_distribution<double> U(0, s);
    std::vector<int> idx(n);
    for (long i = 0; i < n; ++i) idx[i] = (int)(std::lower_bound(cdf.begin(), cdf.end(), U(rng)) - cdf.begin());
    return idx;
}

// permute the entries of every item so that aligned octets hold distinct (index mod 8) where
// possible: order by (rank of the entry inside its residue class, class)
static void bank_order(std::vector<int> &idx, long n_items)
{
    std::vector<int> tmp(ITEM_LEN);
    double conflicts = 0;
    for (long it = 0; it < n_items; ++it) {
        int *p = idx.data() + it * ITEM_LEN;
        int cnt[8] = {0};
        std::vector<std::pair<int, int>> key(ITEM_LEN);
        for (int i = 0; i < ITEM_LEN; ++i) { const int c = p[i] & 7; key[i] = {cnt[c]++ * 8 + c, p[i]}; }
        std::sort(key.begin(), key.end());
        for (int i = 0; i < ITEM_LEN; ++i) p[i] = key[i].second;
        for (int o = 0; o < ITEM_LEN; o += 8) {
            int c8[8] = {0}, mx = 0;
            for (int i = 0; i < 8; ++i) mx = std::max(mx, ++c8[p[o + i] & 7]);
            conflicts += mx;
        }
    }
    printf("  bank order: mean conflict degree per octet %.3f\n", conflicts / (n_items * (ITEM_LEN / 8)));
}

template <class F> static float time_it(F f, int reps = 5)
{
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    f(); f();
    CK(cudaEventRecord(a));
    for (int r = 0; r < reps; ++r) f();
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    CK(cudaGetLastError());
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    return ms / reps;
}

static void report(const char *name, float ms, double entries, int sms, double mhz)
{
    printf("%-46s : %8.1f us  %.3f cyc/entry/SM @%.0f MHz\n", name, ms * 1e3, ms * 1e-3 * mhz * 1e6 * sms / entries, mhz);
}

int main()
{
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    int khz = 0;
    CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
    const double mhz = khz / 1000.0;
    const long n_items = 10'000'000 / ITEM_LEN;
    const long nnz = n_items * ITEM_LEN;
    std::mt19937 rng(1);
    printf("%d SMs, %.0f MHz, %ld items x %d entries, rows of %d floats\n", sms, mhz, n_items, ITEM_LEN, K);

    float *d_own, *d_out, *d_table;
    int2 *d_ent;
    const int big_rows = 100000;
    CK(cudaMalloc(&d_own, n_items * K * 4)); CK(cudaMalloc(&d_out, n_items * K * 4));
    CK(cudaMalloc(&d_table, (size_t)big_rows * 128)); CK(cudaMalloc(&d_ent, (nnz + 4096) * 8));
    CK(cudaMemset(d_ent, 0, (nnz + 4096) * 8));
    {
        std::vector<float> h(n_items * K, 1e-3f), t((size_t)big_rows * 32, 0.01f);
        CK(cudaMemcpy(d_own, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(d_table, t.data(), t.size() * 4, cudaMemcpyHostToDevice));
    }
    auto upload = [&](const std::vector<int> &idx) {
        std::vector<int2> e(nnz);
        for (long i = 0; i < nnz; ++i) e[i] = make_int2(idx[i], __builtin_bit_cast(int, 1.0f + (float)(i & 3)));
        CK(cudaMemcpy(d_ent, e.data(), nnz * 8, cudaMemcpyHostToDevice));
    };
    Args a{};
    a.ent = d_ent; a.own = d_own; a.table = d_table; a.out = d_out; a.n_items = n_items;

    // ---- tile variants: Zipf over the tile's rows (the head of the vocabulary)
    for (int tile_rows : {2048, 2816}) {
        const size_t smem = (size_t)tile_rows * ROW_B;
        CK(cudaFuncSetAttribute(tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        a.tile_rows = tile_rows;
        std::vector<int> idx = zipf_indices(nnz, tile_rows, rng);
        upload(idx);
        char name[96];
        snprintf(name, sizeof name, "tile %d rows, lane-per-entry, random order", tile_rows);
        report(name, time_it([&] { tile_kernel<<<sms, THREADS, smem>>>(a); }), (double)nnz, sms, mhz);
        bank_order(idx, n_items);
        upload(idx);
        snprintf(name, sizeof name, "tile %d rows, lane-per-entry, bank-ordered", tile_rows);
        report(name, time_it([&] { tile_kernel<<<sms, THREADS, smem>>>(a); }), (double)nnz, sms, mhz);
    }
    // ---- ring variants: rows from L2 (50k rows Zipf like the doc pass, 100k uniform like the term pass)
    const size_t ring_smem = (size_t)(RING_THREADS / 32) * STAGES * 32 * ROW_B;
    CK(cudaFuncSetAttribute(ring_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ring_smem));
    CK(cudaFuncSetAttribute(ring_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ring_smem));
    for (int mode = 0; mode < 2; ++mode) {
        const int rows = mode ? 100000 : 50000;
        std::vector<int> idx;
        if (mode) { idx.resize(nnz); std::uniform_int_distribution<int> U(0, rows - 1); for (auto &v : idx) v = U(rng); }
        else idx = zipf_indices(nnz, rows, rng);
        upload(idx);
        for (int pitch : {128, 80}) {
            a.pitch_b = pitch;
            char name[96];
            snprintf(name, sizeof name, "ring LDGSTS, %s %dk rows, pitch %d", mode ? "uniform" : "zipf", rows / 1000, pitch);
            report(name, time_it([&] { ring_kernel<false><<<sms, RING_THREADS, ring_smem>>>(a); }), (double)nnz, sms, mhz);
            snprintf(name, sizeof name, "ring cp.async.bulk 80B/entry, %s %dk rows, pitch %d", mode ? "uniform" : "zipf", rows / 1000, pitch);
            report(name, time_it([&] { ring_kernel<true><<<sms, RING_THREADS, ring_smem>>>(a); }), (double)nnz, sms, mhz);
        }
        char name[96];
        snprintf(name, sizeof name, "red.global.add.v4.f32 scatter, %s %dk rows", mode ? "uniform" : "zipf", rows / 1000);
        report(name, time_it([&] { red_kernel<<<sms * 8, 256>>>(d_ent, nnz, d_table, 32); }), (double)nnz, sms, mhz);
    }
    return 0;
}
