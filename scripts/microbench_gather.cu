// Microbenchmark: what bounds a k-wide row gather per stored entry on B200?
// Isolates the L1/LSU cost of (a) the gather pattern, (b) shuffles, (c) math, so the row-pass
// lane mapping can be chosen on evidence.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a
//   scripts/microbench_gather.cu -o gpurun_out/microbench_gather ; run on the GPU box.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <random>
#include <algorithm>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

constexpr int STRIDE_B = 128;   // bytes per table row (80 used)

enum Pat { G5_V4 = 0, G4_V4_S, G8_V2_S, G5_V4_TEX, G5_LDS, G2_MIX, G10_V2, G5_TEX_LDG_2_1, G5_TEX_LDG_1_1, G5_TEX_LDG_3_2, G8_TEX };

// every warp walks `per_warp` entries (multiple of 32); idx coalesced, then per step NG entries
template <int PAT, int NSHFL, int NFMA>
__global__ void __launch_bounds__(256, 4) gather_kernel(const int *__restrict__ idx, const char *__restrict__ table,
                                                     cudaTextureObject_t tex, int per_warp, float *out, int hot_rows)
{
    extern __shared__ float4 smem[];
    const int lane = threadIdx.x & 31;
    const long warp_id = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int *my = idx + warp_id * per_warp;
    constexpr int G = (PAT == G4_V4_S) ? 4 : (PAT == G8_V2_S || PAT == G8_TEX) ? 8 : (PAT == G2_MIX) ? 2 : (PAT == G10_V2) ? 10 : 5;
    constexpr int NG = 32 / G;
    constexpr int U = (32 / NG);
    constexpr int CH = NG * U;
    const int grp = lane / G, j = lane % G;
    if (PAT == G5_LDS) {   // stage the hot rows once per CTA (80-byte rows packed)
        for (int i = threadIdx.x; i < hot_rows * 5; i += blockDim.x)
            smem[i] = *reinterpret_cast<const float4 *>(table + (long)(i / 5) * STRIDE_B + (i % 5) * 16);
        __syncthreads();
    }
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
    for (int base = 0; base + CH <= per_warp; base += CH) {
        const int mine = (lane < CH) ? my[base + lane] : 0;
        float4 g[U];
        float s[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const unsigned w = (unsigned)__shfl_sync(0xffffffffu, mine, (u * NG + grp) & 31);
            const char *row = table + (unsigned long long)w * STRIDE_B;
            s[u] = 0.f;
            if (PAT == G5_V4) {
                g[u] = __ldg(reinterpret_cast<const float4 *>(row + j * 16));
            } else if (PAT == G4_V4_S) {
                g[u] = __ldg(reinterpret_cast<const float4 *>(row + j * 16));
                s[u] = __ldg(reinterpret_cast<const float *>(row + 64 + j * 4));
            } else if (PAT == G8_V2_S) {   // 8 lanes x float2 (64 B) + 4 B on lanes 0..3 -> 80 B
                const float2 t = __ldg(reinterpret_cast<const float2 *>(row + j * 8));
                g[u] = make_float4(t.x, t.y, 0.f, 0.f);
                if (j < 4) s[u] = __ldg(reinterpret_cast<const float *>(row + 64 + j * 4));
            } else if (PAT == G10_V2) {    // 10 lanes x float2 = 80 B, 3 entries per step
                const float2 t = __ldg(reinterpret_cast<const float2 *>(row + j * 8));
                g[u] = make_float4(t.x, t.y, 0.f, 0.f);
            } else if (PAT == G2_MIX) {    // 2 lanes x (2 float4 + float2) = 80 B
                const float4 a = __ldg(reinterpret_cast<const float4 *>(row + j * 16));
                const float4 b = __ldg(reinterpret_cast<const float4 *>(row + 32 + j * 16));
                const float2 c = __ldg(reinterpret_cast<const float2 *>(row + 64 + j * 8));
                g[u] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z + c.x, a.w + b.w + c.y);
            } else if (PAT == G5_V4_TEX || PAT == G8_TEX) {
                g[u] = tex1Dfetch<float4>(tex, (int)(w * (STRIDE_B / 16) + j));
            } else if (PAT == G5_TEX_LDG_2_1) {
                if (u % 3 == 2) g[u] = __ldg(reinterpret_cast<const float4 *>(row + j * 16));
                else g[u] = tex1Dfetch<float4>(tex, (int)(w * (STRIDE_B / 16) + j));
            } else if (PAT == G5_TEX_LDG_1_1) {
                if (u % 2 == 1) g[u] = __ldg(reinterpret_cast<const float4 *>(row + j * 16));
                else g[u] = tex1Dfetch<float4>(tex, (int)(w * (STRIDE_B / 16) + j));
            } else if (PAT == G5_TEX_LDG_3_2) {
                if (u == 1 || u == 3) g[u] = __ldg(reinterpret_cast<const float4 *>(row + j * 16));
                else g[u] = tex1Dfetch<float4>(tex, (int)(w * (STRIDE_B / 16) + j));
            } else if (PAT == G5_LDS) {
                g[u] = smem[(w % hot_rows) * 5 + j];
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            float part = (g[u].x + g[u].y) + (g[u].z + g[u].w) + s[u];
#pragma unroll
            for (int t = 0; t < NSHFL; ++t) part += __shfl_xor_sync(0xffffffffu, part, 1 << (t % 5));
#pragma unroll
            for (int t = 0; t < NFMA; ++t) {
                acc0 = fmaf(part, g[u].x, acc0); acc1 = fmaf(part, g[u].y, acc1);
                acc2 = fmaf(part, g[u].z, acc2); acc3 = fmaf(part, g[u].w, acc3);
            }
            acc0 += part;
        }
    }
    if (acc0 + acc1 + acc2 + acc3 == 123.456f) out[0] = acc0;
}

template <int PAT, int NSHFL, int NFMA>
static void run(const char *name, const int *d_idx, const char *d_table, cudaTextureObject_t tex, long nnz,
                float *d_out, int hot_rows)
{
    const int per_warp = 960;                      // multiple of 30 and 32
    const long warps = nnz / per_warp;
    const int threads = 256;
    const long blocks = warps / 8;
    size_t smem = (PAT == G5_LDS) ? (size_t)hot_rows * 80 : 0;
    auto k = gather_kernel<PAT, NSHFL, NFMA>;
    if (smem) CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    for (int w = 0; w < 2; ++w) k<<<(unsigned)blocks, threads, smem>>>(d_idx, d_table, tex, per_warp, d_out, hot_rows);
    CK(cudaEventRecord(a));
    const int reps = 5;
    for (int r = 0; r < reps; ++r) k<<<(unsigned)blocks, threads, smem>>>(d_idx, d_table, tex, per_warp, d_out, hot_rows);
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    ms /= reps;
    const double used = (double)blocks * 8 * per_warp;
    printf("%-34s shfl=%d fma=%d : %8.1f us  %6.3f ns/entry  (%.2f cyc/entry/SM @1.9GHz)\n", name, NSHFL, NFMA,
           ms * 1e3, ms * 1e6 / used, ms * 1e-3 * 1.9e9 * 148 / used);
    CK(cudaGetLastError());
}

int main(int argc, char **argv)
{
    const long nnz = 10'000'000 / 7680 * 7680;
    const int rows = (argc > 1) ? atoi(argv[1]) : 50000;
    const int zipf = (argc > 2) ? atoi(argv[2]) : 1;
    std::vector<int> idx(nnz);
    std::mt19937 rng(1);
    if (zipf) {   // Zipf(1) over rows, like term ids in the doc pass
        std::vector<double> cdf(rows);
        double s = 0; for (int r = 0; r < rows; ++r) { s += 1.0 / (r + 1); cdf[r] = s; }
        std::uniform_real_distribution<double> U(0, s);
        for (long i = 0; i < nnz; ++i) idx[i] = (int)(std::lower_bound(cdf.begin(), cdf.end(), U(rng)) - cdf.begin());
    } else {
        std::uniform_int_distribution<int> U(0, rows - 1);
        for (long i = 0; i < nnz; ++i) idx[i] = U(rng);
    }
    int *d_idx; char *d_table; float *d_out;
    CK(cudaMalloc(&d_idx, nnz * 4)); CK(cudaMalloc(&d_table, (size_t)rows * STRIDE_B)); CK(cudaMalloc(&d_out, 16));
    CK(cudaMemcpy(d_idx, idx.data(), nnz * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(d_table, 0, (size_t)rows * STRIDE_B));
    cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeLinear; rd.res.linear.devPtr = d_table;
    rd.res.linear.desc = cudaCreateChannelDesc<float4>(); rd.res.linear.sizeInBytes = (size_t)rows * STRIDE_B;
    cudaTextureDesc td = {}; td.readMode = cudaReadModeElementType;
    cudaTextureObject_t tex; CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
    printf("rows=%d (%s), %ld entries, 80 B gathered per entry\n", rows, zipf ? "zipf" : "uniform", nnz);
    run<G5_V4, 0, 0>("G5 LDG.128", d_idx, d_table, tex, nnz, d_out, 0);
    run<G5_V4, 1, 0>("G5 LDG.128", d_idx, d_table, tex, nnz, d_out, 0);
    run<G5_V4, 3, 0>("G5 LDG.128", d_idx, d_table, tex, nnz, d_out, 0);
    run<G5_V4, 6, 0>("G5 LDG.128", d_idx, d_table, tex, nnz, d_out, 0);
    run<G5_V4, 0, 2>("G5 LDG.128", d_idx, d_table, tex, nnz, d_out, 0);
    run<G5_V4, 3, 2>("G5 LDG.128", d_idx, d_table, tex, nnz, d_out, 0);
    run<G5_V4, 3, 4>("G5 LDG.128", d_idx, d_table, tex, nnz, d_out, 0);
    run<G4_V4_S, 0, 0>("G4 LDG.128+LDG.32", d_idx, d_table, tex, nnz, d_out, 0);
    run<G4_V4_S, 2, 2>("G4 LDG.128+LDG.32", d_idx, d_table, tex, nnz, d_out, 0);
    run<G8_V2_S, 0, 0>("G8 LDG.64+LDG.32(half)", d_idx, d_table, tex, nnz, d_out, 0);
    run<G10_V2, 0, 0>("G10 LDG.64", d_idx, d_table, tex, nnz, d_out, 0);
    run<G2_MIX, 0, 0>("G2 2xLDG.128+LDG.64", d_idx, d_table, tex, nnz, d_out, 0);
    run<G5_V4_TEX, 0, 0>("G5 tex1Dfetch float4", d_idx, d_table, tex, nnz, d_out, 0);
    run<G5_V4_TEX, 3, 2>("G5 tex1Dfetch float4", d_idx, d_table, tex, nnz, d_out, 0);
    run<G5_V4_TEX, 5, 4>("G5 tex1Dfetch float4", d_idx, d_table, tex, nnz, d_out, 0);
    run<G5_V4_TEX, 8, 4>("G5 tex1Dfetch float4", d_idx, d_table, tex, nnz, d_out, 0);
    run<G8_TEX, 0, 0>("G8 tex1Dfetch float4 (128 B)", d_idx, d_table, tex, nnz, d_out, 0);
    run<G5_TEX_LDG_2_1, 0, 0>("G5 tex:ldg 2:1", d_idx, d_table, tex, nnz, d_out, 0);
    run<G5_TEX_LDG_2_1, 3, 2>("G5 tex:ldg 2:1", d_idx, d_table, tex, nnz, d_out, 0);
    run<G5_TEX_LDG_3_2, 0, 0>("G5 tex:ldg 3:2", d_idx, d_table, tex, nnz, d_out, 0);
    run<G5_TEX_LDG_3_2, 3, 2>("G5 tex:ldg 3:2", d_idx, d_table, tex, nnz, d_out, 0);
    run<G5_TEX_LDG_1_1, 0, 0>("G5 tex:ldg 1:1", d_idx, d_table, tex, nnz, d_out, 0);
    run<G5_TEX_LDG_1_1, 3, 2>("G5 tex:ldg 1:1", d_idx, d_table, tex, nnz, d_out, 0);
    run<G5_LDS, 0, 0>("G5 LDS.128 (rows in smem)", d_idx, d_table, tex, nnz, d_out, 600);
    run<G5_LDS, 3, 2>("G5 LDS.128 (rows in smem)", d_idx, d_table, tex, nnz, d_out, 600);
    return 0;
}
