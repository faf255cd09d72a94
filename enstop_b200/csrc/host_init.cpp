/*
 * host_init.cpp — host-side fast path of the reference's random initialisation.
 *
 * enstop/plsa.py:454-456 draws the initial factors with numpy's legacy RandomState
 * (`rng.rand(k, m)` then `rng.rand(n, k)`), plsa.py:510-511 L1-normalises their rows in
 * float64 (enstop/utils.py:22-41) and plsa.py:709-710 casts them to float32.  A seed must
 * give the same factors here, so the generator cannot be replaced — but it can be run
 * faster: this file advances a RandomState's MT19937 state exactly as numpy does
 * (randomkit/mt19937 `genrand` + `random_sample`: (a >> 5, b >> 6) -> 53-bit double),
 * vectorised with AVX2 where available, and fuses the row normalisation and the float32
 * cast.  The caller (enstop_b200/plsa.py) reads the state with `rng.get_state()` and
 * writes it back with `rng.set_state()`, so the stream stays the caller's.
 *
 * tests/test_host_cpu.py checks bit-equality with numpy for many seeds, positions and
 * shapes, and with the reference's own initial factors stored in tests/golden.
 */
#include <sched.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <stdio.h>
#include <thread>
#include <vector>

#if defined(__AVX2__)
#include <immintrin.h>
#endif

#include "../../include/plsa_b200.h"

#define API extern "C" __attribute__((visibility("default")))

namespace {

constexpr int MT_N = 624, MT_M = 397;
constexpr uint32_t MATRIX_A = 0x9908b0dfu, UPPER = 0x80000000u, LOWER = 0x7fffffffu;

inline uint32_t twist(uint32_t cur, uint32_t nxt, uint32_t far)
{
    const uint32_t y = (cur & UPPER) | (nxt & LOWER);
    return far ^ (y >> 1) ^ ((0u - (y & 1u)) & MATRIX_A);
}

/* regenerate the 624-word block (mt19937_gen) */
void mt_refill(uint32_t *key)
{
    int i = 0;
#if defined(__AVX2__)
    const __m256i upper = _mm256_set1_epi32((int)UPPER), lower = _mm256_set1_epi32((int)LOWER),
                  mat = _mm256_set1_epi32((int)MATRIX_A), one = _mm256_set1_epi32(1),
                  zero = _mm256_setzero_si256();
    auto step = [&](int idx, int far_idx) {
        const __m256i cur = _mm256_loadu_si256((const __m256i *)(key + idx));
        const __m256i nxt = _mm256_loadu_si256((const __m256i *)(key + idx + 1));
        const __m256i far = _mm256_loadu_si256((const __m256i *)(key + far_idx));
        const __m256i y = _mm256_or_si256(_mm256_and_si256(cur, upper), _mm256_and_si256(nxt, lower));
        const __m256i mag = _mm256_and_si256(_mm256_sub_epi32(zero, _mm256_and_si256(y, one)), mat);
        const __m256i r = _mm256_xor_si256(_mm256_xor_si256(far, _mm256_srli_epi32(y, 1)), mag);
        _mm256_storeu_si256((__m256i *)(key + idx), r);
    };
    /* i in [0, 227): reads key[i+1] and key[i+397], both still old when 8 lanes are loaded
     * before the store; 227 = 28*8 + 3 */
    for (; i + 8 <= MT_N - MT_M; i += 8) step(i, i + MT_M);
#endif
    for (; i < MT_N - MT_M; ++i) key[i] = twist(key[i], key[i + 1], key[i + MT_M]);
#if defined(__AVX2__)
    /* i in [227, 623): reads key[i-227] (new, written >= 227 words earlier) and key[i+1] (old) */
    for (; i + 8 <= MT_N - 1; i += 8) step(i, i + (MT_M - MT_N));
#endif
    for (; i < MT_N - 1; ++i) key[i] = twist(key[i], key[i + 1], key[i + (MT_M - MT_N)]);
    key[MT_N - 1] = twist(key[MT_N - 1], key[0], key[MT_M - 1]);
}

inline uint32_t temper(uint32_t y)
{
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

/* The generator state plus a block of tempered outputs, consumed two words per double
 * (numpy legacy random_sample / rk_double: (a >> 5, b >> 6) -> 53 bits). */
struct Stream {
    uint32_t *key;
    int pos;
    uint32_t out[MT_N];

    void temper_from(int from)
    {
        for (int i = from; i < MT_N; ++i) out[i] = temper(key[i]); /* auto-vectorised */
    }
    void refill()
    {
        mt_refill(key);
        temper_from(0);
        pos = 0;
    }
    static inline double make(uint32_t a, uint32_t b)
    {
        return ((a >> 5) * 67108864.0 + (b >> 6)) * (1.0 / 9007199254740992.0);
    }
    void fill(double *dst, int64_t n)
    {
        while (n > 0) {
            if (pos == MT_N) refill();
            int64_t pairs = (MT_N - pos) / 2;
            if (pairs > n) pairs = n;
            const uint32_t *src = out + pos;
            int64_t i = 0;
#if defined(__AVX2__)
            /* four doubles from eight words; a >> 5 and b >> 6 are below 2^27, so the signed
             * conversion is exact, and so are a * 2^26 + b (< 2^53) and the scaling by 2^-53:
             * bit-identical to make() in any evaluation order */
            const __m256i even = _mm256_setr_epi32(0, 2, 4, 6, 0, 2, 4, 6),
                          odd = _mm256_setr_epi32(1, 3, 5, 7, 1, 3, 5, 7);
            const __m256d hi = _mm256_set1_pd(67108864.0), scale = _mm256_set1_pd(1.0 / 9007199254740992.0);
            for (; i + 4 <= pairs; i += 4) {
                const __m256i w = _mm256_loadu_si256((const __m256i *)(src + 2 * i));
                const __m128i a = _mm256_castsi256_si128(
                    _mm256_permutevar8x32_epi32(_mm256_srli_epi32(w, 5), even));
                const __m128i b = _mm256_castsi256_si128(
                    _mm256_permutevar8x32_epi32(_mm256_srli_epi32(w, 6), odd));
                const __m256d v = _mm256_add_pd(_mm256_mul_pd(_mm256_cvtepi32_pd(a), hi),
                                                _mm256_cvtepi32_pd(b));
                _mm256_storeu_pd(dst + i, _mm256_mul_pd(v, scale));
            }
#endif
            for (; i < pairs; ++i) dst[i] = make(src[2 * i], src[2 * i + 1]);
            dst += pairs;
            n -= pairs;
            pos += (int)(2 * pairs);
            if (n > 0 && pos == MT_N - 1) { /* a double straddling two blocks */
                const uint32_t a = out[pos];
                refill();
                *dst++ = make(a, out[0]);
                pos = 1;
                --n;
            }
        }
    }
};

} // namespace

/* The factors go to the GPU next (plsa_set_factors, a DMA read of page-locked memory): written
 * with ordinary stores by several cores they sit dirty in those cores' caches and the DMA engine
 * snoops them out line by line (measured: 8 MB at 6 GB/s instead of 45).  Streaming stores put
 * them in memory. */
static inline void store_streaming(float *dst, const float *src, int64_t n)
{
#if defined(__AVX2__)
    auto one = [&](int64_t c) {
        int bits;
        memcpy(&bits, src + c, 4);
        _mm_stream_si32(reinterpret_cast<int *>(dst + c), bits);
    };
    int64_t c = 0;
    for (; c < n && ((uintptr_t)(dst + c) & 15u); ++c) one(c);
    for (; c + 4 <= n; c += 4) _mm_stream_ps(dst + c, _mm_loadu_ps(src + c));
    for (; c < n; ++c) one(c);
#else
    memcpy(dst, src, sizeof(float) * (size_t)n);
#endif
}

static inline void store_fence()
{
#if defined(__AVX2__)
    _mm_sfence();
#endif
}

/* R rows at a time.  The left-to-right marginal of a row is one chain of dependent additions
 * (4 cycles each: more than everything else done per value).  Long rows (P(w|z)) are drawn four
 * at a time and their four chains advance together; every row's sum is still taken in the
 * reference's order.  Short rows (R = 1) overlap by themselves in the out-of-order window. */
template <int R>
static int draw_rows_by(uint32_t *key, int32_t *pos, int64_t r0, int64_t r1, int64_t cols, float *out,
                        double *out_f64)
{
    Stream s;
    s.key = key;
    s.pos = *pos;
    s.temper_from(s.pos < MT_N ? s.pos : MT_N);
    const size_t w = (size_t)(cols > 0 ? cols : 1);
    double *buf = (double *)malloc(sizeof(double) * R * w + sizeof(float) * w);
    if (!buf) return PLSA_ENOMEM;
    float *row32 = reinterpret_cast<float *>(buf + R * w);
    for (int64_t r = r0; r < r1; r += R) {
        const int nr = (int)std::min<int64_t>(R, r1 - r);
        s.fill(buf, (int64_t)nr * cols); /* rng.rand(rows, cols) is row-major: nr whole rows */
        double marginal[R];
        if (R == 4 && nr == 4) {
            const double *b0 = buf, *b1 = buf + cols, *b2 = buf + 2 * cols, *b3 = buf + 3 * cols;
            double m0 = 0.0, m1 = 0.0, m2 = 0.0, m3 = 0.0;
            for (int64_t c = 0; c < cols; ++c) { /* left to right, as utils.py:25-29 */
                m0 += b0[c];
                m1 += b1[c];
                m2 += b2[c];
                m3 += b3[c];
            }
            marginal[0] = m0; marginal[1 % R] = m1; marginal[2 % R] = m2; marginal[3 % R] = m3;
        } else {
            for (int j = 0; j < nr; ++j) {
                double m = 0.0;
                for (int64_t c = 0; c < cols; ++c) m += buf[j * cols + c];
                marginal[j] = m;
            }
        }
        for (int j = 0; j < nr; ++j) {
            const double *src = buf + j * cols;
            float *o = out + (r + j) * cols;
            double *o64 = out_f64 ? out_f64 + (r + j) * cols : nullptr;
            const double m = marginal[j];
            if (m > 0.0) {
                for (int64_t c = 0; c < cols; ++c) {
                    const double v = src[c] / m;
                    row32[c] = (float)v;
                    if (o64) o64[c] = v;
                }
            } else {
                for (int64_t c = 0; c < cols; ++c) {
                    row32[c] = (float)src[c];
                    if (o64) o64[c] = src[c];
                }
            }
            store_streaming(o, row32, cols);
        }
    }
    store_fence();
    free(buf);
    *pos = s.pos;
    return PLSA_OK;
}

/* rows [r0, r1) of the draw from a generator state positioned at the first word of row r0 */
static int draw_rows(uint32_t *key, int32_t *pos, int64_t r0, int64_t r1, int64_t cols, float *out,
                     double *out_f64)
{
    return cols >= 256 ? draw_rows_by<4>(key, pos, r0, r1, cols, out, out_f64)
                       : draw_rows_by<1>(key, pos, r0, r1, cols, out, out_f64);
}

/* advance the state by `words` outputs without producing them (twist only: the tempering, the
 * conversion to doubles and the normalisation are four fifths of a draw's cost) */
static void skip_words(uint32_t *key, int32_t *pos, int64_t words)
{
    int p = *pos;
    while (words > 0) {
        if (p == MT_N) {
            mt_refill(key);
            p = 0;
        }
        const int64_t take = words < (int64_t)(MT_N - p) ? words : (int64_t)(MT_N - p);
        p += (int)take;
        words -= take;
    }
    *pos = p;
}

/* worker threads of one draw: the cores this process may use, shared with the other ranks of
 * the box (LOCAL_WORLD_SIZE); the upload thread beside the draw mostly waits for DMA */
static int init_threads()
{
    static const int n = [] {
        if (const char *e = getenv("ENSTOP_B200_INIT_THREADS")) return std::max(1, std::min(8, atoi(e)));
        cpu_set_t set;
        int cores = 0;
        if (sched_getaffinity(0, sizeof(set), &set) == 0) cores = CPU_COUNT(&set);
        if (cores <= 0) cores = (int)std::thread::hardware_concurrency();
        int ranks = 1;
        if (const char *e = getenv("LOCAL_WORLD_SIZE")) ranks = std::max(1, atoi(e));
        return std::max(1, std::min(8, cores / ranks));
    }();
    return n;
}

/* Draw rows*cols doubles (row-major, the order rng.rand(rows, cols) produces them),
 * L1-normalise every row with a float64 marginal accumulated left to right
 * (enstop/utils.py:22-41; rows whose marginal is not > 0 are left as drawn) and store
 * float32 (plsa.py:709-710).  key[624] / *pos are the RandomState's MT19937 state, updated
 * in place.  out_f64 (optional) receives the normalised float64 values.
 *
 * The MT19937 stream is sequential, but advancing it is cheap next to consuming it: the rows
 * are cut into one slice per worker thread, this thread skips the state ahead to the start of
 * each slice (a snapshot per slice) and the workers produce their slices in parallel — the
 * same words land in the same places as in a serial draw. */
API int plsa_host_random_rows(uint32_t *key, int32_t *pos, int64_t rows, int64_t cols, float *out,
                              double *out_f64)
{
    if (!key || !pos || !out || rows < 0 || cols < 0 || *pos < 0 || *pos > MT_N) return PLSA_EINVAL;
    const int64_t total = rows * cols;
    int T = init_threads();
    if (total < (int64_t)1 << 18) T = 1;
    if (rows < 2 * T) T = (int)std::max<int64_t>(1, rows / 2); /* few, long rows (P(w|z)): fewer slices */
    if (T == 1) return draw_rows(key, pos, 0, rows, cols, out, out_f64);
    /* Worker t can start once this thread has skipped the slices before it, so equal slices
     * would finish one after the other.  A slice costs ~7.5x more to draw than to skip (2.25 against
     * 0.3 ns per word, scripts/time_init.py): slices shrinking by 1/8 per worker finish together.  The last slice is drawn by this
     * thread on the caller's state, which is then where a serial draw would have left it. */
    struct Slice { uint32_t key[MT_N]; int32_t pos; int64_t r0, r1; int rc; double t0, t1; };
    std::vector<Slice> slices((size_t)(T - 1));
    static const bool profile = getenv("ENSTOP_B200_INIT_PROFILE") != nullptr; /* slice time line on stderr */
    const double ratio = 7.0 / 8.0;
    const auto clk0 = std::chrono::steady_clock::now();
    auto now_ms = [clk0]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - clk0).count(); };
    std::vector<double> skip_done((size_t)T, 0.0);
    std::vector<int64_t> cut((size_t)T + 1, 0);
    {
        double w = 1.0, sum = 0.0;
        std::vector<double> share((size_t)T);
        for (int t = 0; t < T; ++t, w *= ratio) sum += (share[(size_t)t] = w);
        double acc = 0.0;
        for (int t = 0; t < T; ++t) {
            acc += share[(size_t)t];
            cut[(size_t)t + 1] = std::min<int64_t>(rows, (int64_t)((double)rows * acc / sum + 0.5));
        }
        cut[(size_t)T] = rows;
    }
    std::vector<std::thread> workers;
    try {
        for (int t = 0; t + 1 < T; ++t) {
            Slice &sl = slices[(size_t)t];
            sl.r0 = cut[(size_t)t];
            sl.r1 = cut[(size_t)t + 1];
            sl.rc = PLSA_OK;
            memcpy(sl.key, key, sizeof(sl.key));
            sl.pos = *pos;
            workers.emplace_back([&sl, cols, out, out_f64, now_ms]() {
                sl.t0 = now_ms();
                sl.rc = draw_rows(sl.key, &sl.pos, sl.r0, sl.r1, cols, out, out_f64);
                sl.t1 = now_ms();
            });
            skip_words(key, pos, 2 * (sl.r1 - sl.r0) * cols); /* the caller's state: behind this slice */
            skip_done[(size_t)t] = now_ms();
        }
    } catch (...) { /* thread creation failed: finish what was started, report */
        for (auto &w : workers) w.join();
        return PLSA_ENOMEM;
    }
    const double t_last0 = now_ms();
    const int rc_last = draw_rows(key, pos, cut[(size_t)T - 1], rows, cols, out, out_f64);
    const double t_last1 = now_ms();
    for (auto &w : workers) w.join();
    if (profile) {
        for (int t = 0; t + 1 < T; ++t)
            fprintf(stderr, "[init] slice %d rows %lld: thread %.3f..%.3f ms, skipped by %.3f\n", t,
                    (long long)(slices[(size_t)t].r1 - slices[(size_t)t].r0), slices[(size_t)t].t0,
                    slices[(size_t)t].t1, skip_done[(size_t)t]);
        fprintf(stderr, "[init] last slice rows %lld: %.3f..%.3f ms, joined %.3f\n",
                (long long)(rows - cut[(size_t)T - 1]), t_last0, t_last1, now_ms());
    }
    for (const Slice &sl : slices)
        if (sl.rc != PLSA_OK) return sl.rc;
    return rc_last;
}
