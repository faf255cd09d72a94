/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the shipped product path.
 *
 * CPU restatement of the pLSA EM arithmetic of lmcinnes/enstop (enstop/plsa.py).
 * This header is included twice by plsa_oracle.c:
 *   REAL = float   SUFFIX = f32  -> faithful restatement (float32 storage and float32
 *                                   accumulators, same loop order as the reference's
 *                                   numba kernels, serial M-step scatter)
 *   REAL = double  SUFFIX = f64  -> the same algorithm carried out in float64 end to end
 *                                   (the "exact" yardstick of SURVEY.md §7 hard part 1)
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  Every function cites the reference file:line it follows.
 */

#define CAT_(a, b) a##_##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUFFIX)

/* plsa.py:89-105  plsa_e_step — prange over nnz; thresholded product, then normalise. */
static void FN(e_step)(const int32_t *rows, const int32_t *cols, int64_t nnz,
                       const REAL *pwz /* [k,m] */, const REAL *pzd /* [n,k] */,
                       REAL *pzwd /* [nnz,k] */, int64_t k, int64_t m, REAL thresh)
{
#pragma omp parallel for schedule(static)
    for (int64_t nz = 0; nz < nnz; ++nz) {
        const int64_t d = rows[nz], w = cols[nz];
        REAL norm = 0;
        REAL *post = pzwd + nz * k;
        for (int64_t z = 0; z < k; ++z) {
            REAL v = pwz[z * m + w] * pzd[d * k + z];
            if (v > thresh) {
                post[z] = v;
                norm += post[z];
            } else {
                post[z] = 0;
            }
        }
        for (int64_t z = 0; z < k; ++z)
            if (norm > 0) post[z] /= norm;
    }
}

/* plsa.py:172-202  plsa_m_step   and   plsa.py:277-308  plsa_m_step_w_sample_weight.
 * The scatter loop over nnz is SERIAL in the reference (plain `range`, plsa.py:182/287);
 * float32 accumulators norm_pwz / norm_pdz are kept exactly as there.  `sw == NULL`
 * selects the unweighted variant; otherwise P(w|z) and norm_pwz receive s*sw[d] while
 * P(z|d) and norm_pdz receive s (plsa.py:293-300). */
static void FN(m_step)(const int32_t *rows, const int32_t *cols, const REAL *vals,
                       int64_t nnz, REAL *pwz, REAL *pzd, const REAL *pzwd,
                       const REAL *sw, REAL *norm_pwz, REAL *norm_pdz, int64_t n,
                       int64_t m, int64_t k)
{
    memset(pwz, 0, sizeof(REAL) * (size_t)(k * m));
    memset(pzd, 0, sizeof(REAL) * (size_t)(n * k));
    memset(norm_pwz, 0, sizeof(REAL) * (size_t)k);
    memset(norm_pdz, 0, sizeof(REAL) * (size_t)n);

    for (int64_t nz = 0; nz < nnz; ++nz) {
        const int64_t d = rows[nz], w = cols[nz];
        const REAL x = vals[nz];
        const REAL *post = pzwd + nz * k;
        if (sw) {
            const REAL wd = sw[d];
            for (int64_t z = 0; z < k; ++z) {
                REAL s = x * post[z];
                REAL t = s * wd;
                pwz[z * m + w] += t;
                pzd[d * k + z] += s;
                norm_pwz[z] += t;
                norm_pdz[d] += s;
            }
        } else {
            for (int64_t z = 0; z < k; ++z) {
                REAL s = x * post[z];
                pwz[z * m + w] += s;
                pzd[d * k + z] += s;
                norm_pwz[z] += s;
                norm_pdz[d] += s;
            }
        }
    }

    /* plsa.py:196-202 — prange over k */
#pragma omp parallel for schedule(static)
    for (int64_t z = 0; z < k; ++z) {
        if (norm_pwz[z] > 0)
            for (int64_t w = 0; w < m; ++w) pwz[z * m + w] /= norm_pwz[z];
        for (int64_t d = 0; d < n; ++d)
            if (norm_pdz[d] > 0) pzd[d * k + z] /= norm_pdz[d];
    }
}

/* plsa.py:795-814  plsa_refit_m_step — P(w|z) frozen, only P(z|d) re-estimated; serial. */
static void FN(refit_m_step)(const int32_t *rows, const REAL *vals, int64_t nnz,
                             REAL *pzd, const REAL *pzwd, REAL *norm_pdz, int64_t n,
                             int64_t k)
{
    memset(pzd, 0, sizeof(REAL) * (size_t)(n * k));
    memset(norm_pdz, 0, sizeof(REAL) * (size_t)n);
    for (int64_t nz = 0; nz < nnz; ++nz) {
        const int64_t d = rows[nz];
        const REAL x = vals[nz];
        const REAL *post = pzwd + nz * k;
        for (int64_t z = 0; z < k; ++z) {
            REAL s = x * post[z];
            pzd[d * k + z] += s;
            norm_pdz[d] += s;
        }
    }
    for (int64_t z = 0; z < k; ++z)
        for (int64_t d = 0; d < n; ++d)
            if (norm_pdz[d] > 0) pzd[d * k + z] /= norm_pdz[d];
}

/* plsa.py:372-386  log_likelihood — prange over nnz with a REAL-typed reduction
 * (`result` is float32 in the reference, plsa.py:322). */
static REAL FN(log_likelihood)(const int32_t *rows, const int32_t *cols, const REAL *vals,
                               int64_t nnz, const REAL *pwz, const REAL *pzd,
                               const REAL *sw, int64_t m, int64_t k)
{
    REAL result = 0;
#pragma omp parallel for schedule(static) reduction(+ : result)
    for (int64_t nz = 0; nz < nnz; ++nz) {
        const int64_t d = rows[nz], w = cols[nz];
        REAL p = 0;
        for (int64_t z = 0; z < k; ++z) p += pwz[z * m + w] * pzd[d * k + z];
        result += vals[nz] * (REAL)log((double)p) * sw[d];
    }
    return result;
}

/* plsa.py:583-640  plsa_fit_inner — EM loop and stopping rule.
 * LL is evaluated once before the loop and then at every i with i % n_iter_per_test == 0
 * (including i == 0); stop when change == 0 or change/|LL| < tolerance.
 * Returns the number of EM iterations carried out; ll_trace (optional, capacity ll_cap)
 * receives LL before the loop followed by each tested value. */
static int64_t FN(fit_inner)(const int32_t *rows, const int32_t *cols, const REAL *vals,
                             int64_t nnz, REAL *pwz, REAL *pzd, const REAL *sw, int64_t n,
                             int64_t m, int64_t k, int64_t n_iter, int64_t n_iter_per_test,
                             double tolerance, REAL thresh, int use_sample_weights,
                             double *ll_trace, int64_t ll_cap, int64_t *n_ll)
{
    REAL *pzwd = (REAL *)calloc((size_t)(nnz * k) + 1, sizeof(REAL)); /* plsa.py:586 */
    REAL *norm_pwz = (REAL *)calloc((size_t)k + 1, sizeof(REAL));
    REAL *norm_pdz = (REAL *)calloc((size_t)n + 1, sizeof(REAL));
    int64_t nl = 0, iters = 0;
    if (!pzwd || !norm_pwz || !norm_pdz) {
        free(pzwd); free(norm_pwz); free(norm_pdz);
        return -1;
    }

    REAL prev = FN(log_likelihood)(rows, cols, vals, nnz, pwz, pzd, sw, m, k);
    if (ll_trace && nl < ll_cap) ll_trace[nl] = (double)prev;
    nl++;

    for (int64_t i = 0; i < n_iter; ++i) {
        FN(e_step)(rows, cols, nnz, pwz, pzd, pzwd, k, m, thresh);
        FN(m_step)(rows, cols, vals, nnz, pwz, pzd, pzwd, use_sample_weights ? sw : NULL,
                   norm_pwz, norm_pdz, n, m, k);
        iters = i + 1;
        if (i % n_iter_per_test == 0) {
            REAL cur = FN(log_likelihood)(rows, cols, vals, nnz, pwz, pzd, sw, m, k);
            if (ll_trace && nl < ll_cap) ll_trace[nl] = (double)cur;
            nl++;
            REAL change = (REAL)fabs((double)(cur - prev));
            /* tolerance is a Python float (f64) in the reference; the ratio is f32/f32
             * promoted for the comparison. */
            if (change == 0 || (double)(change / (REAL)fabs((double)cur)) < tolerance)
                break;
            prev = cur;
        }
    }
    if (n_ll) *n_ll = nl;
    free(pzwd); free(norm_pwz); free(norm_pdz);
    return iters;
}

/* plsa.py:884-920  plsa_refit_inner — frozen topics.  The early stop is guarded by
 * `if current_log_likelihood > 0` (plsa.py:913), which never holds for a log-likelihood,
 * so all n_iter iterations always run; the guard is restated literally. */
static int64_t FN(refit_inner)(const int32_t *rows, const int32_t *cols, const REAL *vals,
                               int64_t nnz, const REAL *topics, REAL *pzd, const REAL *sw,
                               int64_t n, int64_t m, int64_t k, int64_t n_iter,
                               int64_t n_iter_per_test, double tolerance, REAL thresh)
{
    REAL *pzwd = (REAL *)calloc((size_t)(nnz * k) + 1, sizeof(REAL));
    REAL *norm_pdz = (REAL *)calloc((size_t)n + 1, sizeof(REAL));
    int64_t iters = 0;
    if (!pzwd || !norm_pdz) { free(pzwd); free(norm_pdz); return -1; }

    REAL prev = FN(log_likelihood)(rows, cols, vals, nnz, topics, pzd, sw, m, k);
    for (int64_t i = 0; i < n_iter; ++i) {
        FN(e_step)(rows, cols, nnz, topics, pzd, pzwd, k, m, thresh);
        FN(refit_m_step)(rows, vals, nnz, pzd, pzwd, norm_pdz, n, k);
        iters = i + 1;
        if (i % n_iter_per_test == 0) {
            REAL cur = FN(log_likelihood)(rows, cols, vals, nnz, topics, pzd, sw, m, k);
            if (cur > 0) {
                REAL change = (REAL)fabs((double)(cur - prev));
                if ((double)(change / (REAL)fabs((double)cur)) < tolerance) break;
                prev = cur;
            }
        }
    }
    free(pzwd); free(norm_pdz);
    return iters;
}

#undef FN
#undef CAT
#undef CAT_
