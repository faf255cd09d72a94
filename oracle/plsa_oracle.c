/*
 * ORACLE — TEST INFRASTRUCTURE ONLY (see plsa_oracle_impl.h).
 *
 * C-ABI wrapper around the two instantiations of the restated reference algorithm.
 * Built by oracle/Makefile into oracle/libplsa_oracle.so and loaded with ctypes from
 * oracle/oracle.py.  Parity status: PINNED against outputs of the reference itself
 * (enstop/plsa.py run in the build container by tests/golden/make_golden.py; fixtures in
 * the .npz files under tests/golden; checked by tests/test_oracle_golden.py).
 */
#include <math.h>
#include <omp.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define REAL float
#define SUFFIX f32
#include "plsa_oracle_impl.h"
#undef REAL
#undef SUFFIX

#define REAL double
#define SUFFIX f64
#include "plsa_oracle_impl.h"
#undef REAL
#undef SUFFIX

#define API __attribute__((visibility("default")))

/* ---- float32, faithful ------------------------------------------------------------ */
API void oracle_e_step_f32(const int32_t *rows, const int32_t *cols, int64_t nnz,
                           const float *pwz, const float *pzd, float *pzwd, int64_t k,
                           int64_t m, float thresh)
{
    e_step_f32(rows, cols, nnz, pwz, pzd, pzwd, k, m, thresh);
}

API void oracle_m_step_f32(const int32_t *rows, const int32_t *cols, const float *vals,
                           int64_t nnz, float *pwz, float *pzd, const float *pzwd,
                           const float *sw_or_null, float *norm_pwz, float *norm_pdz,
                           int64_t n, int64_t m, int64_t k)
{
    m_step_f32(rows, cols, vals, nnz, pwz, pzd, pzwd, sw_or_null, norm_pwz, norm_pdz, n, m, k);
}

API float oracle_log_likelihood_f32(const int32_t *rows, const int32_t *cols,
                                    const float *vals, int64_t nnz, const float *pwz,
                                    const float *pzd, const float *sw, int64_t m, int64_t k)
{
    return log_likelihood_f32(rows, cols, vals, nnz, pwz, pzd, sw, m, k);
}

API int64_t oracle_fit_inner_f32(const int32_t *rows, const int32_t *cols, const float *vals,
                                 int64_t nnz, float *pwz, float *pzd, const float *sw,
                                 int64_t n, int64_t m, int64_t k, int64_t n_iter,
                                 int64_t n_iter_per_test, double tolerance, float thresh,
                                 int use_sample_weights, double *ll_trace, int64_t ll_cap,
                                 int64_t *n_ll)
{
    return fit_inner_f32(rows, cols, vals, nnz, pwz, pzd, sw, n, m, k, n_iter,
                         n_iter_per_test, tolerance, thresh, use_sample_weights, ll_trace,
                         ll_cap, n_ll);
}

API int64_t oracle_refit_inner_f32(const int32_t *rows, const int32_t *cols,
                                   const float *vals, int64_t nnz, const float *topics,
                                   float *pzd, const float *sw, int64_t n, int64_t m,
                                   int64_t k, int64_t n_iter, int64_t n_iter_per_test,
                                   double tolerance, float thresh)
{
    return refit_inner_f32(rows, cols, vals, nnz, topics, pzd, sw, n, m, k, n_iter,
                           n_iter_per_test, tolerance, thresh);
}

/* ---- float64, "exact" yardstick ---------------------------------------------------- */
API double oracle_log_likelihood_f64(const int32_t *rows, const int32_t *cols,
                                     const double *vals, int64_t nnz, const double *pwz,
                                     const double *pzd, const double *sw, int64_t m,
                                     int64_t k)
{
    return log_likelihood_f64(rows, cols, vals, nnz, pwz, pzd, sw, m, k);
}

API int64_t oracle_fit_inner_f64(const int32_t *rows, const int32_t *cols,
                                 const double *vals, int64_t nnz, double *pwz, double *pzd,
                                 const double *sw, int64_t n, int64_t m, int64_t k,
                                 int64_t n_iter, int64_t n_iter_per_test, double tolerance,
                                 double thresh, int use_sample_weights, double *ll_trace,
                                 int64_t ll_cap, int64_t *n_ll)
{
    return fit_inner_f64(rows, cols, vals, nnz, pwz, pzd, sw, n, m, k, n_iter,
                         n_iter_per_test, tolerance, thresh, use_sample_weights, ll_trace,
                         ll_cap, n_ll);
}

API int64_t oracle_refit_inner_f64(const int32_t *rows, const int32_t *cols,
                                   const double *vals, int64_t nnz, const double *topics,
                                   double *pzd, const double *sw, int64_t n, int64_t m,
                                   int64_t k, int64_t n_iter, int64_t n_iter_per_test,
                                   double tolerance, double thresh)
{
    return refit_inner_f64(rows, cols, vals, nnz, topics, pzd, sw, n, m, k, n_iter,
                           n_iter_per_test, tolerance, thresh);
}

/* utils.py:22-41  normalize(ndarray, axis=1) — in-place L1 row normalisation with a
 * float64 marginal; rows whose marginal is not > 0 are left untouched. */
API void oracle_normalize_rows_f64(double *a, int64_t n_rows, int64_t n_cols)
{
    for (int64_t i = 0; i < n_rows; ++i) {
        double marginal = 0.0;
        for (int64_t j = 0; j < n_cols; ++j) marginal += a[i * n_cols + j];
        if (marginal > 0.0)
            for (int64_t j = 0; j < n_cols; ++j) a[i * n_cols + j] /= marginal;
    }
}

/* Thread count of the `prange` loops (numba: NUMBA_NUM_THREADS).  n <= 0 restores the
 * OpenMP default.  Returns the count now in effect. */
API int oracle_set_num_threads(int n)
{
    omp_set_num_threads(n > 0 ? n : omp_get_num_procs());
    return omp_get_max_threads();
}
