"""Generate golden fixtures from the REAL reference (lmcinnes/enstop, /root/reference).

Run in the build container only (the reference does not travel to the GPU box):

    python tests/golden/make_golden.py

The reference package cannot be imported normally here (``enstop/__init__.py`` pulls in
dask / hdbscan / umap, none installed), so ``enstop`` is registered as a stub package whose
``__path__`` points at the read-only checkout and only ``enstop.plsa`` / ``enstop.utils``
are executed.  ``np.float`` (removed from numpy, used at enstop/utils.py:277) is not
needed because the class entry point is bypassed: the functions called are
``plsa_fit_inner`` (plsa.py:516), ``plsa_fit`` (:643), ``plsa_refit`` (:923),
``log_likelihood`` (:329) — exactly the seam the C-ABI replaces.

Outputs (tests/golden/*.npz) hold the input matrix, the float32 initial factors and the
reference's outputs, so the tests never need the reference at run time.
"""
import os
import sys
import types

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

REF = "/root/reference"


def load_reference():
    pkg = types.ModuleType("enstop")
    pkg.__path__ = [os.path.join(REF, "enstop")]
    sys.modules["enstop"] = pkg
    import enstop.plsa as ref_plsa  # noqa: E402  (numba JIT: ~10 s)
    return ref_plsa


def pack_matrix(X):
    X = sp.csr_matrix(X)
    X.sort_indices()
    return dict(indptr=X.indptr.astype(np.int32), indices=X.indices.astype(np.int32),
                data=np.asarray(X.data), shape=np.array(X.shape, dtype=np.int64))


def random_init(ref, X, k, seed):
    rng = np.random.RandomState(seed)
    pzd, pwz = ref.plsa_init(X, k, init="random", rng=rng)
    return pzd.astype(np.float32, order="C"), pwz.astype(np.float32, order="C")


def run_inner(ref, X, pwz0, pzd0, sw, n_iter, thresh=1e-32, use_sw=False,
              n_iter_per_test=10, tol=0.0):
    A = X.tocoo().astype(np.float32)
    pwz, pzd = pwz0.copy(), pzd0.copy()
    ref.plsa_fit_inner(A.row, A.col, A.data, pwz, pzd, sw, n_iter, n_iter_per_test, tol,
                       thresh, use_sw)
    return pzd, pwz


def ll(ref, X, pwz, pzd, sw):
    A = X.tocoo().astype(np.float32)
    return float(ref.log_likelihood(A.row, A.col, A.data, pwz, pzd, sw))


def main():
    from enstop_b200 import synth

    ref = load_reference()

    # ---- config 1 (2k x 5k, ~200k nnz, k=10): snapshots after 1, 10, 50 iterations ----
    for tag, planted in (("c1_zipf", False), ("c1_planted", True)):
        X = synth.make_config("C1", planted=planted)
        n, m = X.shape
        k = 10
        pzd0, pwz0 = random_init(ref, X, k, 42)
        sw = np.ones(n, dtype=np.float32)
        out = dict(**pack_matrix(X), pzd0=pzd0, pwz0=pwz0, k=np.int64(k))
        out["ll0"] = ll(ref, X, pwz0, pzd0, sw)
        for it in (1, 10, 50):
            pzd, pwz = run_inner(ref, X, pwz0, pzd0, sw, it)
            out[f"pzd_{it}"] = pzd
            out[f"pwz_{it}"] = pwz
            out[f"ll_{it}"] = ll(ref, X, pwz, pzd, sw)
        # plsa_fit end to end from the seed (random init inside), default tolerance
        pzd, pwz = ref.plsa_fit(X, k, sw, n_iter=50, random_state=42)
        out["fit_pzd"], out["fit_pwz"] = pzd, pwz
        np.savez_compressed(os.path.join(HERE, tag + ".npz"), **out)
        print(tag, "nnz", X.nnz, "ll0", out["ll0"], "ll50", out["ll_50"])

    # ---- small matrix: sample weights, threshold, refit, ragged rows, float input ------
    rng = np.random.default_rng(7)
    Xs = synth.make_corpus(600, 900, 30_000, seed=5, planted=True, k_true=6)
    Xs = Xs.tolil()
    Xs[5, :] = 0          # an all-zero document in the middle
    Xs[599, :] = 0        # and at the end
    Xs[:, 17] = 0         # a term that never occurs
    Xs = sp.csr_matrix(Xs)
    Xs.eliminate_zeros()
    n, m = Xs.shape
    k = 7
    pzd0, pwz0 = random_init(ref, Xs, k, 3)
    ones = np.ones(n, dtype=np.float32)
    sw = rng.uniform(0.25, 2.0, size=n).astype(np.float32)
    out = dict(**pack_matrix(Xs), pzd0=pzd0, pwz0=pwz0, k=np.int64(k), sw=sw)
    for it in (1, 20):
        pzd, pwz = run_inner(ref, Xs, pwz0, pzd0, ones, it)
        out[f"pzd_{it}"], out[f"pwz_{it}"] = pzd, pwz
        pzd, pwz = run_inner(ref, Xs, pwz0, pzd0, sw, it, use_sw=True)
        out[f"pzd_sw_{it}"], out[f"pwz_sw_{it}"] = pzd, pwz
        out[f"ll_sw_{it}"] = ll(ref, Xs, pwz, pzd, sw)
    # a visible threshold: products below 1e-3 are dropped (plsa.py:98-102)
    pzd, pwz = run_inner(ref, Xs, pwz0, pzd0, ones, 5, thresh=1e-3)
    out["pzd_thr"], out["pwz_thr"] = pzd, pwz
    # refit against frozen topics (plsa.py:923-997), seed 42 as PLSA.transform uses
    topics = out["pwz_20"]
    out["refit_pzd"] = ref.plsa_refit(Xs, topics, ones, n_iter=50, n_iter_per_test=5,
                                      tolerance=0.001, random_state=np.random.RandomState(42))
    # plsa_fit from a seed with default tolerance (early stop exercised) and with weights
    out["fit_pzd"], out["fit_pwz"] = ref.plsa_fit(Xs, k, ones, n_iter=100, random_state=11)
    out["fit_sw_pzd"], out["fit_sw_pwz"] = ref.plsa_fit(Xs, k, sw, n_iter=30, tolerance=0.0,
                                                        random_state=11)
    # float-valued (L1-normalised rows) input, as standardize_input produces (utils.py:278)
    from sklearn.preprocessing import normalize as sk_normalize
    Xf = sk_normalize(Xs.astype(np.float64), norm="l1")
    pzd, pwz = run_inner(ref, Xf, pwz0, pzd0, ones, 10)
    out["pzd_float"], out["pwz_float"] = pzd, pwz
    np.savez_compressed(os.path.join(HERE, "small_cases.npz"), **out)
    print("small_cases nnz", Xs.nnz)


if __name__ == "__main__":
    main()
