"""CPU: the oracle (C restatement of enstop/plsa.py) against fixtures produced by the
reference itself (tests/golden/make_golden.py).  This is what pins the oracle."""
import numpy as np
import pytest
from conftest import rel_l2

from oracle import oracle

# float32 restatement vs the numba reference: rounding-level agreement (both are float32
# with fastmath reassociation; measured 1e-7..7e-7).
TOL_F32 = 5e-6


@pytest.mark.parametrize("tag", ["golden_c1_zipf", "golden_c1_planted"])
@pytest.mark.parametrize("it", [1, 10, 50])
def test_fit_inner_matches_reference(tag, it, request):
    g, X = request.getfixturevalue(tag)
    sw = np.ones(X.shape[0], dtype=np.float32)
    pzd, pwz, info = oracle.plsa_fit(X, int(g["k"]), sw, init=(g["pzd0"], g["pwz0"]),
                                     n_iter=it, tolerance=0.0, return_info=True)
    assert info["n_iter"] == it
    assert rel_l2(pwz, g[f"pwz_{it}"]) < TOL_F32
    assert rel_l2(pzd, g[f"pzd_{it}"]) < TOL_F32
    # log-likelihood of the reference's own factors (plsa.py:329-386)
    ll = oracle.log_likelihood(X, g[f"pwz_{it}"], g[f"pzd_{it}"], precision="f64")
    assert abs(ll - float(g[f"ll_{it}"])) / abs(ll) < 2e-6


@pytest.mark.parametrize("tag", ["golden_c1_zipf", "golden_c1_planted"])
def test_plsa_fit_seeded_matches_reference(tag, request):
    """Random init from RandomState(42) (plsa.py:455-456 draw order) + default tolerance."""
    g, X = request.getfixturevalue(tag)
    sw = np.ones(X.shape[0], dtype=np.float32)
    pzd, pwz = oracle.plsa_fit(X, int(g["k"]), sw, n_iter=50, random_state=42)
    assert rel_l2(pwz, g["fit_pwz"]) < TOL_F32
    assert rel_l2(pzd, g["fit_pzd"]) < TOL_F32


def test_f64_yardstick_distance_is_the_references_own_error(golden_c1_zipf):
    """The float64 EM sits ~1e-4 from the reference after 50 iterations (the reference's
    float32 accumulation error, SURVEY.md hard part 1) and ~1e-5 after one."""
    g, X = golden_c1_zipf
    sw = np.ones(X.shape[0], dtype=np.float32)
    pzd, pwz = oracle.plsa_fit(X, 10, sw, init=(g["pzd0"], g["pwz0"]), n_iter=1,
                               tolerance=0.0, precision="f64")
    assert rel_l2(pwz, g["pwz_1"]) < 5e-5
    pzd, pwz = oracle.plsa_fit(X, 10, sw, init=(g["pzd0"], g["pwz0"]), n_iter=50,
                               tolerance=0.0, precision="f64")
    assert rel_l2(pwz, g["pwz_50"]) < 5e-4


def test_small_cases(golden_small):
    g, X = golden_small
    k = int(g["k"])
    n = X.shape[0]
    ones = np.ones(n, dtype=np.float32)
    sw = g["sw"]
    init = (g["pzd0"], g["pwz0"])
    for it in (1, 20):
        pzd, pwz = oracle.plsa_fit(X, k, ones, init=init, n_iter=it, tolerance=0.0)
        assert rel_l2(pwz, g[f"pwz_{it}"]) < TOL_F32 and rel_l2(pzd, g[f"pzd_{it}"]) < TOL_F32
        pzd, pwz = oracle.plsa_fit(X, k, sw, init=init, n_iter=it, tolerance=0.0)
        assert rel_l2(pwz, g[f"pwz_sw_{it}"]) < TOL_F32
        assert rel_l2(pzd, g[f"pzd_sw_{it}"]) < TOL_F32
    # empty documents stay all-zero rows, the never-seen term an all-zero column
    assert not pzd[5].any() and not pzd[599].any() and not pwz[:, 17].any()
    pzd, pwz = oracle.plsa_fit(X, k, ones, init=init, n_iter=5, tolerance=0.0,
                               e_step_thresh=1e-3)
    assert rel_l2(pwz, g["pwz_thr"]) < TOL_F32 and rel_l2(pzd, g["pzd_thr"]) < TOL_F32
    pzd = oracle.plsa_refit(X, g["pwz_20"], ones, n_iter=50, n_iter_per_test=5,
                            tolerance=0.001, random_state=np.random.RandomState(42))
    assert rel_l2(pzd, g["refit_pzd"]) < TOL_F32
    pzd, pwz = oracle.plsa_fit(X, k, ones, n_iter=100, random_state=11)
    assert rel_l2(pwz, g["fit_pwz"]) < TOL_F32 and rel_l2(pzd, g["fit_pzd"]) < TOL_F32
    pzd, pwz = oracle.plsa_fit(X, k, sw, n_iter=30, tolerance=0.0, random_state=11)
    assert rel_l2(pwz, g["fit_sw_pwz"]) < TOL_F32 and rel_l2(pzd, g["fit_sw_pzd"]) < TOL_F32


def test_float_input(golden_small):
    from sklearn.preprocessing import normalize as sk_normalize
    g, X = golden_small
    Xf = sk_normalize(X.astype(np.float64), norm="l1")
    ones = np.ones(X.shape[0], dtype=np.float32)
    pzd, pwz = oracle.plsa_fit(Xf, int(g["k"]), ones, init=(g["pzd0"], g["pwz0"]),
                               n_iter=10, tolerance=0.0)
    assert rel_l2(pwz, g["pwz_float"]) < TOL_F32 and rel_l2(pzd, g["pzd_float"]) < TOL_F32


def test_topic_distances_against_reference():
    """all_pairs_kl_divergence / all_pairs_hellinger_distance (enstop_.py:234-263): the
    oracle's numpy restatement against the reference's own numba functions
    (tests/golden/make_golden_distances.py)."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden",
                             "topic_distances.npz"))
    topics = g["topics"]
    K = oracle.all_pairs_kl_divergence(topics)
    H = oracle.all_pairs_hellinger_distance(topics)
    # the reference takes log2 of float32 values in float32: ~1e-7 per term
    assert np.allclose(K, g["kl"], rtol=1e-6, atol=2e-6)
    # the reference's 1 - inner/denominator form leaves ~1e-8 rounding under the root for
    # identical rows: compare squared distances
    assert np.allclose(H ** 2, g["hellinger"] ** 2, atol=1e-7)
    assert H[11, 0] == 1.0 and H[11, 11] == 0.0 and g["hellinger"][11, 0] == 1.0
