"""GPU: the tiled row pass (csrc/plsa_tile.cuh) forced on at sizes where the oracle is cheap.

A small tile (option "tile_kb") makes every code path appear on a 2000 x 5000 corpus: documents
with a head and a tail part, head-only and tail-only documents, term-side items over many blocks
of documents (tile reloads inside a CTA), tiled and untiled terms, sample weights, the fused
log-likelihood, refit.  Tolerances are those of tests/test_gpu_parity.py.
"""
import numpy as np
import pytest
from conftest import rel_l2

from enstop_b200 import _lib, plsa, synth
from oracle import oracle

pytestmark = pytest.mark.gpu

TOL_EXACT = 1e-5
TOL_REF_50 = 3e-4


def _fit(X, k, sw, init, n_iter, tile_kb, term_tiled=1, term_tile_min=2, per_test=10, tol=0.0,
         thresh=1e-32, options=()):
    with _lib.Context(0) as ctx:
        ctx.set_option("tiled", 1)
        ctx.set_option("tile_kb", tile_kb)
        ctx.set_option("term_tiled", term_tiled)
        ctx.set_option("term_tile_min", term_tile_min)
        for name, value in options:
            ctx.set_option(name, value)
        ctx.upload_csr(X)
        pzd, pwz, info = plsa.plsa_fit(X, k, sw, init=init, n_iter=n_iter, n_iter_per_test=per_test,
                                       tolerance=tol, e_step_thresh=thresh, context=ctx,
                                       return_info=True)
        ll = ctx.log_likelihood()
    return pzd, pwz, info, ll


@pytest.mark.parametrize("tile_kb,term_tiled", [(8, 1), (8, 0), (40, 1), (200, 1)])
def test_tiled_fit_vs_reference_and_exact(golden_c1_planted, tile_kb, term_tiled):
    g, X = golden_c1_planted
    sw = np.ones(X.shape[0], dtype=np.float32)
    init = (g["pzd0"], g["pwz0"])
    pzd, pwz, info, ll = _fit(X, 10, sw, init, 50, tile_kb, term_tiled)
    assert info["n_iter"] == 50
    assert rel_l2(pwz, g["pwz_50"]) < TOL_REF_50 and rel_l2(pzd, g["pzd_50"]) < TOL_REF_50
    ez, ew, einfo = oracle.plsa_fit(X, 10, sw, init=init, n_iter=50, tolerance=0.0, precision="f64",
                                    return_info=True)
    assert rel_l2(pwz, ew) < TOL_EXACT and rel_l2(pzd, ez) < TOL_EXACT
    assert np.allclose(info["ll_trace"], einfo["ll_trace"], rtol=1e-6)
    assert abs(ll - oracle.log_likelihood(X, ew, ez)) / abs(ll) < 1e-6


def test_tiled_equals_untiled_closely(golden_c1_zipf):
    """Same model to float32 rounding, same early stop, same trace as the group-per-row path."""
    g, X = golden_c1_zipf
    sw = np.ones(X.shape[0], dtype=np.float32)
    init = (g["pzd0"], g["pwz0"])
    a = _fit(X, 10, sw, init, 200, 16, per_test=5, tol=1e-4)
    with _lib.Context(0) as ctx:
        ctx.set_option("tiled", 0)
        ctx.upload_csr(X)
        b = plsa.plsa_fit(X, 10, sw, init=init, n_iter=200, n_iter_per_test=5, tolerance=1e-4,
                          context=ctx, return_info=True)
    assert a[2]["n_iter"] == b[2]["n_iter"] < 200
    assert rel_l2(a[0], b[0]) < 5e-6 and rel_l2(a[1], b[1]) < 5e-6
    assert np.allclose(a[2]["ll_trace"], b[2]["ll_trace"], rtol=1e-7)


@pytest.mark.parametrize("k", [1, 3, 7, 12, 20, 24])
def test_tiled_widths_weights_threshold(k):
    """Every row width the tile kernel is built for, with sample weights (term-side values carry
    them) and a visible E-step threshold."""
    X = synth.make_corpus(700, 900, 40_000, seed=k, planted=True, k_true=5)
    rng = np.random.RandomState(k)
    sw = rng.uniform(0.5, 1.5, size=X.shape[0]).astype(np.float32)
    init = plsa.plsa_init(X, k, "random", np.random.RandomState(k))
    init = (init[0].astype(np.float32), init[1].astype(np.float32))
    # a threshold near the median product P(w|z) P(z|d) ~ 1 / (n_terms k): about half are dropped
    for thresh in (1e-32, 0.5 / (X.shape[1] * k)):
        pzd, pwz, info, _ = _fit(X, k, sw, init, 4, 6, thresh=thresh)
        ez, ew, einfo = oracle.plsa_fit(X, k, sw, init=init, n_iter=4, tolerance=0.0,
                                        e_step_thresh=thresh, precision="f64", return_info=True)
        assert ew.any() and ez.any()
        # a visible threshold: the tiled pass cuts at thresh (1 +- 1e-7), so a
        # product that sits on the boundary may fall the other way than in the float64 oracle
        tol = TOL_EXACT if thresh < 1e-30 else 1e-4
        assert rel_l2(pwz, ew) < tol and rel_l2(pzd, ez) < tol, (k, thresh)
        if thresh < 1e-30:
            assert np.allclose(info["ll_trace"], einfo["ll_trace"], rtol=1e-6)


def test_tiled_refit_and_bit_repeatable(golden_small):
    g, X = golden_small
    ones = np.ones(X.shape[0], dtype=np.float32)
    outs = []
    for _ in range(2):
        with _lib.Context(0) as ctx:
            ctx.set_option("tiled", 1)
            ctx.set_option("tile_kb", 4)
            ctx.upload_csr(X)
            pzd = plsa.plsa_refit(X, g["pwz_20"], ones, n_iter=50, n_iter_per_test=5, tolerance=0.001,
                                  random_state=np.random.RandomState(42), context=ctx)
            outs.append(pzd)
    assert rel_l2(outs[0], g["refit_pzd"]) < TOL_REF_50
    assert np.array_equal(outs[0], outs[1])
    # empty documents stay zero rows, a never-seen term a zero column
    init = (g["pzd0"], g["pwz0"])
    a = _fit(X, int(g["k"]), ones, init, 20, 4)
    b = _fit(X, int(g["k"]), ones, init, 20, 4)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert not a[0][5].any() and not a[0][599].any() and not a[1][:, 17].any()
    assert rel_l2(a[1], g["pwz_20"]) < TOL_REF_50 and rel_l2(a[0], g["pzd_20"]) < TOL_REF_50


def test_tiled_host_plan_matches_device_plan(golden_c1_planted):
    g, X = golden_c1_planted
    sw = np.ones(X.shape[0], dtype=np.float32)
    init = (g["pzd0"], g["pwz0"])
    a = _fit(X, 10, sw, init, 5, 8)
    b = _fit(X, 10, sw, init, 5, 8, options=(("device_plan", 0),))
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
