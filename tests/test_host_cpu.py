"""CPU: the C-ABI library loads and exports every declared symbol, the host-side mirror of
the reference API behaves like the reference before any device work, and the product path
fails loudly (no CPU fallback) when there is no GPU."""
import os
import re

import numpy as np
import pytest
import scipy.sparse as sp
from conftest import ROOT

from enstop_b200 import _lib, plsa, synth, utils
from oracle import oracle


def header_functions():
    text = open(os.path.join(ROOT, "include", "plsa_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(plsa_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    _lib.build()
    L = _lib.lib()
    names = header_functions()
    assert len(names) >= 25
    for name in names:
        assert hasattr(L, name), name
    # and the binding table covers exactly the header
    assert sorted(_lib.SIGNATURES) == names
    assert L.plsa_version() >= 200


def test_binary_is_sm100a_only():
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", _lib.SO_PATH], stdout=subprocess.PIPE, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_\d+a?", out.stdout))
    assert archs == {"sm_100a"}, archs


def test_no_silent_cpu_fallback():
    if _lib.device_count() > 0:
        pytest.skip("a GPU is present")
    X = synth.make_corpus(50, 60, 500, seed=0)
    with pytest.raises(_lib.PlsaError):
        plsa.plsa_fit(X, 3, np.ones(50, dtype=np.float32), n_iter=2)
    with pytest.raises(_lib.PlsaError):
        plsa.PLSA(n_components=3).fit(X)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "enstop_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(base, f)).read()
                assert "oracle" not in text.lower().replace("oracle (", ""), os.path.join(base, f)


def test_random_init_is_the_references_stream(golden_c1_zipf):
    """plsa.py:455-456: P(w|z) drawn first, then P(z|d); float64 L1 normalisation; the
    float32 cast equals the reference's own initial factors stored in the fixture."""
    g, X = golden_c1_zipf
    pzd, pwz = plsa.plsa_init(X, 10, "random", np.random.RandomState(42))
    assert pzd.dtype == np.float64 and pwz.dtype == np.float64
    assert np.array_equal(pzd.astype(np.float32), g["pzd0"])
    assert np.array_equal(pwz.astype(np.float32), g["pwz0"])
    opzd, opwz = oracle.plsa_init_random(X.shape[0], X.shape[1], 10, np.random.RandomState(42))
    assert np.allclose(pzd, opzd, rtol=1e-15) and np.allclose(pwz, opwz, rtol=1e-15)


@pytest.mark.parametrize("seed,burn,rows,cols", [(42, 0, 10, 5000), (42, 13, 2000, 10),
                                                 (7, 1, 5, 3), (123, 311, 1000, 7), (5, 0, 311, 1),
                                                 (9, 2, 3, 624), (9, 0, 1, 1249), (0, 623, 64, 20)])
def test_host_random_rows_is_numpy_bit_for_bit(seed, burn, rows, cols):
    """csrc/host_init.cpp advances the RandomState's MT19937 state exactly as numpy's
    rand() does and reproduces rand + L1 row normalisation + float32 cast bit for bit."""
    a, b = np.random.RandomState(seed), np.random.RandomState(seed)
    if burn:
        a.randint(0, 10, size=burn)   # odd numbers of 32-bit draws shift the pair alignment
        b.randint(0, 10, size=burn)
    x = a.rand(rows, cols)
    x /= x.sum(axis=1, keepdims=True)
    y32, y64 = _lib.random_rows(b, rows, cols, want_f64=True)
    assert np.array_equal(x.astype(np.float32), y32)
    assert np.allclose(x, y64, rtol=1e-13, atol=0)     # numpy sums pairwise, the reference left to right
    assert np.array_equal(a.rand(7), b.rand(7))           # the streams stay in step
    assert a.get_state()[2] == b.get_state()[2]


@pytest.mark.parametrize("seed,burn,rows,cols", [(42, 13, 40000, 10), (3, 5, 7, 60001),
                                                 (11, 622, 30011, 13), (8, 0, 17, 20000)])
def test_host_random_rows_parallel_slices(seed, burn, rows, cols):
    """From 2^18 values the draw is cut into slices of unequal length, one per thread, each
    started from the state skipped ahead to its first word; the last slice runs on the
    caller's state: same bits, same final state as numpy's serial draw."""
    a, b = np.random.RandomState(seed), np.random.RandomState(seed)
    if burn:
        a.randint(0, 10, size=burn)
        b.randint(0, 10, size=burn)
    x = a.rand(rows, cols)
    x /= np.cumsum(x, axis=1)[:, -1:]          # left-to-right float64 marginal (utils.py:25-29)
    y32 = _lib.random_rows(b, rows, cols)
    assert np.array_equal(x.astype(np.float32), y32)
    assert np.array_equal(a.rand(7), b.rand(7))
    assert a.get_state()[2] == b.get_state()[2]


@pytest.mark.parametrize("offset", [1, 2, 3])
def test_host_random_rows_into_a_misaligned_buffer(offset):
    """The factors are written with streaming stores (16-byte where aligned, 4-byte at the
    edges): any float32-aligned destination works, e.g. a view that starts 4 bytes in."""
    rows, cols = 301, 21
    a, b = np.random.RandomState(17), np.random.RandomState(17)
    flat = np.full(rows * cols + 8, -1.0, dtype=np.float32)
    out = flat[offset:offset + rows * cols].reshape(rows, cols)
    got = _lib.random_rows(b, rows, cols, out=out)
    assert got is out
    x = a.rand(rows, cols)
    x /= np.cumsum(x, axis=1)[:, -1:]
    assert np.array_equal(out, x.astype(np.float32))
    assert np.all(flat[:offset] == -1.0) and np.all(flat[offset + rows * cols:] == -1.0)


def test_fast_random_init_equals_plsa_init(golden_c1_zipf):
    g, X = golden_c1_zipf
    fast = plsa._random_init_f32(X.shape[0], X.shape[1], 10, np.random.RandomState(42))
    assert np.array_equal(fast[0], g["pzd0"]) and np.array_equal(fast[1], g["pwz0"])
    # the module-level generator (random_state=None) is advanced in place as well
    st = np.random.get_state()
    try:
        np.random.seed(5)
        ref = plsa.plsa_init(X, 10, "random", np.random)
        after_ref = np.random.rand()
        np.random.seed(5)
        fast = plsa._random_init_f32(X.shape[0], X.shape[1], 10, np.random)
        assert np.random.rand() == after_ref
        assert np.array_equal(fast[0], ref[0].astype(np.float32))
        assert np.array_equal(fast[1], ref[1].astype(np.float32))
    finally:
        np.random.set_state(st)
    assert plsa._random_init_f32(5, 6, 2, np.random.default_rng(0)) is None   # not MT19937


def test_init_variants():
    X = synth.make_corpus(200, 300, 6000, seed=2, planted=True, k_true=4).astype(np.float64)
    for init in ("nndsvd", "nmf"):
        pzd, pwz = plsa.plsa_init(X, 4, init, np.random.RandomState(0))
        assert pzd.shape == (200, 4) and pwz.shape == (4, 300)
        assert pzd.min() >= 0 and pwz.min() >= 0
        assert np.allclose(pwz.sum(axis=1), 1.0)
        rows = pzd.sum(axis=1)
        assert np.all(np.isclose(rows, 1.0) | (rows == 0.0))
    a, b = np.random.rand(200, 4), np.random.rand(4, 300)
    pzd, pwz = plsa.plsa_init(X, 4, (a, b))
    assert np.allclose(pzd, a / a.sum(1, keepdims=True))
    with pytest.raises(ValueError, match="Unrecognized init"):
        plsa.plsa_init(X, 4, "bogus")
    with pytest.raises(ValueError):
        plsa.plsa_init(X, 4, (a[:, :3], b))


def test_normalize_semantics():
    a = np.array([[1.0, 3.0], [0.0, 0.0], [2.0, 2.0]])
    utils.normalize(a, axis=1)
    assert np.allclose(a, [[0.25, 0.75], [0.0, 0.0], [0.5, 0.5]])
    b = np.array([[1.0, 0.0], [3.0, 0.0]])
    utils.normalize(b, axis=0)
    assert np.allclose(b, [[0.25, 0.0], [0.75, 0.0]])
    c = np.random.RandomState(0).rand(5, 7)
    d = c.copy()
    utils.normalize(c, axis=1)
    oracle.normalize_rows(d)
    assert np.allclose(c, d, rtol=1e-15)


def test_standardize_input_and_validation():
    Xi = sp.csr_matrix(np.array([[1, 2], [0, 3]], dtype=np.int64))
    assert utils.standardize_input(Xi) is Xi
    Xf = utils.standardize_input(Xi.astype(np.float64))
    assert np.allclose(Xf.toarray(), [[1 / 3, 2 / 3], [0, 1]])
    Xn = sp.csr_matrix(np.array([[1.0, -2.0], [0.0, 3.0]]))
    with pytest.raises(ValueError, match="non-negative"):
        plsa.PLSA(n_components=2).fit(Xn)   # raised before any device work
    m = plsa.PLSA()
    assert m.get_params()["n_components"] == 10 and m.get_params()["e_step_thresh"] == 1e-32
    assert m.get_params()["transform_random_seed"] == 42 and m.get_params()["n_iter"] == 100


def test_metrics_match_a_direct_computation():
    X = synth.make_corpus(120, 80, 2500, seed=4, planted=True, k_true=3)
    topics = np.random.RandomState(0).rand(3, 80)
    topics /= topics.sum(axis=1, keepdims=True)
    # log lift over the full vocabulary (utils.py:70-74)
    p = np.asarray(X.sum(axis=0), dtype=np.float64).ravel()
    p /= p.sum()
    ok = p > 0
    expect = np.log(np.sum(topics[1][ok] / p[ok]) / 80)
    assert np.isclose(utils.log_lift(topics, 1, X), expect)
    # coherence (utils.py:196-207)
    top = np.argsort(topics[0])[-5:]
    D = (X > 0).toarray()
    ndw = D.sum(axis=0)
    tot = 0.0
    for i in range(4):
        for j in range(i + 1, 5):
            tot += np.log((np.sum(D[:, top[i]] & D[:, top[j]]) + 1.0) / ndw[top[i]])
    assert np.isclose(utils.coherence(topics, 0, X, n_words=5), tot)
    assert np.isfinite(utils.mean_coherence(topics, X, 5)) and np.isfinite(utils.mean_log_lift(topics, X))


def test_synth_recipe():
    X, info = synth.make_corpus(2000, 5000, 200_000, seed=1, return_info=True)
    assert abs(X.nnz - 200_000) / 200_000 < 0.03
    assert X.dtype == np.int32 and X.has_sorted_indices
    assert info["sum_counts"] == info["tokens"]
    Y = synth.make_corpus(2000, 5000, 200_000, seed=1)
    assert (X != Y).nnz == 0


def _ragged_indptr(rng, rows, long_rows):
    """Row pointers with empty rows, short rows and a few rows far longer than a chunk."""
    lens = rng.integers(0, 40, size=rows)
    lens[rng.random(rows) < 0.2] = 0
    for r in rng.choice(rows, size=long_rows, replace=False):
        lens[r] = int(rng.integers(300, 20_000))
    indptr = np.zeros(rows + 1, dtype=np.int32)
    np.cumsum(lens, out=indptr[1:])
    return indptr


@pytest.mark.parametrize("chunk,align", [(256, 4), (32, 4), (64, 1), (2048, 2), (100, 4)])
def test_work_item_plan_covers_every_entry_once(chunk, align):
    """plan_items (host side of the row pass, csrc/plsa_b200.cu): the
    items tile the stored entries exactly, chunks of a split row own consecutive slots, items
    start on entry-block boundaries and are sorted longest first."""
    rng = np.random.default_rng(chunk * 7 + align)
    indptr = _ragged_indptr(rng, 3000, 12)
    p = _lib.plan_items(indptr, chunk, align=align)
    start, row, ln, slot, skip = p["start"], p["row"], p["len"], p["slot"], p["skip"]
    eff = chunk // align * align
    assert np.all(ln <= eff) and np.all(ln >= 0) and np.all(skip >= 0) and np.all(skip < align)
    assert np.all(start % align == 0)
    assert np.all(np.diff(ln) <= 0), "items must be sorted by length, longest first"
    # every stored entry is covered by exactly one item of its own row
    cover = np.zeros(int(indptr[-1]), dtype=np.int32)
    owner = np.full(int(indptr[-1]), -1, dtype=np.int64)
    for s, r, l, k in zip(start, row, ln, skip):
        cover[s + k:s + l] += 1
        owner[s + k:s + l] = r
        assert indptr[r] <= s + k and s + l <= indptr[r + 1]
    assert np.all(cover == 1)
    assert np.array_equal(owner, np.repeat(np.arange(indptr.shape[0] - 1), np.diff(indptr)))
    # one item per unsplit row (empty rows included: their result is a zero row)
    whole = slot < 0
    n_rows = indptr.shape[0] - 1
    split_rows = np.unique(row[~whole])
    assert np.array_equal(np.sort(np.concatenate([row[whole], split_rows])), np.arange(n_rows))
    assert p["n_split"] == split_rows.shape[0]
    # slots: a permutation of 0..n_slots-1; within a row consecutive and in entry order
    assert np.array_equal(np.sort(slot[~whole]), np.arange(p["n_slots"]))
    for r in split_rows:
        sel = np.flatnonzero(row == r)
        by_start = sel[np.argsort(start[sel])]
        assert np.array_equal(slot[by_start], slot[by_start][0] + np.arange(sel.shape[0]))
        assert np.all(skip[by_start][1:] == 0)


def test_work_item_plan_rejects_bad_alignment():
    rng = np.random.default_rng(11)
    indptr = _ragged_indptr(rng, 200, 40)
    with pytest.raises(_lib.PlsaError):
        _lib.plan_items(indptr, 256, align=3)


def _emulate_pass(indptr, idx, val, own, gat, plan, thresh, normalise):
    """What row_pass_kernel + fixup_kernel compute from a work-item plan (float64, numpy):
    per item the E-step posterior of plsa.py:95-104 folded into the M-step sums of
    plsa.py:189-194; whole rows are written directly, chunks into their partial-sum slot, and
    the slots of a split row are added in slot order."""
    k = own.shape[1]
    own_new = np.zeros_like(own)
    partial = np.zeros((max(plan["n_slots"], 1), k))
    for s, r, l, slot, skip in zip(plan["start"], plan["row"], plan["len"], plan["slot"],
                                   plan["skip"]):
        e = slice(s + skip, s + l)
        v = own[r][None, :] * gat[idx[e]]
        v[v <= thresh] = 0.0
        norm = v.sum(axis=1)
        c = np.divide(val[e], norm, out=np.zeros_like(norm), where=norm > 0)
        acc = (c[:, None] * v).sum(axis=0)
        if slot < 0:
            own_new[r] = acc
        else:
            partial[slot] = acc
    chunks = np.flatnonzero(plan["slot"] >= 0)
    for r in np.unique(plan["row"][chunks]):
        sl = np.sort(plan["slot"][chunks][plan["row"][chunks] == r])
        assert np.array_equal(sl, np.arange(sl[0], sl[0] + sl.shape[0]))
        own_new[r] = partial[sl].sum(axis=0)
    if normalise:  # plsa.py:199-202
        tot = own_new.sum(axis=1, keepdims=True)
        np.divide(own_new, tot, out=own_new, where=tot > 0)
    return own_new


def test_planned_passes_are_one_em_iteration():
    """Host-side plan semantics end to end on the CPU: a doc pass and a term pass carried out
    item by item from plsa_plan_items (aligned items, split rows), with
    the lazily normalised P(w|z), equal one EM iteration of the oracle's float64
    restatement of plsa_fit_inner."""
    X = synth.make_corpus(400, 300, 14_000, seed=9, planted=True, k_true=4).astype(np.float64)
    n, m = X.shape
    k = 6
    rng = np.random.RandomState(5)
    pzd, pwz = oracle.plsa_init_random(n, m, k, rng)
    Xt = X.T.tocsr()
    Xt.sort_indices()
    thresh = 1e-32
    plan_d = _lib.plan_items(X.indptr, 32, align=4)
    plan_t = _lib.plan_items(Xt.indptr, 32, align=4)
    assert plan_d["n_split"] > 0 and plan_t["n_split"] > 0
    new_pzd = _emulate_pass(X.indptr, X.indices, X.data, pzd, pwz.T.copy(), plan_d, thresh, True)
    raw = _emulate_pass(Xt.indptr, Xt.indices, Xt.data, pwz.T.copy(), pzd, plan_t, thresh, False)
    col = raw.sum(axis=0)                                   # plsa.py:196-198
    new_pwz = (raw / np.where(col > 0, col, 1.0)).T
    A = X.tocoo()
    e_pwz, e_pzd = pwz.copy(), pzd.copy()
    iters, _ = oracle.fit_inner(A.row.astype(np.int32), A.col.astype(np.int32),
                                A.data.astype(np.float64), e_pwz, e_pzd, np.ones(n), n_iter=1,
                                tolerance=0.0, e_step_thresh=thresh, precision="f64")
    assert iters == 1
    assert np.allclose(new_pzd, e_pzd, rtol=1e-10, atol=1e-15)
    assert np.allclose(new_pwz, e_pwz, rtol=1e-10, atol=1e-15)


def test_combine_matches_reference_golden():
    """Cluster representatives (mean of sqrt(topic), squared, renormalised) against the
    reference's own statements (enstop_.py:309-312, :397-411; tests/golden/make_golden_combine.py)."""
    from enstop_b200 import enstop_
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "topic_combine.npz"))
    out = enstop_._combine(g["all_topics"], g["labels"])
    assert out.dtype == np.float32 and out.shape == g["combined_mean"].shape
    assert np.allclose(out, g["combined_mean"], rtol=1e-6, atol=0)
    outw = enstop_._combine(g["all_topics"], g["labels"], g["strengths"])
    assert np.allclose(outw, g["combined_weighted"], rtol=1e-6, atol=0)
