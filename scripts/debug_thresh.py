"""Localise the thresh=1e-3 mismatch of the tiled path (GPU box)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from enstop_b200 import _lib, plsa, synth
from oracle import oracle

def rel(a, b):
    return float(np.linalg.norm(a.astype(np.float64) - b) / np.linalg.norm(b))

for k in (7, 12, 20):
    X = synth.make_corpus(700, 900, 40_000, seed=k, planted=True, k_true=5)
    sw = np.ones(X.shape[0], dtype=np.float32)
    init = plsa.plsa_init(X, k, "random", np.random.RandomState(k))
    init = (init[0].astype(np.float32), init[1].astype(np.float32))
    for n_iter in (1, 2):
        ez, ew = oracle.plsa_fit(X, k, sw, init=init, n_iter=n_iter, tolerance=0.0, e_step_thresh=1e-3, precision="f64")
        e32z, e32w = oracle.plsa_fit(X, k, sw, init=init, n_iter=n_iter, tolerance=0.0, e_step_thresh=1e-3, precision="f32")
        print("k", k, "iters", n_iter, "oracle f32 vs f64", rel(e32w, ew), rel(e32z, ez))
        for name, opts in (("untiled", dict(tiled=0)), ("doc tiled, small tile", dict(tiled=1, tile_kb=6, term_tiled=0)),
                           ("doc tiled, all head", dict(tiled=1, tile_kb=200, term_tiled=0)),
                           ("both tiled", dict(tiled=1, tile_kb=6, term_tiled=1, term_tile_min=2))):
            with _lib.Context(0) as ctx:
                for o, v in opts.items():
                    ctx.set_option(o, v)
                ctx.upload_csr(X)
                pzd, pwz = plsa.plsa_fit(X, k, sw, init=init, n_iter=n_iter, tolerance=0.0, e_step_thresh=1e-3, context=ctx)
            print("   %-24s pwz %.2e pzd %.2e   rows differing: pzd %d pwz-cols %d" % (
                name, rel(pwz, ew), rel(pzd, ez),
                int((np.abs(pzd - ez).max(axis=1) > 1e-4).sum()), int((np.abs(pwz - ew).max(axis=0) > 1e-6).sum())))
