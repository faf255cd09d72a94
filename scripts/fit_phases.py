"""Phases of PLSA(n_iter=20).fit at C2 on a pooled (warm) context, inputs page-locked or
pageable.  Run on the GPU box: python scripts/fit_phases.py [C2] [pinned|pageable]"""
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from sklearn.utils import check_random_state
from enstop_b200 import _lib, plsa, synth

cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
pinned = (sys.argv[2] if len(sys.argv) > 2 else "pinned") == "pinned"
X = synth.make_config(cfg)
if pinned:
    X = _lib.pinned_csr(X)
k, n_iter = 20, 20
DRAIN = os.environ.get("DRAIN", "0") == "1"
if DRAIN:
    import torch
    torch.cuda.init()
est = lambda: plsa.PLSA(n_components=k, n_iter=n_iter, tolerance=0.0, random_state=42)
est().fit(X)
for rep in range(3):
    t0 = time.perf_counter()
    est().fit(X)
    print("PLSA.fit %s inputs: %.2f ms" % ("pinned" if pinned else "pageable", 1e3 * (time.perf_counter() - t0)))

ms = lambda a, b: round(1e3 * (b - a), 2)
for rep in range(4):
    t = {}
    t0 = time.perf_counter()
    ctx = _lib.acquire_context(0)
    stamps = {}

    def stage():
        a = time.perf_counter()
        ctx.upload_csr(X)
        b = time.perf_counter()
        ctx.prepare(k, False)
        c = time.perf_counter()
        stamps.update(upload=(a, b), prepare=(b, c))

    th = threading.Thread(target=stage)
    th.start()
    a = time.perf_counter()
    rng = check_random_state(42)
    p, w = plsa._random_init_f32(X.shape[0], X.shape[1], k, rng, ctx=ctx)
    b = time.perf_counter()
    th.join()
    c0 = time.perf_counter()
    if DRAIN:
        torch.cuda.synchronize()      # whatever prepare left running on the device
    c = time.perf_counter()
    ctx.set_factors(p, w)
    c1 = time.perf_counter()
    ctx.set_sample_weight(None)
    d = time.perf_counter()
    ctx.em(n_iter, 10, 0.0, 1e-32)
    e = time.perf_counter()
    ctx.get_factors()
    f = time.perf_counter()
    _lib.release_context(ctx)
    g = time.perf_counter()
    print("phases", rep, {"acquire": ms(t0, a), "init draw (main)": ms(a, b),
                           "upload (helper, from t0)": [ms(t0, stamps["upload"][0]), ms(t0, stamps["upload"][1])],
                           "prepare (helper)": ms(*stamps["prepare"]), "join at": ms(t0, c0),
                           "drain": ms(c0, c), "set_factors": ms(c, c1), "set_sample_weight": ms(c1, d), "em": ms(d, e), "device em": round(ctx.last_em_ms, 2),
                           "get_factors": ms(e, f), "release": ms(f, g), "total": ms(t0, g)})
