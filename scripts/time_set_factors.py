"""Where do plsa_set_factors' 2 ms go?  (GPU box)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from enstop_b200 import _lib, synth
X = synth.make_config("C2")
n, m = X.shape
k = 20
ctx = _lib.Context(0)
ctx.upload_csr(X)
ctx.prepare(k, False)
pin_pzd, pin_pwz = ctx.pinned_factors(n, m, k)
rng = np.random.RandomState(0)
pin_pzd[:] = rng.rand(n, k); pin_pwz[:] = rng.rand(k, m)
page_pzd, page_pwz = pin_pzd.copy(), pin_pwz.copy()
for name, a, b in (("pinned", pin_pzd, pin_pwz), ("pageable", page_pzd, page_pwz)):
    ts = []
    for _ in range(6):
        t0 = time.perf_counter(); ctx.set_factors(a, b); ts.append((time.perf_counter() - t0) * 1e3)
    print("set_factors from %-8s ms: %s" % (name, [round(t, 2) for t in ts]))
ts = []
for _ in range(4):
    t0 = time.perf_counter(); ctx.get_factors(); ts.append((time.perf_counter() - t0) * 1e3)
print("get_factors ms:", [round(t, 2) for t in ts])
ts = []
for _ in range(4):
    t0 = time.perf_counter(); ctx.upload_csr(X); t1 = time.perf_counter(); ctx.prepare(k, False); t2 = time.perf_counter()
    ts.append((round((t1 - t0) * 1e3, 2), round((t2 - t1) * 1e3, 2)))
print("upload_csr, prepare ms:", ts)
ctx.close()
