"""Where does PLSA.fit spend its wall time at C2?  (run on the GPU box)"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from sklearn.utils import check_array, check_random_state
from enstop_b200 import _lib, plsa, synth
from enstop_b200.utils import standardize_input, _check_sample_weight, normalize

X = synth.make_config(sys.argv[1] if len(sys.argv) > 1 else "C2")
k = 20
n_iter = 100
plsa.PLSA(n_components=k, n_iter=3, tolerance=0.0, random_state=42).fit(X)  # warm

def T(label, fn, reps=1):
    t0 = time.perf_counter(); out = fn(); dt = time.perf_counter() - t0
    print("%-34s %8.2f ms" % (label, dt * 1e3)); return out

t_all = time.perf_counter()
Xc = T("check_array", lambda: check_array(X, accept_sparse="csr"))
Xc = T("standardize_input", lambda: standardize_input(Xc))
sw = T("_check_sample_weight", lambda: _check_sample_weight(None, Xc, dtype=np.float32))
T("negative check", lambda: np.any(Xc.data < 0))
rs = T("row sums", lambda: np.array(Xc.sum(axis=1).T)[0])
rng = check_random_state(42)
pw = T("rng.rand(k,m)", lambda: rng.rand(k, X.shape[1]))
pz = T("rng.rand(n,k)", lambda: rng.rand(X.shape[0], k))
T("normalize x2", lambda: (normalize(pw, axis=1), normalize(pz, axis=1)))
pz32 = T("astype f32 x2", lambda: pz.astype(np.float32, order="C")); pw32 = pw.astype(np.float32, order="C")
T("any(sw != 1)", lambda: np.any(sw != 1.0))
ctx = T("Context()", lambda: _lib.Context(0))
T("upload_csr", lambda: ctx.upload_csr(Xc))
T("set_factors", lambda: ctx.set_factors(pz32, pw32))
T("set_sample_weight", lambda: ctx.set_sample_weight(sw))
T("em(1) incl. term-major build", lambda: ctx.em(1, 10, 0.0))
T("em(%d)" % n_iter, lambda: ctx.em(n_iter, 10, 0.0))
print("   device EM ms", ctx.last_em_ms)
T("get_factors", lambda: ctx.get_factors())
T("close", lambda: ctx.close())
print("sum of phases %.2f ms" % ((time.perf_counter() - t_all) * 1e3))
t0 = time.perf_counter()
plsa.PLSA(n_components=k, n_iter=n_iter, tolerance=0.0, random_state=42).fit(X)
print("PLSA.fit end to end %.2f ms" % ((time.perf_counter() - t0) * 1e3))

# ---- steady state: phases inside plsa_fit on a pooled context -------------------------
import threading
from enstop_b200.plsa import _Staging, plsa_init
for rep in range(3):
    t = {}
    t0 = time.perf_counter()
    st = _Staging(X, k, None, None, refit=False)
    t["thread start"] = time.perf_counter() - t0
    t1 = time.perf_counter()
    rng = check_random_state(42)
    from enstop_b200.plsa import _random_init_f32
    p, w = _random_init_f32(X.shape[0], X.shape[1], k, rng)
    t["init (main thread)"] = time.perf_counter() - t1
    t1 = time.perf_counter(); ctx = st.wait(); t["wait for staging"] = time.perf_counter() - t1
    t1 = time.perf_counter(); ctx.set_factors(p, w); ctx.set_sample_weight(None); t["set_factors"] = time.perf_counter() - t1
    t1 = time.perf_counter(); ctx.em(n_iter, 10, 0.0); t["em"] = time.perf_counter() - t1
    t1 = time.perf_counter(); ctx.get_factors(); t["get_factors"] = time.perf_counter() - t1
    t1 = time.perf_counter(); st.close(); t["release"] = time.perf_counter() - t1
    t["total"] = time.perf_counter() - t0
    print("steady", rep, {a: round(b * 1e3, 2) for a, b in t.items()})

# ---- PLSA.fit_transform pieces before plsa_fit
for rep in range(2):
    t0 = time.perf_counter(); Xc = check_array(X, accept_sparse="csr"); t1 = time.perf_counter()
    dm = Xc.data.min(); t2 = time.perf_counter()
    good = np.diff(Xc.indptr) != 0; allg = np.all(good); t3 = time.perf_counter()
    sw = _check_sample_weight(None, Xc, dtype=np.float32); t4 = time.perf_counter()
    print("front", rep, "check_array %.2f min %.2f rows %.2f sw %.2f ms" % ((t1-t0)*1e3,(t2-t1)*1e3,(t3-t2)*1e3,(t4-t3)*1e3))
for rep in range(5):
    t0 = time.perf_counter()
    plsa.PLSA(n_components=k, n_iter=n_iter, tolerance=0.0, random_state=42).fit(X)
    print("PLSA.fit end to end %.2f ms" % ((time.perf_counter() - t0) * 1e3))
