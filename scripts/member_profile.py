"""Where does an ensemble member's wall time go, and how does it scale with host threads?
(C2 corpus, k=20, 80 iterations; run on a GPU box)"""
import os, sys, time, threading, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from sklearn.utils import check_random_state
from enstop_b200 import _lib, enstop_, synth

X = synth.make_config("C2", planted=True).astype(np.float32)
n, m = X.shape
k = 20


def member(ctx, seed, T):
    def tick(name, t0):
        T[name] += time.perf_counter() - t0
    t0 = time.perf_counter(); idx = enstop_.bootstrap_indices(n, seed); tick("randint", t0)
    t0 = time.perf_counter(); ctx.bootstrap(idx); tick("bootstrap", t0)
    t0 = time.perf_counter(); ctx.prepare(k, False); tick("prepare(term-major+items)", t0)
    rng = check_random_state(seed)
    t0 = time.perf_counter(); pwz = _lib.random_rows(rng, k, m); pzd = _lib.random_rows(rng, ctx.shape[0], k); tick("seeded init", t0)
    t0 = time.perf_counter(); ctx.set_factors(pzd, pwz); ctx.set_sample_weight(None); tick("set_factors", t0)
    t0 = time.perf_counter(); it, _ = ctx.em(80, 10, 0.001, 1e-32); tick("em", t0)
    T["iters"] += it; T["em_device_ms"] += ctx.last_em_ms
    t0 = time.perf_counter(); ctx.get_factors(); tick("get_factors", t0)
    t0 = time.perf_counter(); ctx.stash_topics(0, 1); tick("stash", t0)


def run(devices, lanes, per_lane=4):
    stats = []
    def worker(dev, lane):
        T = collections.defaultdict(float)
        ctx = _lib.Context(dev); ctx.upload_csr(X)
        member(ctx, 1000 + dev * 10 + lane, collections.defaultdict(float))   # warm
        t0 = time.perf_counter()
        for i in range(per_lane):
            member(ctx, 2000 + 100 * dev + 10 * lane + i, T)
        T["wall"] = time.perf_counter() - t0
        stats.append(T); ctx.close()
    th = [threading.Thread(target=worker, args=(d, l)) for d in devices for l in range(lanes)]
    t0 = time.perf_counter()
    [t.start() for t in th]; [t.join() for t in th]
    tot = time.perf_counter() - t0
    agg = collections.defaultdict(float)
    for T in stats:
        for key, v in T.items():
            agg[key] += v
    nm = per_lane * len(th)
    print("devices %s lanes %d: %d members, per member (ms): " % (devices, lanes, nm) +
          ", ".join("%s %.2f" % (key, 1e3 * agg[key] / nm) for key in agg if key not in ("iters", "em_device_ms", "wall")) +
          " | iters %.0f, em device %.2f ms, lane wall/member %.2f ms" % (agg["iters"] / nm, agg["em_device_ms"] / nm, 1e3 * agg["wall"] / nm))

G = _lib.device_count()
run([0], 1); run([0], 2)
if G > 1:
    run([0, 1], 1); run([0, 1], 2)
