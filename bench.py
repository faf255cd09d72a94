#!/usr/bin/env python
"""bench.py — EM throughput of the pLSA hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config C2] [--impl reference]

A *step* is one EM iteration (E-step + M-step, plus the periodic log-likelihood test of
plsa_fit_inner, enstop/plsa.py:595-638) over the whole synthetic corpus of the named config
(default C2: 100k docs x 50k terms, ~10M stored entries, k=20 — SURVEY.md §8d recipe).
``value`` is nnz*k*K / device time of K steps with corpus and factors resident in HBM
(CUDA events on the context's stream, max over ranks).  ``e2e`` is the same metric through
the public API (``PLSA.fit`` at N=1, the ensemble member call at N>1) from HOST buffers:
validation, seeded init, H2D of the CSR arrays and factors, term-major build, EM, D2H of
both factors, all inside the timed region.

N > 1 (launched by torch.distributed.run, one rank per GPU): the ensemble path — every rank
fits one bootstrapped member (enstop_.py:84-114) of the same corpus, no data-path
collective; the NCCL gather of the topic matrices (enstop_.py:231) is timed separately.
torch is used for the launcher's rendezvous/barrier only.

``--impl reference`` times the CPU restatement of the reference algorithm (oracle/, C +
OpenMP mirroring numba's prange/serial structure) on the host cores for the same metric.
"""
import argparse
import json
import os
import shutil
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "EM iters/sec (nnz*k/s) at k=%d"
UNIT = "nnz*k/s"
SPEC_HBM_GBS = 8000.0
FALLBACK_HBM_GBS = 6650.0


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.lines = []

    def start(self):
        exe = shutil.which("nvidia-smi")
        if not exe:
            return
        try:
            self.proc = subprocess.Popen(
                [exe, "-i", str(self.device), "--query-gpu=" + self.QUERY,
                 "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0=None, t1=None):
        if not self.proc:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for ts, line in self.lines:
            if t0 is not None and not (t0 - 0.05 <= ts <= t1 + 0.15):
                continue
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                  "sw_power_cap"), parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax),
                "reasons": sorted(reasons), "samples": len(sm)}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    return rank, world, local


class Plumbing:
    """Launcher rendezvous: barrier, max-reduce and a small broadcast.  torch.distributed
    (gloo, CPU tensors) when launched by torchrun; trivial at world size 1."""

    def __init__(self, rank, world):
        self.rank, self.world = rank, world
        self.dist = None
        if world > 1:
            import torch
            import torch.distributed as dist
            self.torch = torch
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group(backend="gloo", rank=rank, world_size=world)
            self.dist = dist

    def barrier(self):
        if self.dist:
            self.dist.barrier()

    def max(self, x):
        if not self.dist:
            return float(x)
        t = self.torch.tensor([float(x)], dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t[0])

    def sum(self, x):
        if not self.dist:
            return float(x)
        t = self.torch.tensor([float(x)], dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t[0])

    def allgather(self, obj):
        if not self.dist:
            return [obj]
        box = [None] * self.world
        self.dist.all_gather_object(box, obj)
        return box

    def bcast_bytes(self, payload):
        if not self.dist:
            return payload
        box = [payload]
        self.dist.broadcast_object_list(box, src=0)
        return box[0]

    def close(self):
        if self.dist:
            self.dist.destroy_process_group()


class IpcExchange:
    """Peer-memory all-reduce of a document-sharded fit between PROCESSES (one rank per GPU):
    every rank exports the CUDA IPC handle of its exchange block through the launcher's
    rendezvous and maps its peers'.  Falls back to ncclAllReduce on every rank if any rank
    cannot."""

    def __init__(self, plumb, no_p2p=False):
        self.plumb, self.no_p2p = plumb, no_p2p

    def __call__(self, c, r):
        from enstop_b200 import _lib
        if self.no_p2p:
            return False
        try:
            c.shard_p2p_prepare()
            mine = c.shard_p2p_export()
        except _lib.PlsaError:
            mine = None
        handles = self.plumb.allgather(mine)
        ok = all(h is not None for h in handles)
        if ok:
            try:
                for p, h in enumerate(handles):
                    if p != r:
                        c.shard_p2p_attach(p, -1, handle=h)
            except _lib.PlsaError:
                ok = False
        ok = all(self.plumb.allgather(ok))
        if not ok:
            c.set_option("p2p", 0)
        return ok

    def finish(self, c):     # importers unmap before any exporter frees its block
        c.shard_p2p_detach()
        self.plumb.barrier()


def _rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


TOL_SHARD_VS_SINGLE = 5e-6      # tests/test_gpu_sharded.py
TOL_VS_REFERENCE = 3e-4         # tests/test_gpu_parity.py TOL_REF_50 (the reference's own drift)


def _exchange_name(world):
    """The peer-memory exchange plsa_em picks (option p2p_two_shot: -1 = two-shot from 4 ranks)."""
    forced = int(os.environ.get("ENSTOP_B200_TWO_SHOT", "-1"))
    if forced > 0 or (forced < 0 and world >= 4):
        return ("two-shot over NVLink peer memory: every rank adds its slice of the rows, "
                "the finished slices are fetched by the peers, fused with the column sums")
    return "one-shot: one kernel per rank reads all partials over NVLink peer memory, fused with the column sums"


def sharded_fit_checks(plumb, rank, world, device):
    """N > 1, before anything is timed: ONE fit of the committed C1 golden corpus with its
    documents sharded over all N ranks — through the peer-memory kernels (one-shot and two-shot exchange) and
    through ncclAllReduce — against the same fit on one GPU and against the reference's own output
    (tests/golden/c1_planted.npz, made by tests/golden/make_golden.py from enstop/plsa.py).
    Returns the list of check records (rank 0; every rank learns the verdict)."""
    import scipy.sparse as sp
    from enstop_b200 import _lib, plsa
    g = np.load(os.path.join(ROOT, "tests", "golden", "c1_planted.npz"))
    X = sp.csr_matrix((g["data"], g["indices"], g["indptr"]), shape=tuple(g["shape"]))
    k, n = int(g["k"]), X.shape[0]
    n_iter = 10
    sw = np.ones(n, dtype=np.float32)
    bounds = plsa.shard_rows(X.indptr, world)
    lo, hi = bounds[rank], bounds[rank + 1]
    single = None
    if rank == 0:
        single = plsa.plsa_fit(X, k, sw, init=(g["pzd0"], g["pwz0"]), n_iter=n_iter, tolerance=0.0,
                               device=device)
    uid = plumb.bcast_bytes(_lib.Comm.unique_id() if rank == 0 else None)
    comm = _lib.Comm(device, world, rank, uid)
    checks = []
    forced = os.environ.get("ENSTOP_B200_TWO_SHOT")
    for path in ("peer-memory kernel, one-shot", "peer-memory kernel, two-shot", "ncclAllReduce"):
        ex = IpcExchange(plumb, no_p2p=(path == "ncclAllReduce"))
        os.environ["ENSTOP_B200_TWO_SHOT"] = "1" if path.endswith("two-shot") else "0"
        try:
            pzd_rows, pwz, info = plsa.plsa_fit_shard(X[lo:hi], k, g["pzd0"][lo:hi], g["pwz0"],
                                                      sw[lo:hi], comm, device, n_iter=n_iter,
                                                      tolerance=0.0, exchange=ex)
        finally:
            if forced is None:
                del os.environ["ENSTOP_B200_TWO_SHOT"]
            else:
                os.environ["ENSTOP_B200_TWO_SHOT"] = forced
        parts = plumb.allgather(pzd_rows)
        pwzs = plumb.allgather(pwz if rank in (0, world - 1) else None)
        used = plumb.allgather(bool(info["p2p"]))
        if rank == 0:
            pzd = np.concatenate(parts)
            rec = {"check": "doc-sharded fit of C1 golden over %d ranks, %s" % (world, path),
                   "exchange_used": path if all(used) else "ncclAllReduce",
                   "iters": int(info["n_iter"]),
                   "components_vs_single_gpu": _rel_l2(pwz, single[1]),
                   "embedding_vs_single_gpu": _rel_l2(pzd, single[0]),
                   "components_vs_reference_golden": _rel_l2(pwz, g["pwz_%d" % n_iter]),
                   "embedding_vs_reference_golden": _rel_l2(pzd, g["pzd_%d" % n_iter]),
                   "ranks_agree_bitwise": bool(np.array_equal(pwzs[0], pwzs[world - 1]))}
            rec["ok"] = bool(rec["iters"] == n_iter and rec["ranks_agree_bitwise"]
                             and rec["components_vs_single_gpu"] < TOL_SHARD_VS_SINGLE
                             and rec["embedding_vs_single_gpu"] < TOL_SHARD_VS_SINGLE
                             and rec["components_vs_reference_golden"] < TOL_VS_REFERENCE
                             and rec["embedding_vs_reference_golden"] < TOL_VS_REFERENCE)
            checks.append(rec)
    comm.close()
    return checks


C4_KW = dict(init="random", n_iter=80, n_iter_per_test=10, tolerance=0.001, e_step_thresh=1e-32,
             bootstrap=True)   # EnsembleTopics' constructor defaults (enstop_.py:709-730)


def run_c4(plumb, rank, world, device, X, k, n_starts=16, seed=42):
    """BASELINE.json config 4: the fan-out of EnsembleTopics(n_components=k, n_starts=16) over
    the N ranks — member r on rank r mod N (enstop_.py:209-217), two lanes per GPU, topics left
    on the device — and the gather of the stacked topic matrices to rank 0 over NCCL
    (enstop_.py:231).  Strong scaling: the 16 members are fixed.  Members 0 and n_starts-1 are
    then fitted alone on rank 0 and must equal their rows of the gathered stack bit for bit."""
    from enstop_b200 import _lib, enstop_
    seeds = enstop_.member_seeds(seed, n_starts)
    shards = enstop_.shard_members(n_starts, world)
    counts = [len(m) for m in shards]
    comm = None
    if world > 1:
        uid = plumb.bcast_bytes(_lib.Comm.unique_id() if rank == 0 else None)
        comm = _lib.Comm(device, world, rank, uid)

    def once():
        plumb.barrier()
        t0 = time.perf_counter()
        ctx, order, used = enstop_.fit_members_on_device(X, k, device, shards[rank], seeds, **C4_KW)
        t_fit = time.perf_counter() - t0
        if comm is not None:
            stacked = comm.gather_topics(ctx, counts, root=0)
        else:
            stacked = _lib.gather_topics([ctx], counts)
        dt = time.perf_counter() - t0
        enstop_.release_member_contexts(used)
        plumb.barrier()
        return stacked, order, plumb.max(dt), plumb.max(t_fit)

    once()          # warm-up: NCCL connects its peers, contexts and pinned staging are created
    runs = [once() for _ in range(3)]
    walls = [r[2] for r in runs]
    stacked, order, wall, fit_wall = runs[int(np.argsort(walls)[1])]     # the median run
    orders = plumb.allgather(order)
    if comm is not None:
        comm.close()
    out = None
    if rank == 0:
        all_topics = enstop_.stack_in_member_order(stacked, orders, k)
        checks = []
        for r in (0, n_starts - 1):
            alone = enstop_.plsa_topics(X, k, random_state=seeds[r], device=device, **C4_KW)
            mine = all_topics[r * k:(r + 1) * k]
            checks.append({"check": "C4 member %d (fitted on rank %d, gathered) == the same member "
                                    "fitted alone" % (r, r % world),
                           "bit_equal": bool(np.array_equal(alone, mine)),
                           "rel_l2": _rel_l2(mine, alone),
                           "ok": bool(np.array_equal(alone, mine))})
        rows_ok = bool(np.allclose(all_topics.sum(axis=1), 1.0, atol=1e-4))
        checks.append({"check": "C4 stack is [%d, %d], rows sum to 1" % all_topics.shape,
                       "ok": bool(all_topics.shape == (n_starts * k, X.shape[1]) and rows_ok)})
        out = {"c4_wall_s": wall, "c4_fit_wall_s": fit_wall, "c4_gather_s": wall - fit_wall,
               "c4_wall_s_all_runs": walls, "c4_statistic": "median of 3 fan-outs + gathers",
               "c4": {"workload": "EnsembleTopics(n_components=%d, n_starts=%d) fan-out + gather on "
                                  "the C2 corpus, members %s per rank, two lanes per GPU"
                                  % (k, n_starts, counts),
                      "scaling": "strong", "member_seeds_from": seed, "kwargs": C4_KW},
               "checks": checks}
    return out


def c1_parity(device):
    """BASELINE.json config 1 from the committed goldens (no reference, no oracle at run time):
    50 EM iterations from the goldens' start through the C ABI; relative L2 of components_
    (P(w|z)) and embedding_ (P(z|d)) between the engine, the reference's own output
    (tests/golden/c1_*.npz, made from enstop/plsa.py by make_golden.py) and the float64-exact
    EM (tests/golden/c1_exact.npz, make_exact.py)."""
    import scipy.sparse as sp
    from enstop_b200 import plsa
    out = {}
    try:
        ex = np.load(os.path.join(ROOT, "tests", "golden", "c1_exact.npz"))
        for tag in ("c1_planted", "c1_zipf"):
            g = np.load(os.path.join(ROOT, "tests", "golden", tag + ".npz"))
            X = sp.csr_matrix((g["data"], g["indices"], g["indptr"]), shape=tuple(g["shape"]))
            sw = np.ones(X.shape[0], dtype=np.float32)
            pzd, pwz = plsa.plsa_fit(X, int(g["k"]), sw, init=(g["pzd0"], g["pwz0"]), n_iter=50,
                                     tolerance=0.0, device=device)
            out[tag] = {
                "components_engine_vs_reference": _rel_l2(pwz, g["pwz_50"]),
                "components_engine_vs_f64_exact": _rel_l2(pwz, ex[tag + "_pwz_50"]),
                "components_reference_vs_f64_exact": _rel_l2(g["pwz_50"], ex[tag + "_pwz_50"]),
                "embedding_engine_vs_reference": _rel_l2(pzd, g["pzd_50"]),
                "embedding_engine_vs_f64_exact": _rel_l2(pzd, ex[tag + "_pzd_50"]),
                "embedding_reference_vs_f64_exact": _rel_l2(g["pzd_50"], ex[tag + "_pzd_50"]),
            }
        out["note"] = ("relative L2 after 50 EM iterations from the same float32 start, k=10, "
                       "2000 x 5000; the reference's serial float32 accumulators put IT ~1e-4 from "
                       "exact arithmetic, which bounds engine_vs_reference from below")
    except Exception as exc:   # a report, never a reason to lose the bench line
        out["error"] = repr(exc)
    return out


def finish_checks(plumb, rank, checks):
    """Every rank learns whether all checks passed; a failure ends the run non-zero."""
    ok = all(c.get("ok", False) for c in checks) if rank == 0 else True
    ok = all(plumb.allgather(ok))
    if not ok:
        if rank == 0:
            print(json.dumps({"parity_checks": checks, "error": "parity check failed"}), file=sys.stderr)
        plumb.close()
        sys.exit(3)


def measured_traffic(config):
    """DRAM bytes per launch of the doc pass from the committed `ncu --set full` capture of
    this same command (profiles/traffic.json, written by scripts/ncu_summary.py --traffic)."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f).get(config)
        return None if t is None else t["doc_pass"]["dram_bytes_read"] + t["doc_pass"]["dram_bytes_write"]
    except Exception:
        return None


def l2_side(config, doc_ms, word_ms):
    """Second roof of the row pass: bytes the SMs pull from L2 per launch (32-byte sectors of
    the committed ncu capture, profiles/traffic.json) over the live launch time.  The L2 of a
    B300-class part delivers ~6300 B/clk to the SMs at best (B300_MICROARCH.md), 12.4 TB/s at
    1965 MHz; the term pass runs close to it, the doc pass at about half."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)[config]
        out = {"peak_gbs_approx": 6300 * 1.965, "source": "lts__t_sectors_srcunit_tex_op_read of "
               + t["term_pass"]["source"] + " x 32 B / live launch time"}
        for name, ms in (("doc_pass", doc_ms), ("term_pass", word_ms)):
            b = t[name]["lts_sectors_read_from_sm"] * 32.0
            out[name] = {"bytes_from_l2": b, "achieved_gbs": b / (ms * 1e-3) / 1e9 if ms > 0 else None,
                         "l1_sector_hit_rate_pct": t[name].get("l1_sector_hit_rate_pct")}
        return out
    except Exception:
        return None


def algorithmic_bytes(n, m, nnz, k):
    """per EM iteration 8*nnz + 4*(n+1) + 8*k*(n+m); E-step reads
    8*nnz + 4*(n+1) + 4*k*(n+m)."""
    return (8 * nnz + 4 * (n + 1) + 8 * k * (n + m), 8 * nnz + 4 * (n + 1) + 4 * k * (n + m))


def cpu_port_run(X, k, n_iter, seed=42):
    """The oracle's float32 restatement of plsa_fit_inner on the host cores.
    Returns seconds for n_iter EM iterations (tolerance 0: all of them run)."""
    from oracle import oracle
    rng = np.random.RandomState(seed)
    n, m = X.shape
    pzd, pwz = oracle.plsa_init_random(n, m, k, rng)
    pzd = np.ascontiguousarray(pzd, dtype=np.float32)
    pwz = np.ascontiguousarray(pwz, dtype=np.float32)
    A = X.tocoo()
    rows = np.ascontiguousarray(A.row, dtype=np.int32)
    cols = np.ascontiguousarray(A.col, dtype=np.int32)
    vals = np.ascontiguousarray(A.data, dtype=np.float32)
    sw = np.ones(n, dtype=np.float32)
    t0 = time.perf_counter()
    iters, _ = oracle.fit_inner(rows, cols, vals, pwz, pzd, sw, n_iter=n_iter,
                                n_iter_per_test=10, tolerance=0.0)
    dt = time.perf_counter() - t0
    assert iters == n_iter
    return dt


def pick_cpu_threads(X, k):
    """Credit the CPU path with its best thread count (the M-step scatter is serial, so more
    threads mostly add OpenMP noise — BASELINE.md §2): probe {1, n/2, n} on a row slice."""
    from oracle import oracle
    ncpu = os.cpu_count() or 1
    probe = X[: max(64, X.shape[0] // 25)]
    best_t, best = 1, None
    for t in sorted({1, max(1, ncpu // 2), ncpu}):
        oracle.set_num_threads(t)
        cpu_port_run(probe, k, 1)
        dt = min(cpu_port_run(probe, k, 2) for _ in range(2))
        if best is None or dt < best:
            best_t, best = t, dt
    return oracle.set_num_threads(best_t)


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU algorithm (oracle port) on the host cores."""
    if rank != 0:
        return
    from enstop_b200 import synth
    from oracle import oracle
    oracle.build()
    cfg = synth.CONFIGS[args.config]
    X, info = synth.make_config(args.config, return_info=True)
    n, m = X.shape
    k = cfg["k"]
    budget_s = 150.0
    cores = pick_cpu_threads(X, k)
    # bound the sample: probe one iteration on a 1/16 row slice, then size the slice
    probe = X[: max(64, n // 16)]
    cpu_port_run(probe, k, 1)  # warm (page faults, OpenMP pool)
    per_iter_probe = cpu_port_run(probe, k, 2) / 2.0
    per_iter_full = per_iter_probe * (X.nnz / max(1, probe.nnz))
    total_iters = args.steps + args.warmup
    frac = min(1.0, budget_s / max(1e-9, per_iter_full * total_iters))
    rows = n if frac >= 1.0 else max(64, int(n * frac))
    Xs = X[:rows]
    if args.warmup > 0:
        cpu_port_run(Xs, k, min(args.warmup, 2))
    dt = cpu_port_run(Xs, k, args.steps)
    value = Xs.nnz * k * args.steps / dt
    sample = ("%d EM iterations of plsa_fit_inner (oracle C port: OpenMP E-step/log-likelihood, "
              "serial M-step scatter as enstop/plsa.py:182) on the first %d of %d documents "
              "(%d of %d stored entries) of %s" % (args.steps, rows, n, Xs.nnz, X.nnz, args.config))
    line = {
        "impl": "reference", "metric": METRIC % k, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.config, cfg, info), "n_docs": n, "n_terms": m,
                   "nnz": int(X.nnz), "k": k, "sample_nnz": int(Xs.nnz)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def make_config_once(name, plumb, rank, world):
    """The synthetic corpus of a config, generated by rank 0 only and handed to the other
    ranks of the box through /dev/shm (the generator sorts up to 3e8 keys; N copies of that at
    once would be most of the run)."""
    from enstop_b200 import synth
    import scipy.sparse as sp
    cache = os.environ.get("ENSTOP_B200_CORPUS_CACHE")   # A/B scripts: reuse a generated corpus
    if world == 1 and cache:
        path = os.path.join(cache, "enstop_b200_%s.npz" % name)
        if os.path.exists(path):
            with np.load(path) as z:
                X = sp.csr_matrix((z["data"], z["indices"], z["indptr"]), shape=tuple(z["shape"]))
                info = json.loads(str(z["info"]))
            X.has_sorted_indices = True
            return X, info
        X, info = synth.make_config(name, return_info=True)
        np.savez(path, data=X.data, indices=X.indices, indptr=X.indptr, shape=np.array(X.shape),
                 info=np.array(json.dumps(info)))
        return X, info
    if world == 1:
        return synth.make_config(name, return_info=True)
    path = "/dev/shm/enstop_b200_%s_%s.npz" % (name, os.environ.get("MASTER_PORT", "0"))
    if rank == 0:
        X, info = synth.make_config(name, return_info=True)
        np.savez(path, data=X.data, indices=X.indices, indptr=X.indptr, shape=np.array(X.shape),
                 info=np.array(json.dumps(info)))
        plumb.barrier()
        plumb.barrier()
        os.remove(path)
        return X, info
    plumb.barrier()
    with np.load(path) as z:
        X = sp.csr_matrix((z["data"], z["indices"], z["indptr"]), shape=tuple(z["shape"]))
        info = json.loads(str(z["info"]))
    X.has_sorted_indices = True
    plumb.barrier()
    return X, info


def run_shard(args, plumb, rank, world, device):
    """--mode shard, N > 1: one fit of the whole corpus, documents sharded over the ranks
    (include/plsa_b200.h plsa_set_shard); strong scaling."""
    import types
    from enstop_b200 import _lib, plsa, synth
    cfg = synth.CONFIGS[args.config]
    k = cfg["k"]
    X, info = make_config_once(args.config, plumb, rank, world)
    n, m = X.shape
    bounds = plsa.shard_rows(X.indptr, world)
    lo, hi = bounds[rank], bounds[rank + 1]
    Xs = X[lo:hi]
    uid = plumb.bcast_bytes(_lib.Comm.unique_id() if rank == 0 else None)
    comm = _lib.Comm(device, world, rank, uid)
    rng = np.random.RandomState(42)
    pzd0, pwz0 = plsa.plsa_init(types.SimpleNamespace(shape=(n, m)), k, "random", rng)
    ctx = _lib.Context(device)
    ctx.upload_csr(Xs)
    ctx.set_shard(comm)
    ctx.set_factors(pzd0[lo:hi].astype(np.float32), pwz0.astype(np.float32))
    ctx.set_sample_weight(None)

    exchange = IpcExchange(plumb, args.no_p2p)
    p2p = exchange(ctx, rank)
    ctx.em(args.warmup, n_iter_per_test=10, tolerance=0.0)
    sampler = ClockSampler(device)
    sampler.start()
    time.sleep(0.25)
    plumb.barrier()
    launches0 = ctx.launches
    t0 = time.time()
    iters, trace = ctx.em(args.steps, n_iter_per_test=10, tolerance=0.0)
    em_ms_local = ctx.last_em_ms
    t1 = time.time()
    plumb.barrier()
    launches = plumb.sum(ctx.launches - launches0)
    clocks = sampler.stop(t0, t1)
    assert iters == args.steps
    em_ms = plumb.max(em_ms_local)
    value = float(X.nnz) * k * args.steps / (em_ms * 1e-3)
    ctx.set_profiling(True)
    ctx.em(args.profile_iters, n_iter_per_test=10, tolerance=0.0)
    prof = ctx.profile()
    ctx.set_profiling(False)
    kernel_ms = {s: plumb.max(prof[s]["ms"] / args.profile_iters) for s in prof}

    # end to end: the public call, host buffers, every rank fits its shard (rank 0 reports)
    sw = np.ones(hi - lo, dtype=np.float32)

    def e2e_call(n_iter):
        return plsa.plsa_fit_shard(Xs, k, pzd0[lo:hi], pwz0, sw, comm, device, n_iter=n_iter,
                                   tolerance=0.0, exchange=exchange)
    e2e_call(3)
    runs = []
    for _ in range(args.e2e_repeats):
        plumb.barrier()
        w0 = time.perf_counter()
        e2e_call(args.steps)
        dt = time.perf_counter() - w0
        plumb.barrier()
        runs.append(plumb.max(dt))
    e2e_s = statistics.median(runs)
    if p2p:
        exchange.finish(ctx)
    ctx.set_shard(None)
    ctx.close()
    comm.close()
    if rank == 0:
        b_iter, _ = algorithmic_bytes(n, m, X.nnz, k)
        peak, peak_src = load_peaks()
        line = {
            "metric": METRIC % k, "value": value, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": em_ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": workload_name(args.config, cfg, info), "n_docs": n, "n_terms": m,
                       "nnz": int(X.nnz), "k": k, "shard_bounds": bounds,
                       "parallelism": "ONE fit, documents sharded over %d GPUs, raw P(w|z) sums "
                                      "all-reduced once per EM iteration (%s)"
                                      % (world, _exchange_name(world) if p2p else "ncclAllReduce"),
                       "ll_first_last": [float(trace[0]), float(trace[-1])]},
            "roofline": {"bound": "hbm", "achieved": b_iter / (em_ms / args.steps * 1e-3) / 1e9,
                         "peak": peak * world, "unit": "GB/s",
                         "frac": b_iter / (em_ms / args.steps * 1e-3) / 1e9 / (peak * world),
                         "peak_source": peak_src + " x n_gpus", "traffic": None,
                         "kernel": "whole EM iteration, aggregate over the GPUs",
                         "kernel_ms_per_iter_max_over_ranks": kernel_ms},
            "cpu_baseline": None,
            "e2e": {"value": float(X.nnz) * k * args.steps / e2e_s, "unit": UNIT,
                    "h2d_bytes_per_step": (8 * X.nnz + 4 * (n + world) + 4 * k * (n + m * world)) / args.steps,
                    "d2h_bytes_per_step": 4 * k * (n + m * world) / args.steps, "seconds": e2e_s,
                    "call": "plsa_fit_shard per rank (upload shard, set factors, EM, download)"},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        print(json.dumps(line))


def workload_name(name, cfg, info):
    return ("%s: PLSA(n_components=%d) EM on synthetic Zipf(s=1) token-sampled CSR %dx%d, "
            "%d stored entries (seed %d, %d tokens)"
            % (name, cfg["k"], cfg["n"], cfg["m"], info["nnz"], info["seed"], info["tokens"]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default="C2", choices=["C1", "C2", "C3", "C5"])
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mode", default="ensemble", choices=["ensemble", "shard"],
                    help="N > 1: 'ensemble' = one bootstrapped member per GPU (weak scaling, the "
                         "BASELINE.json multi-GPU config); 'shard' = ONE fit, documents sharded "
                         "over the GPUs, P(w|z) all-reduced per EM iteration (strong scaling)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-p2p", action="store_true", help="--mode shard: ncclAllReduce instead of "
                                                          "the peer-memory all-reduce kernel")
    ap.add_argument("--pageable", action="store_true", help="e2e: keep the CSR arrays in pageable memory")
    ap.add_argument("--no-checks", action="store_true", help="N > 1: skip the sharded-fit parity checks")
    ap.add_argument("--no-c4", action="store_true", help="skip the 16-member ensemble (config 4) leg")
    ap.add_argument("--cpu-iters", type=int, default=10)
    ap.add_argument("--profile-iters", type=int, default=20)
    ap.add_argument("--e2e-repeats", type=int, default=3)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank, world, local = dist_env()
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    from enstop_b200 import _lib, plsa, synth
    if _lib.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    plumb = Plumbing(rank, world)
    device = local % _lib.device_count()
    if world > 1 and args.mode == "shard":
        run_shard(args, plumb, rank, world, device)
        plumb.close()
        return
    cfg = synth.CONFIGS[args.config]
    k = cfg["k"]
    # ---- N > 1: parity of the multi-GPU paths, before anything is timed ---------------------
    parity_checks = []
    if world > 1 and not args.no_checks:
        parity_checks += sharded_fit_checks(plumb, rank, world, device) or []
        finish_checks(plumb, rank, parity_checks)
    X, info = make_config_once(args.config, plumb, rank, world)
    n, m = X.shape
    peak, peak_src = load_peaks()

    # N > 1: this rank's ensemble member = bootstrap resample of the corpus (enstop_.py:86-88)
    member_seed = 1000 + rank
    ctx = _lib.Context(device)
    ctx.upload_csr(X)
    if world > 1:
        idx = np.random.RandomState(member_seed).randint(0, n, size=n)
        ctx.bootstrap(idx)
    n_fit, m_fit, nnz_fit = ctx.shape
    rng = np.random.RandomState(42 + rank)
    import types
    pzd0, pwz0 = plsa.plsa_init(types.SimpleNamespace(shape=(n_fit, m_fit)), k, "random", rng)
    pzd0 = pzd0.astype(np.float32)
    pwz0 = pwz0.astype(np.float32)
    ctx.set_factors(pzd0, pwz0)
    ctx.set_sample_weight(None)

    # ---- device-timed K steps (inputs resident in HBM) -----------------------------------
    ctx.em(args.warmup, n_iter_per_test=10, tolerance=0.0)          # warm-up, untimed
    sampler = ClockSampler(device)
    sampler.start()
    time.sleep(0.25)
    plumb.barrier()
    launches0 = ctx.launches
    t0 = time.time()
    iters, trace = ctx.em(args.steps, n_iter_per_test=10, tolerance=0.0)
    em_ms_local = ctx.last_em_ms
    t1 = time.time()
    plumb.barrier()
    launches = ctx.launches - launches0
    clocks = sampler.stop(t0, t1)
    assert iters == args.steps
    em_ms = plumb.max(em_ms_local)
    total_units = plumb.sum(float(nnz_fit) * k * args.steps)
    value = total_units / (em_ms * 1e-3)

    # ---- per-kernel durations, live CUDA events (profiling mode adds event overhead, so it
    # is a separate short run; the kernels and their inputs are the same) ------------------
    ctx.set_profiling(True)
    ctx.em(args.profile_iters, n_iter_per_test=10, tolerance=0.0)
    prof = ctx.profile()
    ctx.set_profiling(False)
    b_iter, b_e = algorithmic_bytes(n_fit, m_fit, nnz_fit, k)
    doc_ms = prof["doc_pass"]["ms"] / max(1, prof["doc_pass"]["launches"])
    word_ms = prof["word_pass"]["ms"] / max(1, prof["word_pass"]["launches"])
    achieved = b_e / (doc_ms * 1e-3) / 1e9
    tiled = prof["doc_head"]["launches"] > 0     # the library chose the tiled passes (>= 64 M entries)
    roofline = {
        "bound": "hbm",
        "kernel": ("doc pass = tile_pass_kernel (head: shared-memory tile, TMA-staged) + row_pass_kernel<doc> "
                   "(tail), E-step + P(z|d) M-step") if tiled else "row_pass_kernel<doc> (E-step + P(z|d) M-step)",
        "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "peak_source": peak_src,
        "traffic": measured_traffic(args.config) if world == 1 and not tiled else None,
        "algorithmic_bytes_per_launch": b_e, "avg_launch_ms": doc_ms,
        "frac_of_spec_8TBs": achieved / SPEC_HBM_GBS,
        "word_pass": {"avg_launch_ms": word_ms,
                      "achieved": b_e / (word_ms * 1e-3) / 1e9 if word_ms > 0 else None},
        "iteration": {"algorithmic_bytes": b_iter, "ms": em_ms_local / args.steps,
                      "achieved": b_iter / (em_ms_local / args.steps * 1e-3) / 1e9,
                      "frac": b_iter / (em_ms_local / args.steps * 1e-3) / 1e9 / peak},
        "kernel_ms_per_iter": {s: prof[s]["ms"] / args.profile_iters for s in prof},
        "l2": l2_side(args.config, doc_ms, word_ms) if world == 1 and not tiled else None,
    }

    # ---- end to end through the public API, HOST buffers ---------------------------------
    sw = np.ones(n_fit, dtype=np.float32)
    # the step's inputs (the three CSR arrays) start in page-locked HOST memory, as the bench
    # contract has it: the upload inside the timed call is then one DMA transfer per array
    # (pageable arrays work the same way through a staged copy; --pageable times that)
    to_host = (lambda M: M) if args.pageable else _lib.pinned_csr
    if world == 1:
        def e2e_call(n_iter):
            model = plsa.PLSA(n_components=k, n_iter=n_iter, tolerance=0.0, random_state=42,
                              device=device)
            model.fit(Xe)
            return model
        Xe = to_host(X)
    else:
        Xe = to_host(X[np.random.RandomState(member_seed).randint(0, n, size=n)])

        def e2e_call(n_iter):
            return plsa.plsa_fit(Xe, k, sw, n_iter=n_iter, tolerance=0.0,
                                 random_state=42 + rank, device=device)
    plumb.barrier()                                                   # first call: a new pooled
    w0 = time.perf_counter()                                          # context allocates its pinned
    e2e_call(args.steps)                                              # staging and device buffers
    e2e_first_s = plumb.max(time.perf_counter() - w0)
    e2e_runs = []
    for _ in range(args.e2e_repeats):                                 # each: one whole K-step fit
        plumb.barrier()
        w0 = time.perf_counter()
        e2e_call(args.steps)
        e2e_s_local = time.perf_counter() - w0
        plumb.barrier()
        e2e_runs.append(plumb.max(e2e_s_local))
    e2e_s = statistics.median(e2e_runs)
    e2e_value = plumb.sum(float(Xe.nnz) * k * args.steps) / e2e_s
    h2d = (4 * (n_fit + 1) + 8 * Xe.nnz + 4 * k * (n_fit + m_fit) + 4 * n_fit) / args.steps
    d2h = (4 * k * (n_fit + m_fit)) / args.steps

    # ---- ensemble gather over NCCL (N > 1), outside the EM timing -------------------------
    gather_ms = None
    if world > 1:
        uid = plumb.bcast_bytes(_lib.Comm.unique_id() if rank == 0 else None)
        comm = _lib.Comm(device, world, rank, uid)
        ctx.stash_topics(0, 1)
        comm.gather_topics(ctx, [1] * world, root=0)     # first use: NCCL connects its peers
        plumb.barrier()
        g0 = time.perf_counter()
        stacked = comm.gather_topics(ctx, [1] * world, root=0)
        plumb.barrier()
        gather_ms = (time.perf_counter() - g0) * 1e3
        if rank == 0:
            assert stacked.shape == (world * k, m_fit)
            assert np.allclose(stacked.sum(axis=1), 1.0, atol=1e-4)
        comm.close()

    # ---- CPU baseline beside it (rank 0, N = 1 only) ----------------------------------------
    cpu = None
    if world == 1 and rank == 0 and not args.no_cpu_baseline:
        from oracle import oracle
        oracle.build()
        cores = pick_cpu_threads(X, k)
        dt = cpu_port_run(X, k, args.cpu_iters)
        cpu = {"value": X.nnz * k * args.cpu_iters / dt, "unit": UNIT,
               "cores": cores, "host_cores": os.cpu_count() or 1, "kind": "port",
               "ms_per_iter": 1e3 * dt / args.cpu_iters,
               "sample": "%d EM iterations of plsa_fit_inner on the full %s corpus (oracle C "
                         "port: OpenMP E-step/log-likelihood, serial M-step scatter as "
                         "enstop/plsa.py:182)" % (args.cpu_iters, args.config)}

    ctx.close()
    _lib.release_device_memory()

    # ---- BASELINE.json config 4: the 16-member ensemble over the N ranks (strong scaling) ---
    c4 = None
    if args.config == "C2" and not args.no_c4:
        c4 = run_c4(plumb, rank, world, device, X, k)
        if rank == 0:
            parity_checks += c4.pop("checks")
        finish_checks(plumb, rank, parity_checks)
    if rank == 0:
        line = {
            "metric": METRIC % k, "value": value, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": em_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": workload_name(args.config, cfg, info),
                       "n_docs": n, "n_terms": m, "nnz": int(X.nnz), "k": k,
                       "parallelism": "single fit" if world == 1 else
                       "ensemble members sharded one per GPU (bootstrap resample per rank)",
                       "l2": "working set (doc-major + term-major CSR + factors, %.0f MB) exceeds "
                             "the 126 MB L2; no flush" % ((16 * nnz_fit + 16 * k * (n_fit + m_fit)) / 1e6),
                       "em_iters_per_s": args.steps / (em_ms * 1e-3),
                       "ll_first_last": [float(trace[0]), float(trace[-1])]},
            "roofline": roofline, "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "seconds": e2e_s,
                    "seconds_all_runs": e2e_runs, "statistic": "median of %d fits" % len(e2e_runs),
                    "first_call_seconds": e2e_first_s,
                    "first_call_note": "the same call the first time in the process after CUDA start-up: "
                                       "it creates the pooled context (pinned staging, device buffers, "
                                       "sort scratch); `seconds` are the calls after it",
                    "call": "PLSA(n_components=%d, n_iter=%d, tolerance=0).fit(X)" % (k, args.steps)
                    if world == 1 else "plsa_fit(bootstrap member) per rank",
                    "host_buffers": "pageable numpy arrays (staged upload)" if args.pageable else
                    "CSR arrays in page-locked host memory (enstop_b200._lib.pinned_csr); factors are "
                    "drawn on the host inside the call"},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        if gather_ms is not None:
            line["ensemble_gather_ms"] = gather_ms
        if c4 is not None:
            line.update(c4)
        line["parity_checks"] = parity_checks
        if world == 1:
            line["parity"] = c1_parity(device)
        print(json.dumps(line))
    plumb.close()


if __name__ == "__main__":
    main()
