"""CPU, world size 2, gloo: the N > 1 host path (launcher plumbing of bench.py, member
sharding and the member-order re-stacking of the gathered topics)."""
import os
import socket
import subprocess
import sys

import numpy as np
from conftest import ROOT

from enstop_b200 import enstop_


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_two_ranks_gloo():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "_dist_worker.py")]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                         timeout=240, env=env, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-3000:]
    assert "MULTIRANK_OK 2" in res.stdout


def test_sharded_em_two_ranks_gloo():
    """Document-sharded EM: local E-step + all-reduced P(w|z) sums == the unsharded oracle."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "_shard_worker.py")]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                         timeout=240, env=env, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-3000:]
    assert "SHARDED_EM_OK 2" in res.stdout


def test_shard_rows():
    from enstop_b200.plsa import shard_rows
    ip = np.concatenate([[0], np.cumsum(np.random.RandomState(0).randint(0, 50, size=1000))])
    for g in (1, 2, 3, 8):
        b = shard_rows(ip, g)
        assert b[0] == 0 and b[-1] == 1000 and len(b) == g + 1
        assert all(b[i] < b[i + 1] for i in range(g))
        sizes = [ip[b[i + 1]] - ip[b[i]] for i in range(g)]
        assert max(sizes) - min(sizes) <= 100          # within two rows of equal entries
    assert shard_rows(np.array([0, 0, 0, 5]), 3) == [0, 1, 2, 3]   # every shard gets a row
    assert shard_rows(np.array([0, 5, 5, 5]), 3) == [0, 1, 2, 3]
    try:
        shard_rows(np.array([0, 1, 2]), 3)
        assert False
    except ValueError:
        pass


def test_reference_arm_other_ranks_do_nothing():
    """bench.py --impl reference under torchrun: rank 0 alone works, the others exit 0."""
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                          "--gpus", "2", "--steps", "1", "--warmup", "0", "--config", "C1"],
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120,
                         env=env, cwd=ROOT)
    assert res.returncode == 0 and res.stdout.strip() == ""


def test_reference_arm_json_contract():
    """bench.py --impl reference on rank 0: one JSON line with the keys the driver reads (same
    metric / unit / config as the GPU arm, `impl`, a cpu_baseline describing this run, an e2e
    object with zero copy bytes) — run here on C1, the reference's own CPU-sized case."""
    import json
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                          "--steps", "2", "--warmup", "1", "--config", "C1"],
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300,
                         env=env, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "nnz*k/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("EM iters/sec") and d["steps"] == 2 and d["n_gpus"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None
    assert d["config"]["workload"].startswith("C1:") and d["config"]["k"] == 10
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_sharding_and_seeds():
    assert enstop_.shard_members(16, 8)[3] == [3, 11]
    assert enstop_.shard_members(5, 2) == [[0, 2, 4], [1, 3]]
    assert sorted(sum(enstop_.shard_members(7, 3), [])) == list(range(7))
    a, b = enstop_.member_seeds(42, 16), enstop_.member_seeds(42, 16)
    assert a == b and len(set(a)) == 16                       # deterministic and distinct
    assert enstop_.member_seeds(43, 16) != a
    idx = enstop_.bootstrap_indices(1000, a[0])
    # enstop_.py:85-87: the member's RandomState draws randint(0, n, size=n)
    assert np.array_equal(idx, np.random.RandomState(a[0]).randint(0, 1000, size=1000))
    stacked = np.arange(5 * 2 * 3, dtype=np.float32).reshape(10, 3)
    shards = enstop_.shard_members(5, 2)                       # gather order: 0,2,4,1,3
    out = enstop_.stack_in_member_order(stacked, shards, 2)
    assert np.array_equal(out[2:4], stacked[6:8])              # member 1 was 4th in the gather
    assert np.array_equal(out[4:6], stacked[2:4])              # member 2 was 2nd


def test_topic_distances_and_combiners():
    """The oracle's all-pairs distances against a direct evaluation, and the host-side
    combiners on precomputed matrices (the product computes them on the GPU)."""
    from oracle import oracle
    rng = np.random.RandomState(0)
    base = rng.dirichlet(np.full(30, 0.3), size=4)
    topics = np.vstack([b * (1 + 0.02 * rng.rand(30)) for b in base for _ in range(6)])
    topics /= topics.sum(axis=1, keepdims=True)
    topics = topics.astype(np.float32)
    H = oracle.all_pairs_hellinger_distance(topics)
    i, j = 3, 17
    direct = np.sqrt(1 - np.sum(np.sqrt(topics[i].astype(float) * topics[j]))
                     / np.sqrt(topics[i].sum(dtype=float) * topics[j].sum(dtype=float)))
    assert np.isclose(H[i, j], direct, atol=1e-6) and np.allclose(np.diag(H), 0, atol=1e-7)
    K = oracle.all_pairs_kl_divergence(topics)
    a, b = topics[i].astype(float), topics[j].astype(float)
    ok = (a > 0) & (b > 0)
    assert np.isclose(K[i, j], np.sum(a[ok] * (np.log2(a[ok]) - np.log2(b[ok]))))
    for name, D in (("hellinger", H), ("kl_divergence", K)):
        stable = enstop_._topic_combiner[name](topics, 3, 4, distances=D)
        assert stable.shape == (4, 30) and stable.dtype == np.float32
        assert np.allclose(stable.sum(axis=1), 1.0, atol=1e-5)
