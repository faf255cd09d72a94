"""ctypes binding of ``libplsa_b200.so`` (C ABI: ``include/plsa_b200.h``).

The shared library is built in-tree by :func:`build` (nvcc, sm_100a only) and loaded with
``ctypes.CDLL``, which releases the GIL for the duration of every call — the same property
the reference relies on when it runs ``nogil`` numba kernels from ensemble worker threads
(enstop/enstop_.py:209-217).

There is no CPU fallback: if the library cannot be loaded or no CUDA device is present,
every entry point raises.
"""
import ctypes
import os
import shutil
import subprocess
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
SO_PATH = os.environ.get("ENSTOP_B200_LIB") or os.path.join(_HERE, "libplsa_b200.so")
SOURCES = [os.path.join(_HERE, "csrc", "plsa_b200.cu"), os.path.join(_HERE, "csrc", "host_init.cpp")]
HEADERS = [os.path.join(_HERE, "csrc", "plsa_kernels.cuh"),
           os.path.join(ROOT, "include", "plsa_b200.h")]

NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-shared", "-Xcompiler", "-fPIC,-fvisibility=hidden,-mavx2", "-diag-suppress", "550"]

PLSA_OK, PLSA_EINVAL, PLSA_ECUDA, PLSA_ENOMEM, PLSA_ENCCL = 0, 1, 2, 3, 4
ABI_VERSION = 200   # plsa_version() of the library this binding was written against
PROF_SLOTS = ("doc_pass", "word_pass", "fixup", "normalize", "loglik", "doc_head", "term_head")

_i32p = ctypes.POINTER(ctypes.c_int32)
_i64p = ctypes.POINTER(ctypes.c_int64)
_f32p = ctypes.POINTER(ctypes.c_float)
_f64p = ctypes.POINTER(ctypes.c_double)
_ctx = ctypes.c_void_p
_i32, _i64, _f32, _f64 = ctypes.c_int32, ctypes.c_int64, ctypes.c_float, ctypes.c_double

# name -> (restype, argtypes); every symbol include/plsa_b200.h declares
SIGNATURES = {
    "plsa_version": (ctypes.c_int, []),
    "plsa_device_count": (ctypes.c_int, [ctypes.POINTER(ctypes.c_int)]),
    "plsa_last_error": (ctypes.c_char_p, [_ctx]),
    "plsa_ctx_create": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(_ctx)]),
    "plsa_ctx_destroy": (ctypes.c_int, [_ctx]),
    "plsa_upload_csr": (ctypes.c_int, [_ctx, _i32p, _i32p, _f32p, _i64, _i64, _i64]),
    "plsa_upload_csr_typed": (ctypes.c_int, [_ctx, _i32p, _i32p, ctypes.c_void_p, _i32, _i64, _i64,
                                             _i64]),
    "plsa_upload_coo": (ctypes.c_int, [_ctx, _i32p, _i32p, _f32p, _i64, _i64, _i64]),
    "plsa_bootstrap": (ctypes.c_int, [_ctx, _i32p, _i64]),
    "plsa_corpus_shape": (ctypes.c_int, [_ctx, _i64p, _i64p, _i64p]),
    "plsa_set_factors": (ctypes.c_int, [_ctx, _f32p, _f32p, _i32]),
    "plsa_set_sample_weight": (ctypes.c_int, [_ctx, _f32p]),
    "plsa_pinned_factors": (ctypes.c_int, [_ctx, _i64, _i64, _i32, ctypes.POINTER(_f32p),
                                           ctypes.POINTER(_f32p)]),
    "plsa_host_alloc": (ctypes.c_int, [_i64, ctypes.POINTER(ctypes.c_void_p)]),
    "plsa_host_free": (ctypes.c_int, [ctypes.c_void_p]),
    "plsa_get_factors": (ctypes.c_int, [_ctx, _f32p, _f32p]),
    "plsa_stash_topics": (ctypes.c_int, [_ctx, _i32, _i32]),
    "plsa_topics_device": (ctypes.c_int, [_ctx, ctypes.POINTER(ctypes.c_void_p), _i64p]),
    "plsa_em": (ctypes.c_int, [_ctx, _i32, _i32, _f64, _f32, _i32, _i32, _i32p, _f64p, _i32,
                               _i32p]),
    "plsa_prepare": (ctypes.c_int, [_ctx, _i32, _i32]),
    "plsa_log_likelihood": (ctypes.c_int, [_ctx, _f64p]),
    "plsa_last_em_ms": (ctypes.c_int, [_ctx, _f32p]),
    "plsa_set_profiling": (ctypes.c_int, [_ctx, _i32]),
    "plsa_get_profile": (ctypes.c_int, [_ctx, _f64p, _i64p]),
    "plsa_launch_count": (ctypes.c_int, [_ctx, _i64p]),
    "plsa_set_option": (ctypes.c_int, [_ctx, ctypes.c_char_p, _i64]),
    "plsa_plan_items": (ctypes.c_int, [_i32p, _i64, _i64, _i32, _i64, _i64p, _i32p, _i32p,
                                       _i32p, _i32p, _i64p, _i32p, _i32p]),
    "plsa_last_distances_ms": (ctypes.c_int, [ctypes.POINTER(ctypes.c_float)]),
    "plsa_gathered_distances": (ctypes.c_int, [_ctx, _i32, _f64p, _i64p]),
    "plsa_debug_items": (ctypes.c_int, [_ctx, _i32, _i64, _i64p, _i32p, _i32p, _i32p, _i32p, _i64p,
                                        _i32p, _i32p, _i64p, _i32p, _i32p, _i64]),
    "plsa_host_random_rows": (ctypes.c_int, [ctypes.POINTER(ctypes.c_uint32), _i32p, _i64, _i64, _f32p,
                                             _f64p]),
    "plsa_b200_fit_inner": (ctypes.c_int, [_i32p, _i32p, _f32p, _i64, _f32p, _f32p, _f32p, _i64,
                                           _i64, _i32, _i32, _i32, _f64, _f32, _i32, _i32, _i32p]),
    "plsa_b200_refit_inner": (ctypes.c_int, [_i32p, _i32p, _f32p, _i64, _f32p, _f32p, _f32p,
                                             _i64, _i64, _i32, _i32, _i32, _f64, _f32, _i32,
                                             _i32p]),
    "plsa_gather_topics": (ctypes.c_int, [ctypes.POINTER(_ctx), _i32, _i32p, _f32p]),
    "plsa_nccl_unique_id": (ctypes.c_int, [ctypes.c_char_p]),
    "plsa_comm_create": (ctypes.c_int, [ctypes.c_int, _i32, _i32, ctypes.c_char_p,
                                        ctypes.POINTER(ctypes.c_void_p)]),
    "plsa_comm_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "plsa_comm_abort": (ctypes.c_int, [ctypes.c_void_p]),
    "plsa_comm_gather_topics": (ctypes.c_int, [ctypes.c_void_p, _ctx, _i32p, _i32, _f32p]),
    "plsa_gather_warmup": (ctypes.c_int, [_i32p, _i32]),
    "plsa_stash_append": (ctypes.c_int, [_ctx, _ctx, _i32, _i32]),
    "plsa_topic_distances": (ctypes.c_int, [_i32, _f32p, _i64, _i64, _i32, _f64p]),
    "plsa_set_shard": (ctypes.c_int, [_ctx, ctypes.c_void_p]),
    "plsa_shard_p2p_prepare": (ctypes.c_int, [_ctx, ctypes.POINTER(ctypes.c_uint64), _i64p]),
    "plsa_shard_p2p_export": (ctypes.c_int, [_ctx, ctypes.c_char_p]),
    "plsa_shard_p2p_attach": (ctypes.c_int, [_ctx, _i32, _i32, ctypes.c_uint64, ctypes.c_char_p]),
    "plsa_shard_p2p_detach": (ctypes.c_int, [_ctx]),
}

_lib = None
_lock = threading.Lock()


class PlsaError(RuntimeError):
    """A call into libplsa_b200.so failed (message from plsa_last_error)."""


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise PlsaError("nvcc not found; cannot build libplsa_b200.so")


def needs_build():
    if not os.path.exists(SO_PATH):
        return True
    t = os.path.getmtime(SO_PATH)
    return any(os.path.getmtime(p) > t for p in SOURCES + HEADERS)


def build(force=False, verbose=False):
    """Compile the CUDA extension in-tree for sm_100a.  Returns the path of the .so."""
    if not force and not needs_build():
        return SO_PATH
    cmd = [nvcc_path()] + NVCC_FLAGS + ["-o", SO_PATH] + SOURCES + ["-ldl"]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env)
    if res.returncode != 0:
        raise PlsaError("nvcc failed:\n" + res.stdout)
    if verbose:
        print(res.stdout)
    return SO_PATH


def lib():
    """The loaded library.  Raises PlsaError if it is missing — never falls back to CPU."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(SO_PATH):
            raise PlsaError(
                "libplsa_b200.so is not built (run `python -c 'import __graft_entry__ as g; "
                "g.build()'` or enstop_b200._lib.build()); there is no CPU fallback")
        try:
            L = ctypes.CDLL(SO_PATH)
        except OSError as exc:
            raise PlsaError("cannot load %s: %s" % (SO_PATH, exc)) from exc
        try:
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(L, name)
                fn.restype = res
                fn.argtypes = args
        except AttributeError as exc:
            raise PlsaError("%s is older than this binding (%s): rebuild it with "
                            "enstop_b200._lib.build(force=True)" % (SO_PATH, exc)) from exc
        if L.plsa_version() != ABI_VERSION:
            raise PlsaError("%s has ABI version %d, this binding expects %d: rebuild it with "
                            "enstop_b200._lib.build(force=True)" % (SO_PATH, L.plsa_version(), ABI_VERSION))
        if not os.environ.get("ENSTOP_B200_LIB") and needs_build():
            import warnings
            warnings.warn("libplsa_b200.so is older than its sources (csrc/, include/): the loaded "
                          "library may not match them; run enstop_b200._lib.build()", RuntimeWarning)
        _lib = L
    return _lib


def check(rc, ctx=None):
    if rc != PLSA_OK:
        msg = lib().plsa_last_error(ctx)
        msg = msg.decode("utf-8", "replace") if msg else "unknown error"
        kind = {PLSA_EINVAL: "invalid argument", PLSA_ECUDA: "CUDA error",
                PLSA_ENOMEM: "out of memory", PLSA_ENCCL: "NCCL error"}.get(rc, "error %d" % rc)
        raise PlsaError("libplsa_b200: %s: %s" % (kind, msg))


def random_rows(rng, rows, cols, want_f64=False, out=None):
    """``rng.rand(rows, cols)`` L1-row-normalised in float64 and cast to float32, computed by
    the library from the RandomState's own MT19937 state (which is advanced exactly as numpy
    would).  Returns None when ``rng`` is not a legacy MT19937 RandomState (caller falls back
    to numpy).  ``out``: a C-contiguous float32 [rows, cols] array to fill (e.g. a view of the
    context's pinned staging, Context.pinned_factors)."""
    owner = np.random if rng is np.random else rng
    get_state = getattr(owner, "get_state", None)
    if get_state is None or not (owner is np.random or isinstance(owner, np.random.RandomState)):
        return None
    state = get_state()
    if not isinstance(state, tuple) or state[0] != "MT19937":
        return None
    key = np.ascontiguousarray(state[1], dtype=np.uint32).copy()
    pos = _i32(int(state[2]))
    if out is None:
        out = np.empty((rows, cols), dtype=np.float32)
    elif out.shape != (rows, cols) or out.dtype != np.float32 or not out.flags.c_contiguous:
        raise ValueError("random_rows: out must be a C-contiguous float32 [rows, cols] array")
    out64 = np.empty((rows, cols), dtype=np.float64) if want_f64 else None
    check(lib().plsa_host_random_rows(key.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)),
                                      ctypes.byref(pos), rows, cols, _ptr(out, _f32p),
                                      _ptr(out64, _f64p)))
    owner.set_state(("MT19937", key, pos.value) + tuple(state[3:]))
    return (out, out64) if want_f64 else out


def plan_items(indptr, chunk, align=4):
    """Host-only: the work items a row pass over a CSR with these row pointers launches, in
    launch order (plsa_plan_items).  Returns a dict of arrays start/row/len/slot/skip plus
    n_split and n_slots."""
    indptr = _as(indptr, np.int32)
    rows = indptr.shape[0] - 1
    n = _i64(0)
    ns, nl = _i32(0), _i32(0)
    L = lib()
    check(L.plsa_plan_items(_ptr(indptr, _i32p), rows, int(chunk), int(align), 0,
                            None, None, None, None, None, ctypes.byref(n), ctypes.byref(ns),
                            ctypes.byref(nl)))
    out = dict(start=np.empty(n.value, np.int64), row=np.empty(n.value, np.int32),
               len=np.empty(n.value, np.int32), slot=np.empty(n.value, np.int32),
               skip=np.empty(n.value, np.int32))
    check(L.plsa_plan_items(_ptr(indptr, _i32p), rows, int(chunk), int(align), n.value,
                            _ptr(out["start"], _i64p), _ptr(out["row"], _i32p),
                            _ptr(out["len"], _i32p), _ptr(out["slot"], _i32p),
                            _ptr(out["skip"], _i32p), ctypes.byref(n), ctypes.byref(ns),
                            ctypes.byref(nl)))
    out["n_split"], out["n_slots"] = ns.value, nl.value
    return out


def pinned_empty(shape, dtype):
    """An uninitialised numpy array in page-locked host memory (freed with the array): uploads
    from it are plain DMA transfers."""
    import weakref
    dtype = np.dtype(dtype)
    count = int(np.prod(shape))
    ptr = ctypes.c_void_p()
    check(lib().plsa_host_alloc(count * dtype.itemsize, ctypes.byref(ptr)))
    buf = (ctypes.c_char * max(1, count * dtype.itemsize)).from_address(ptr.value)
    arr = np.frombuffer(buf, dtype=dtype, count=count).reshape(shape)
    weakref.finalize(buf, lib().plsa_host_free, ptr)
    return arr


def pinned_csr(X):
    """A copy of a scipy CSR matrix whose three arrays live in page-locked memory."""
    import scipy.sparse as sp
    X = X.tocsr()
    parts = []
    for a in (X.data, X.indices, X.indptr):
        p = pinned_empty(a.shape, a.dtype)
        p[...] = a
        parts.append(p)
    out = sp.csr_matrix(tuple(parts), shape=X.shape, copy=False)
    out.has_sorted_indices = X.has_sorted_indices
    return out


def device_count():
    n = ctypes.c_int(0)
    rc = lib().plsa_device_count(ctypes.byref(n))
    return n.value if rc == PLSA_OK else 0


def _ptr(a, ct):
    return a.ctypes.data_as(ct) if a is not None else None


def _as(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


class Context:
    """One corpus resident on one GPU (opaque ``plsa_ctx``).  Not thread-safe; use one per
    worker thread, as the reference's ensemble does with its numba kernels."""

    def __init__(self, device=0):
        self._h = _ctx()
        self._L = lib()
        check(self._L.plsa_ctx_create(int(device), ctypes.byref(self._h)))
        self.device = int(device)
        self.k = 0
        if os.environ.get("ENSTOP_B200_CHUNK"):     # work-item length experiments
            self.set_option("chunk", int(os.environ["ENSTOP_B200_CHUNK"]))
        if os.environ.get("ENSTOP_B200_TEXTURE"):   # 0: gather with LDG instead of the texture pipe
            self.set_option("texture", int(os.environ["ENSTOP_B200_TEXTURE"]))
        if os.environ.get("ENSTOP_B200_VEC"):       # 0: unaligned items, one 8-byte load per entry
            self.set_option("vec_entries", int(os.environ["ENSTOP_B200_VEC"]))

    def close(self):
        if self._h:
            self._L.plsa_ctx_destroy(self._h)
            self._h = _ctx()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- corpus -----------------------------------------------------------------------
    _DTYPES = {np.dtype(np.float32): 0, np.dtype(np.float64): 1, np.dtype(np.int32): 2,
               np.dtype(np.int64): 3}

    def upload_csr(self, X):
        """X: scipy CSR matrix.  Values go up in their own dtype when it is float32/64 or
        int32/64 and are cast to float32 on the device (plsa.py:714); other dtypes are cast
        on the host first."""
        indptr = _as(X.indptr, np.int32)
        indices = _as(X.indices, np.int32)
        data = np.ascontiguousarray(X.data)
        if data.dtype not in self._DTYPES:
            data = data.astype(np.float32)
        n, m = X.shape
        check(self._L.plsa_upload_csr_typed(self._h, _ptr(indptr, _i32p), _ptr(indices, _i32p),
                                            data.ctypes.data_as(ctypes.c_void_p),
                                            self._DTYPES[data.dtype], n, m, data.shape[0]),
              self._h)

    def prepare(self, k, refit=False):
        """Build work items sized for k topics (and the term-major copy for a full fit) now
        rather than inside the first em() — the host overlaps its seeded initialisation
        with this."""
        check(self._L.plsa_prepare(self._h, int(bool(refit)), int(k)), self._h)

    def upload_coo(self, rows, cols, vals, n, m):
        rows, cols, vals = _as(rows, np.int32), _as(cols, np.int32), _as(vals, np.float32)
        check(self._L.plsa_upload_coo(self._h, _ptr(rows, _i32p), _ptr(cols, _i32p),
                                      _ptr(vals, _f32p), n, m, vals.shape[0]), self._h)

    def bootstrap(self, row_idx):
        if row_idx is None:
            check(self._L.plsa_bootstrap(self._h, None, 0), self._h)
            return
        idx = _as(row_idx, np.int32)
        check(self._L.plsa_bootstrap(self._h, _ptr(idx, _i32p), idx.shape[0]), self._h)

    @property
    def shape(self):
        n, m, z = _i64(0), _i64(0), _i64(0)
        check(self._L.plsa_corpus_shape(self._h, ctypes.byref(n), ctypes.byref(m),
                                        ctypes.byref(z)), self._h)
        return n.value, m.value, z.value

    # -- model ------------------------------------------------------------------------
    def set_factors(self, p_z_given_d, p_w_given_z):
        pzd, pwz = _as(p_z_given_d, np.float32), _as(p_w_given_z, np.float32)
        n, m, _ = self.shape
        k = pwz.shape[0]
        if pzd.shape != (n, k) or pwz.shape != (k, m):
            raise ValueError("factor shapes %s, %s do not match corpus %s with k=%d"
                             % (pzd.shape, pwz.shape, (n, m), k))
        check(self._L.plsa_set_factors(self._h, _ptr(pzd, _f32p), _ptr(pwz, _f32p), k), self._h)
        self.k = k

    def pinned_factors(self, n_docs, n_terms, k):
        """Two float32 numpy views ([n_docs, k], [k, n_terms]) of page-locked memory owned by
        this context, valid until the next call or the context's end: draw the initial
        factors into them, then set_factors.  May be called while another thread uploads
        the corpus through this context."""
        a, b = _f32p(), _f32p()
        check(self._L.plsa_pinned_factors(self._h, int(n_docs), int(n_terms), int(k),
                                          ctypes.byref(a), ctypes.byref(b)), self._h)
        pzd = np.ctypeslib.as_array(a, shape=(int(n_docs), int(k))) if n_docs else \
            np.empty((0, int(k)), dtype=np.float32)
        pwz = np.ctypeslib.as_array(b, shape=(int(k), int(n_terms))) if n_terms else \
            np.empty((int(k), 0), dtype=np.float32)
        return pzd, pwz

    def set_sample_weight(self, sample_weight):
        if sample_weight is None:
            check(self._L.plsa_set_sample_weight(self._h, None), self._h)
            return
        sw = _as(sample_weight, np.float32)
        if sw.shape != (self.shape[0],):
            raise ValueError("sample_weight.shape == {}, expected {}!".format(
                sw.shape, (self.shape[0],)))
        check(self._L.plsa_set_sample_weight(self._h, _ptr(sw, _f32p)), self._h)

    def get_factors(self, want_pzd=True, want_pwz=True):
        n, m, _ = self.shape
        pzd = np.empty((n, self.k), dtype=np.float32) if want_pzd else None
        pwz = np.empty((self.k, m), dtype=np.float32) if want_pwz else None
        check(self._L.plsa_get_factors(self._h, _ptr(pzd, _f32p), _ptr(pwz, _f32p)), self._h)
        return pzd, pwz

    # -- EM ---------------------------------------------------------------------------
    def em(self, n_iter, n_iter_per_test=10, tolerance=0.001, e_step_thresh=1e-32, refit=False,
           use_sample_weights=False, trace=False):
        """Run the loop of plsa_fit_inner / plsa_refit_inner.  Returns (iters_run, ll_trace)."""
        iters, n_ll = _i32(0), _i32(0)
        cap = int(n_iter) // max(1, int(n_iter_per_test)) + 3
        buf = np.zeros(cap, dtype=np.float64) if (trace or not refit) else None
        check(self._L.plsa_em(self._h, int(n_iter), int(n_iter_per_test), float(tolerance),
                              float(e_step_thresh), int(bool(refit)),
                              int(bool(use_sample_weights)), ctypes.byref(iters),
                              _ptr(buf, _f64p), cap if buf is not None else 0,
                              ctypes.byref(n_ll)), self._h)
        ll = buf[: min(cap, n_ll.value)].copy() if buf is not None else np.zeros(0)
        return iters.value, ll

    def log_likelihood(self):
        out = _f64(0.0)
        check(self._L.plsa_log_likelihood(self._h, ctypes.byref(out)), self._h)
        return out.value

    # -- measurement --------------------------------------------------------------------
    @property
    def last_em_ms(self):
        ms = _f32(0.0)
        check(self._L.plsa_last_em_ms(self._h, ctypes.byref(ms)), self._h)
        return ms.value

    def set_profiling(self, on=True):
        check(self._L.plsa_set_profiling(self._h, int(bool(on))), self._h)

    def profile(self):
        ms = np.zeros(len(PROF_SLOTS), dtype=np.float64)
        cnt = np.zeros(len(PROF_SLOTS), dtype=np.int64)
        check(self._L.plsa_get_profile(self._h, _ptr(ms, _f64p), _ptr(cnt, _i64p)), self._h)
        return {name: {"ms": float(ms[i]), "launches": int(cnt[i])}
                for i, name in enumerate(PROF_SLOTS)}

    @property
    def launches(self):
        out = _i64(0)
        check(self._L.plsa_launch_count(self._h, ctypes.byref(out)), self._h)
        return out.value

    def set_option(self, name, value):
        check(self._L.plsa_set_option(self._h, name.encode(), int(value)), self._h)

    def set_shard(self, comm):
        """Attach (or, with None, detach) the communicator of a document-sharded fit: this
        context holds one shard of the rows; plsa_em adds the shards' P(w|z) sums and
        log-likelihoods over it."""
        check(self._L.plsa_set_shard(self._h, comm._h if comm is not None else None), self._h)
        self._shard = comm   # keep the communicator alive while attached

    def shard_p2p_prepare(self):
        """Allocate this rank's exchange block of the peer-memory all-reduce (after set_shard
        and set_factors).  Returns (device address, bytes)."""
        base, nbytes = ctypes.c_uint64(), _i64()
        check(self._L.plsa_shard_p2p_prepare(self._h, ctypes.byref(base), ctypes.byref(nbytes)),
              self._h)
        return int(base.value), int(nbytes.value)

    def shard_p2p_export(self):
        """CUDA IPC handle (64 bytes) of the exchange block, for peers in other processes."""
        buf = ctypes.create_string_buffer(64)
        check(self._L.plsa_shard_p2p_export(self._h, buf), self._h)
        return buf.raw

    def shard_p2p_attach(self, peer_rank, peer_device, base=0, handle=None):
        """Map a peer's exchange block: ``base`` (its device address) when the peer lives in
        this process, ``handle`` (its IPC handle) otherwise."""
        check(self._L.plsa_shard_p2p_attach(self._h, int(peer_rank), int(peer_device), int(base),
                                            handle), self._h)

    def shard_p2p_detach(self):
        """Unmap the peers' exchange blocks (before any rank frees its own)."""
        check(self._L.plsa_shard_p2p_detach(self._h), self._h)

    def debug_items(self, which):
        """Test hook: the device-resident work items of a pass (0 doc, 1 term, 2 tiled tail) as a
        dict like plan_items', plus chunk, align and the row pointers they were planned from."""
        n, ns, nl, ch, al = _i64(0), _i32(0), _i32(0), _i64(0), _i32(0)
        check(self._L.plsa_debug_items(self._h, int(which), 0, None, None, None, None, None,
                                       ctypes.byref(n), ctypes.byref(ns), ctypes.byref(nl),
                                       ctypes.byref(ch), ctypes.byref(al), None, 0), self._h)
        rows = self.shape[1] if which == 1 else self.shape[0]
        out = dict(start=np.empty(n.value, np.int64), row=np.empty(n.value, np.int32),
                   len=np.empty(n.value, np.int32), slot=np.empty(n.value, np.int32),
                   skip=np.empty(n.value, np.int32), indptr=np.empty(rows + 1, np.int32))
        check(self._L.plsa_debug_items(self._h, int(which), n.value, _ptr(out["start"], _i64p),
                                       _ptr(out["row"], _i32p), _ptr(out["len"], _i32p),
                                       _ptr(out["slot"], _i32p), _ptr(out["skip"], _i32p),
                                       ctypes.byref(n), ctypes.byref(ns), ctypes.byref(nl),
                                       ctypes.byref(ch), ctypes.byref(al), _ptr(out["indptr"], _i32p),
                                       rows + 1), self._h)
        out.update(n_split=ns.value, n_slots=nl.value, chunk=ch.value, align=al.value)
        return out

    def stash_topics(self, slot, n_slots):
        check(self._L.plsa_stash_topics(self._h, int(slot), int(n_slots)), self._h)


# ---- context pool ---------------------------------------------------------------------------
# A fit that is not handed a context borrows one from here and gives it back afterwards, so
# repeated fits (PLSA.fit then transform, cross-validation loops, ...) reuse the device
# buffers instead of paying cudaMalloc/cudaFree each time.  release_device_memory() frees them.
_pool = {}
_pool_lock = threading.Lock()
_POOL_MAX_PER_DEVICE = 2
# A pooled context keeps its device buffers (two copies of the corpus, factors, sort scratch:
# roughly 40 bytes per stored entry).  Contexts of corpora above this many stored entries are
# destroyed instead of pooled, so that a long-lived process does not sit on gigabytes of HBM;
# ENSTOP_B200_POOL_MAX_NNZ overrides (0 disables the pool).
_POOL_MAX_NNZ = int(os.environ.get("ENSTOP_B200_POOL_MAX_NNZ", 64_000_000))


def acquire_context(device=0):
    device = int(device)
    with _pool_lock:
        free = _pool.get(device)
        if free:
            return free.pop()
    return Context(device)


def release_context(ctx):
    try:
        small = ctx._h and ctx.shape[2] <= _POOL_MAX_NNZ
    except PlsaError:
        small = False
    if small:
        with _pool_lock:
            free = _pool.setdefault(ctx.device, [])
            if len(free) < _POOL_MAX_PER_DEVICE:
                free.append(ctx)
                return
    ctx.close()


def release_device_memory():
    """Destroy the pooled contexts (and with them every cached device buffer)."""
    with _pool_lock:
        ctxs = [c for free in _pool.values() for c in free]
        _pool.clear()
    for c in ctxs:
        c.close()


def gather_warmup(devices):
    """Create the NCCL communicators gather_topics will need for this device list (kept by
    the library); call it from a helper thread while the members are being fitted."""
    devs = np.ascontiguousarray(devices, dtype=np.int32)
    check(lib().plsa_gather_warmup(_ptr(devs, _i32p), len(devs)))


def stash_append(dst, src, n_dst, n_src):
    """Append src's first n_src stashed topic matrices to dst's first n_dst (same device)."""
    check(lib().plsa_stash_append(dst._h, src._h, int(n_dst), int(n_src)), dst._h)


def gather_topics(contexts, n_slots):
    """np.vstack of the stashed P(w|z) of every context (enstop_.py:231): slots
    [0, n_slots[i]) of contexts[i], context order then slot order, moved to contexts[0]'s
    device over NCCL send/recv and copied to the host once."""
    L = lib()
    k = contexts[0].k
    m = contexts[0].shape[1]
    counts = np.ascontiguousarray(n_slots, dtype=np.int32)
    out = np.empty((int(counts.sum()) * k, m), dtype=np.float32)
    arr = (_ctx * len(contexts))(*[c._h for c in contexts])
    check(L.plsa_gather_topics(arr, len(contexts), _ptr(counts, _i32p), _ptr(out, _f32p)))
    return out


def topic_distances(topics, kind, device=0):
    """All-pairs distances between topic vectors on the GPU: kind "hellinger" or "kl"
    (enstop_.py:234-263).  topics [N, m] -> float64 [N, N]."""
    topics = np.ascontiguousarray(topics, dtype=np.float32)
    if topics.ndim != 2:
        raise ValueError("topics must be a 2-d array")
    n, m = topics.shape
    out = np.zeros((n, n), dtype=np.float64)
    code = {"hellinger": 0, "kl": 1}[kind]
    check(lib().plsa_topic_distances(int(device), _ptr(topics, _f32p), n, m, code,
                                     _ptr(out, _f64p)))
    return out


def last_distances_ms():
    """Device time of the kernels of this thread's last topic_distances / gathered_distances."""
    ms = _f32(0.0)
    check(lib().plsa_last_distances_ms(ctypes.byref(ms)))
    return ms.value


def gathered_distances(ctx, kind):
    """All-pairs distances ("hellinger" / "kl") of the topic stack the last gather left on
    ``ctx``'s device (rows in the gather's own order: context / rank major)."""
    n = _i64(0)
    code = {"hellinger": 0, "kl": 1}[kind]
    check(lib().plsa_gathered_distances(ctx._h, code, None, ctypes.byref(n)), ctx._h)
    out = np.zeros((n.value, n.value), dtype=np.float64)
    check(lib().plsa_gathered_distances(ctx._h, code, _ptr(out, _f64p), ctypes.byref(n)), ctx._h)
    return out


class Comm:
    """One rank of a one-process-per-GPU job (NCCL communicator behind ``plsa_comm``)."""

    def __init__(self, device, n_ranks, rank, unique_id):
        self._L = lib()
        self._h = ctypes.c_void_p()
        self.n_ranks, self.rank = int(n_ranks), int(rank)
        check(self._L.plsa_comm_create(int(device), self.n_ranks, self.rank, unique_id,
                                       ctypes.byref(self._h)))

    @staticmethod
    def unique_id():
        buf = ctypes.create_string_buffer(128)
        check(lib().plsa_nccl_unique_id(buf))
        return buf.raw

    def gather_topics(self, ctx, n_per_rank, root=0):
        counts = np.ascontiguousarray(n_per_rank, dtype=np.int32)
        out = None
        if self.rank == root:
            out = np.empty((int(counts.sum()) * ctx.k, ctx.shape[1]), dtype=np.float32)
        check(self._L.plsa_comm_gather_topics(self._h, ctx._h, _ptr(counts, _i32p), int(root),
                                              _ptr(out, _f32p)), ctx._h)
        return out

    def abort(self):
        """Cancel outstanding collectives so that peers blocked on this communicator return."""
        if self._h:
            self._L.plsa_comm_abort(self._h)

    def close(self):
        if self._h:
            self._L.plsa_comm_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
