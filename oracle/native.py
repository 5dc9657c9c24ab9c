"""TEST INFRASTRUCTURE ONLY -- ctypes bindings for the two CPU checkers.

``OracleSampler``   -> oracle/_build/libsgcn_oracle.so  (plain-C restatement, sgcn_oracle.c)
``RefSampler``      -> oracle/_ref/libsgcn_ref.so       (unmodified reference C++ behind ref_shim.cpp)

Both expose the same methods so a test can be parametrised over them.  ``build()`` runs
``make -C oracle`` (compiling the checker is not using it).
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "_build", "libsgcn_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libsgcn_ref.so")
REF_TEST_MULT = os.path.join(HERE, "_ref", "test_mult")

_f32p = C.POINTER(C.c_float)
_i32p = C.POINTER(C.c_int)


def build(verbose=False):
    """Compile the C restatement and, when /root/reference is present, oracle/_ref."""
    out = subprocess.run(["make", "-C", HERE, "all"], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + out.stdout + out.stderr)
    if verbose:
        print(out.stdout)


def have_ref():
    return os.path.exists(REF_SO)


def _fp(a):
    return a.ctypes.data_as(_f32p)


def _ip(a):
    return a.ctypes.data_as(_i32p)


_oracle = None
_ref = None


def oracle_lib():
    global _oracle
    if _oracle is None:
        if not os.path.exists(ORACLE_SO):
            build()
        lib = C.CDLL(ORACLE_SO)
        lib.orc_sampler_create.restype = C.c_void_p
        lib.orc_sampler_create.argtypes = [_f32p, _i32p, _i32p, C.c_int, C.c_int, C.c_int, C.c_int]
        lib.orc_sampler_destroy.argtypes = [C.c_void_p]
        lib.orc_sampler_seed.argtypes = [C.c_void_p, C.c_int]
        lib.orc_sampler_start_batch.argtypes = [C.c_void_p, C.c_int, _i32p]
        lib.orc_sampler_expand.argtypes = [C.c_void_p, C.c_int]
        lib.orc_sampler_expand.restype = C.c_int
        lib.orc_sampler_int_vec.argtypes = [C.c_void_p, C.c_int, C.POINTER(_i32p)]
        lib.orc_sampler_float_vec.argtypes = [C.c_void_p, C.c_int, C.POINTER(_f32p)]
        lib.orc_mult_create.restype = C.c_void_p
        lib.orc_mult_create.argtypes = [_f32p, C.c_int]
        lib.orc_mult_destroy.argtypes = [C.c_void_p]
        lib.orc_mult_draw.argtypes = [C.c_void_p]
        lib.orc_mult_descend.argtypes = [C.c_void_p, C.c_float]
        lib.orc_mult_tree.argtypes = [C.c_void_p, C.POINTER(_f32p)]
        lib.orc_mt_seed.argtypes = [C.c_void_p, C.c_uint32]
        lib.orc_mt_next.argtypes = [C.c_void_p]
        lib.orc_mt_next.restype = C.c_uint32
        lib.orc_u32_to_canonical.argtypes = [C.c_uint32]
        lib.orc_u32_to_canonical.restype = C.c_float
        lib.orc_slice_indptr.argtypes = [C.c_int, _i32p, _i32p, _i32p]
        lib.orc_slice_rows.argtypes = [C.c_int, _i32p, _f32p, _i32p, _i32p, _f32p, _i32p, _i32p]
        lib.orc_dense_slice.argtypes = [C.c_int, C.c_int, _i32p, _f32p, _f32p]
        lib.orc_spmm_coo.argtypes = [C.c_int, _i32p, _i32p, _f32p, _f32p, C.c_int, _f32p, C.c_int]
        lib.orc_spmm_coo_t.argtypes = [C.c_int, _i32p, _i32p, _f32p, _f32p, C.c_int, _f32p, C.c_int]
        lib.orc_gather_rows.argtypes = [C.c_int, C.c_int, _i32p, _f32p, _f32p]
        lib.orc_scatter_rows.argtypes = [C.c_int, C.c_int, _i32p, _f32p, _f32p]
        lib.orc_spmm_csr_omp.argtypes = [C.c_int, _i32p, _i32p, _f32p, _f32p, C.c_int, _f32p, C.c_int]
        lib.orc_cv_forward_omp.argtypes = [C.c_int, _i32p, _i32p, _f32p, _f32p, _i32p, _f32p, C.c_int,
                                           _i32p, _i32p, _i32p, _f32p, _f32p, C.c_int]
        _oracle = lib
    return _oracle


def ref_lib():
    global _ref
    if _ref is None:
        if not os.path.exists(REF_SO):
            build()
        if not os.path.exists(REF_SO):
            raise FileNotFoundError("oracle/_ref/libsgcn_ref.so is absent and /root/reference is not mounted")
        lib = C.CDLL(REF_SO)
        lib.ref_sched_create.restype = C.c_void_p
        lib.ref_sched_create.argtypes = [_f32p, _i32p, _i32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        lib.ref_sched_destroy.argtypes = [C.c_void_p]
        lib.ref_sched_seed.argtypes = [C.c_void_p, C.c_int]
        lib.ref_sched_start_batch.argtypes = [C.c_void_p, C.c_int, _i32p]
        lib.ref_sched_expand.argtypes = [C.c_void_p, C.c_int]
        lib.ref_sched_expand.restype = C.c_int
        lib.ref_sched_int_vec.argtypes = [C.c_void_p, C.c_int, C.POINTER(_i32p)]
        lib.ref_sched_float_vec.argtypes = [C.c_void_p, C.c_int, C.POINTER(_f32p)]
        lib.ref_mult_create.restype = C.c_void_p
        lib.ref_mult_create.argtypes = [_f32p, C.c_int]
        lib.ref_mult_destroy.argtypes = [C.c_void_p]
        lib.ref_mult_query.argtypes = [C.c_void_p]
        lib.ref_mult_query_u.argtypes = [C.c_void_p, C.c_float]
        lib.ref_mult_bit.argtypes = [C.c_void_p, C.POINTER(_f32p)]
        lib.ref_c_indptr.argtypes = [C.c_int, _i32p, _i32p, _i32p]
        lib.ref_c_slice.argtypes = [C.c_int, _i32p, _f32p, _i32p, _i32p, _f32p, _i32p, _i32p]
        lib.ref_c_dense_slice.argtypes = [C.c_int, C.c_int, _i32p, _f32p, _f32p]
        if hasattr(lib, "ref_compute_history"):
            lib.ref_compute_history.argtypes = [_f32p, _i32p, _i32p, C.c_int, C.c_int, _i32p, C.c_int, _f32p,
                                                C.c_int, _f32p]
        _ref = lib
    return _ref


INT_VECS = {"field": 0, "ffield": 1, "edg_s": 2, "edg_t": 3, "fedg_s": 4, "fedg_t": 5,
            "adj_i": 6, "adj_p": 7, "visited": 8, "fvisited": 9}
FLOAT_VECS = {"scales": 0, "edg_w": 1, "medg_w": 2, "fedg_w": 3, "adj_w": 4, "importance": 5}


class _SamplerBase:
    """Common surface of Scheduler (gcn/scheduler.h:6-28)."""

    _prefix = None

    def __init__(self, lib, adj_w, adj_i, adj_p, cv=False, importance=False):
        self._lib = lib
        w = np.ascontiguousarray(adj_w, dtype=np.float32)
        i = np.ascontiguousarray(adj_i, dtype=np.int32)
        p = np.ascontiguousarray(adj_p, dtype=np.int32)
        self.num_data = len(p) - 1
        self.num_edges = len(i)
        self.cv, self.importance = bool(cv), bool(importance)
        self._h = self._create(w, i, p)

    def vec(self, name):
        if name in INT_VECS:
            ptr = _i32p()
            n = self._get_i(self._h, INT_VECS[name], C.byref(ptr))
            return np.ctypeslib.as_array(ptr, shape=(n,)).copy() if n > 0 else np.zeros(0, np.int32)
        ptr = _f32p()
        n = self._get_f(self._h, FLOAT_VECS[name], C.byref(ptr))
        return np.ctypeslib.as_array(ptr, shape=(n,)).copy() if n > 0 else np.zeros(0, np.float32)

    def snapshot(self):
        names = ["field", "ffield", "edg_s", "edg_t", "fedg_s", "fedg_t", "scales", "edg_w", "medg_w", "fedg_w"]
        return {k: self.vec(k) for k in names}


class OracleSampler(_SamplerBase):
    def __init__(self, adj_w, adj_i, adj_p, cv=False, importance=False):
        super().__init__(oracle_lib(), adj_w, adj_i, adj_p, cv, importance)

    def _create(self, w, i, p):
        self._get_i = self._lib.orc_sampler_int_vec
        self._get_f = self._lib.orc_sampler_float_vec
        return self._lib.orc_sampler_create(_fp(w), _ip(i), _ip(p), self.num_data, self.num_edges,
                                            int(self.cv), int(self.importance))

    def seed(self, s):
        self._lib.orc_sampler_seed(self._h, int(s))

    def start_batch(self, ids):
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        self._lib.orc_sampler_start_batch(self._h, len(ids), _ip(ids))

    def expand(self, degree):
        return self._lib.orc_sampler_expand(self._h, int(degree))

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.orc_sampler_destroy(self._h)
            self._h = None


class RefSampler(_SamplerBase):
    """The unmodified reference Scheduler (gcn/scheduler.cpp) behind oracle/ref_shim.cpp."""

    def __init__(self, adj_w, adj_i, adj_p, cv=False, importance=False, L=1):
        self._L = L
        super().__init__(ref_lib(), adj_w, adj_i, adj_p, cv, importance)

    def _create(self, w, i, p):
        self._get_i = self._lib.ref_sched_int_vec
        self._get_f = self._lib.ref_sched_float_vec
        # the reference ctor takes indptr WITHOUT the trailing nnz (it pushes it itself,
        # gcn/scheduler.cpp:16,20) and num_data = labels.shape[0]
        return self._lib.ref_sched_create(_fp(w), _ip(i), _ip(p), self.num_data, self.num_edges,
                                          self._L, int(self.cv), int(self.importance))

    def seed(self, s):
        self._lib.ref_sched_seed(self._h, int(s))

    def start_batch(self, ids):
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        self._lib.ref_sched_start_batch(self._h, len(ids), _ip(ids))

    def expand(self, degree):
        return int(self._lib.ref_sched_expand(self._h, int(degree)))

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.ref_sched_destroy(self._h)
            self._h = None


class OracleMult:
    def __init__(self, prob):
        self._lib = oracle_lib()
        p = np.ascontiguousarray(prob, dtype=np.float32)
        self._h = self._lib.orc_mult_create(_fp(p), len(p))
        if not self._h:
            raise RuntimeError("Prob is empty")

    def query(self, u=None):
        if u is None:
            return self._lib.orc_mult_draw(self._h)
        return self._lib.orc_mult_descend(self._h, float(u))

    def bit(self):
        ptr = _f32p()
        n = self._lib.orc_mult_tree(self._h, C.byref(ptr))
        return np.ctypeslib.as_array(ptr, shape=(n,)).copy()

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.orc_mult_destroy(self._h)
            self._h = None


class RefMult:
    def __init__(self, prob):
        self._lib = ref_lib()
        p = np.ascontiguousarray(prob, dtype=np.float32)
        self._h = self._lib.ref_mult_create(_fp(p), len(p))
        if not self._h:
            raise RuntimeError("Prob is empty")

    def query(self, u=None):
        if u is None:
            return self._lib.ref_mult_query(self._h)
        return self._lib.ref_mult_query_u(self._h, float(u))

    def bit(self):
        ptr = _f32p()
        n = self._lib.ref_mult_bit(self._h, C.byref(ptr))
        return np.ctypeslib.as_array(ptr, shape=(n,)).copy()

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.ref_mult_destroy(self._h)
            self._h = None


class MT19937:
    """std::mt19937 restated (oracle/sgcn_oracle.c); used to pin the device generator."""

    class _State(C.Structure):
        _fields_ = [("x", C.c_uint32 * 624), ("pos", C.c_int)]

    def __init__(self, seed):
        self._lib = oracle_lib()
        self._s = MT19937._State()
        self._lib.orc_mt_seed(C.byref(self._s), C.c_uint32(seed & 0xFFFFFFFF))

    def next(self):
        return self._lib.orc_mt_next(C.byref(self._s))

    def draws(self, n):
        return np.array([self.next() for _ in range(n)], dtype=np.uint32)


def u32_to_canonical(r):
    return oracle_lib().orc_u32_to_canonical(C.c_uint32(int(r)))


# ---- slicers ------------------------------------------------------------------------------

def _slice_with(fn_indptr, fn_slice, a, r):
    """history.slice semantics (gcn/_history.pyx:25-51) on top of the given C entry points."""
    r = np.ascontiguousarray(r, dtype=np.int32)
    n = len(r)
    indptr = np.zeros(n + 1, dtype=np.int32)
    a_p = np.ascontiguousarray(a.indptr, dtype=np.int32)
    a_i = np.ascontiguousarray(a.indices, dtype=np.int32)
    a_d = np.ascontiguousarray(a.data, dtype=np.float32)
    fn_indptr(n, _ip(r), _ip(a_p), _ip(indptr))
    nnz = int(indptr[n])
    if nnz == 0:
        return None
    data = np.zeros(nnz, dtype=np.float32)
    indices = np.zeros((nnz, 2), dtype=np.int32)
    fn_slice(n, _ip(r), _fp(a_d), _ip(a_i), _ip(a_p), _fp(data), _ip(indices), _ip(indptr))
    return indices, data, np.array([n, a.shape[1]], dtype=np.int32)


def oracle_slice(a, r):
    lib = oracle_lib()
    return _slice_with(lib.orc_slice_indptr, lib.orc_slice_rows, a, r)


def ref_slice(a, r):
    lib = ref_lib()
    return _slice_with(lib.ref_c_indptr, lib.ref_c_slice, a, r)


def oracle_dense_slice(a, r):
    a = np.ascontiguousarray(a, dtype=np.float32)
    r = np.ascontiguousarray(r, dtype=np.int32)
    out = np.zeros((len(r), a.shape[1]), dtype=np.float32)
    oracle_lib().orc_dense_slice(len(r), a.shape[1], _ip(r), _fp(a), _fp(out))
    return out


def ref_dense_slice(a, r):
    a = np.ascontiguousarray(a, dtype=np.float32)
    r = np.ascontiguousarray(r, dtype=np.int32)
    out = np.zeros((len(r), a.shape[1]), dtype=np.float32)
    ref_lib().ref_c_dense_slice(len(r), a.shape[1], _ip(r), _fp(a), _fp(out))
    return out


def ref_compute_history(adj_w, adj_i, adj_p, rows, history):
    """The reference's own (commented-out) CSR aggregation loop, gcn/history.cpp:10-37, compiled by
    oracle/Makefile: output[i, :] = sum over the stored row rows[i] of adj_w * history[adj_i, :], fp32,
    in storage order.  adj_p has N + 1 entries."""
    lib = ref_lib()
    if not hasattr(lib, "ref_compute_history"):
        raise FileNotFoundError("oracle/_ref/libsgcn_ref.so was built without compute_history; run make -C oracle")
    w = np.ascontiguousarray(adj_w, dtype=np.float32)
    i = np.ascontiguousarray(adj_i, dtype=np.int32)
    p = np.ascontiguousarray(adj_p, dtype=np.int32)
    r = np.ascontiguousarray(rows, dtype=np.int32)
    h = np.ascontiguousarray(history, dtype=np.float32)
    out = np.zeros((len(r), h.shape[1]), dtype=np.float32)
    lib.ref_compute_history(_fp(w), _ip(i), _ip(p), len(p) - 1, len(i), _ip(r), len(r), _fp(h), h.shape[1], _fp(out))
    return out


# ---- fp32 numeric kernels ------------------------------------------------------------------

def spmm_coo(idx, val, shape, x):
    """tf.sparse_tensor_dense_matmul in storage order, fp32 (gcn/layers.py:31-37)."""
    idx = np.asarray(idx, dtype=np.int32).reshape(-1, 2)
    rows = np.ascontiguousarray(idx[:, 0])
    cols = np.ascontiguousarray(idx[:, 1])
    val = np.ascontiguousarray(val, dtype=np.float32)
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.zeros((int(shape[0]), x.shape[1]), dtype=np.float32)
    oracle_lib().orc_spmm_coo(len(val), _ip(rows), _ip(cols), _fp(val), _fp(x), x.shape[1], _fp(y), int(shape[0]))
    return y


def spmm_coo_t(idx, val, shape, dy):
    idx = np.asarray(idx, dtype=np.int32).reshape(-1, 2)
    rows = np.ascontiguousarray(idx[:, 0])
    cols = np.ascontiguousarray(idx[:, 1])
    val = np.ascontiguousarray(val, dtype=np.float32)
    dy = np.ascontiguousarray(dy, dtype=np.float32)
    dx = np.zeros((int(shape[1]), dy.shape[1]), dtype=np.float32)
    oracle_lib().orc_spmm_coo_t(len(val), _ip(rows), _ip(cols), _fp(val), _fp(dy), dy.shape[1], _fp(dx), int(shape[1]))
    return dx
