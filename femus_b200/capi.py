"""ctypes binding of include/femus_b200.h -- the harness side of the C ABI (tests, bench, smoke).

The product is the shared library; this module only loads it, declares the prototypes and wraps
handles in small Python classes.  It fails loudly when the library is missing: there is no CPU
fallback anywhere in femus_b200."""
import ctypes
import os
import re
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libfemus_b200.so")
HEADER = os.path.join(os.path.dirname(HERE), "include", "femus_b200.h")

_lib = None

vp, ci, cd, i64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_int64

_PROTOS = {
    # name: (restype, argtypes)
    "b2_last_error": (ctypes.c_char_p, []),
    "b2_version": (ci, []),
    "b2_ctx_create": (ci, [ci, vp]),
    "b2_ctx_destroy": (ci, [vp]),
    "b2_ctx_sync": (ci, [vp]),
    "b2_ctx_stream": (vp, [vp]),
    "b2_ctx_device": (ci, [vp]),
    "b2_nccl_unique_id": (ci, [vp]),
    "b2_ctx_comm_init": (ci, [vp, ci, ci, vp]),
    "b2_ctx_nranks": (ci, [vp]),
    "b2_ctx_rank": (ci, [vp]),
    "b2_ctx_bytes_in_use": (i64, [vp]),
    "b2_ctx_launch_count": (i64, [vp, ci]),
    "b2_timer_start": (ci, [vp]),
    "b2_timer_stop_ms": (ci, [vp, vp]),
    "b2_ctx_flush_l2": (ci, [vp]),
    "b2_ctx_measure_fp64_tensor": (ci, [vp, vp]),
    "b2_ctx_measure_fp64_fma": (ci, [vp, vp]),
    "b2_asm_kernel_name": (ctypes.c_char_p, [vp]),
    "b2_ctx_set_option": (ci, [vp, ctypes.c_char_p, ci]),
    "b2_ctx_profile": (ci, [vp, ci]),
    "b2_ctx_profile_only": (ci, [vp, vp]),
    "b2_ctx_profile_read": (ci, [vp, vp, vp, vp]),
    "b2_ctx_profile_clear": (ci, [vp]),
    "b2_vec_put_async": (ci, [vp, vp, i64]),
    "b2_vec_prefetch": (ci, [vp, vp, i64]),
    "b2_vec_fetch": (ci, [vp, vp, i64]),
    "b2_ctx_open_copies": (ci, [vp]),
    "b2_ctx_join_copies": (ci, [vp]),
    "b2_ctx_mark_copies": (ci, [vp]),
    "b2_ctx_wait_marked": (ci, [vp]),
    "b2_mesh_prefetch": (ci, [vp, vp, vp]),
    "b2_mesh_swap": (ci, [vp]),
    "b2_vec_get_async": (ci, [vp, vp, i64]),
    "b2_mesh_update": (ci, [vp, vp, vp]),
    "b2_vec_create": (ci, [vp, i64, vp]),
    "b2_vec_destroy": (ci, [vp]),
    "b2_vec_size": (i64, [vp]),
    "b2_vec_device_ptr": (vp, [vp]),
    "b2_vec_zero": (ci, [vp]),
    "b2_vec_fill": (ci, [vp, cd]),
    "b2_vec_put": (ci, [vp, vp, i64]),
    "b2_vec_get": (ci, [vp, vp, i64]),
    "b2_vec_copy": (ci, [vp, vp]),
    "b2_vec_axpy": (ci, [vp, cd, vp]),
    "b2_vec_aypx": (ci, [vp, cd, vp]),
    "b2_vec_scale": (ci, [vp, cd]),
    "b2_vec_add_scalar": (ci, [vp, cd]),
    "b2_vec_pointwise_mult": (ci, [vp, vp, vp]),
    "b2_vec_dot": (ci, [vp, vp, vp]),
    "b2_vec_norm": (ci, [vp, ci, vp]),
    "b2_vec_sum": (ci, [vp, vp]),
    "b2_vec_minmax": (ci, [vp, vp, vp]),
    "b2_vec_abs": (ci, [vp]),
    "b2_mg_set_level_nullspace": (ci, [vp, ci, vp]),
    "b2_mg_set_timing": (ci, [vp, ci]),
    "b2_mg_get_timing": (ci, [vp, vp]),
    "b2_ctx_peer_export": (ci, [vp, i64, vp]),
    "b2_ctx_peer_open": (ci, [vp, vp]),
    "b2_ctx_peer_error": (ci, [vp, vp]),
    "b2_halo_set_exchange": (ci, [vp, vp, vp, vp, vp]),
    "b2_csr_matmat": (ci, [vp, vp, vp]),
    "b2_csr_axpy": (ci, [vp, cd, vp]),
    "b2_csr_pattern_contains": (ci, [vp, vp, vp]),
    "b2_vec_set_indexed": (ci, [vp, vp, vp, i64]),
    "b2_vec_add_indexed": (ci, [vp, vp, vp, i64]),
    "b2_vec_fill_indexed": (ci, [vp, vp, i64, cd]),
    "b2_vec_get_indexed": (ci, [vp, vp, vp, i64]),
    "b2_vec_copy_masked": (ci, [vp, vp, vp, cd]),
    "b2_csr_create": (ci, [vp, i64, i64, vp, vp, vp, vp]),
    "b2_csr_create_from_elements": (ci, [vp, i64, i64, ci, vp, vp]),
    "b2_csr_destroy": (ci, [vp]),
    "b2_csr_nrows": (i64, [vp]),
    "b2_csr_ncols": (i64, [vp]),
    "b2_csr_nnz": (i64, [vp]),
    "b2_csr_get": (ci, [vp, vp, vp, vp]),
    "b2_csr_put_vals": (ci, [vp, vp]),
    "b2_csr_zero": (ci, [vp]),
    "b2_csr_copy_vals": (ci, [vp, vp]),
    "b2_csr_add_blocks": (ci, [vp, i64, ci, ci, vp, vp, vp]),
    "b2_csr_set_rows": (ci, [vp, i64, vp, vp, vp, vp]),
    "b2_csr_zero_rows": (ci, [vp, vp, i64, cd]),
    "b2_csr_zero_cols": (ci, [vp, vp, i64]),
    "b2_csr_diag": (ci, [vp, vp]),
    "b2_csr_transpose": (ci, [vp, vp]),
    "b2_csr_spmv": (ci, [vp, vp, vp]),
    "b2_csr_spmv_add": (ci, [vp, vp, vp]),
    "b2_csr_spmv_t": (ci, [vp, vp, vp]),
    "b2_csr_resid": (ci, [vp, vp, vp, vp]),
    "b2_csr_jacobi_sweep": (ci, [vp, vp, vp, vp, vp, cd]),
    "b2_csr_ptap": (ci, [vp, vp, vp]),
    "b2_csr_last_kernel_ms": (cd, [vp]),
    "b2_galerkin_create": (ci, [vp, vp, i64, ci, ci, vp, vp, vp, vp, vp, vp, vp, vp]),
    "b2_galerkin_apply": (ci, [vp]),
    "b2_galerkin_record_elements": (ci, [vp, ci]),
    "b2_galerkin_apply_from_elements": (ci, [vp, vp]),
    "b2_galerkin_destroy": (ci, [vp]),
    "b2_mesh_create": (ci, [vp, i64, i64, vp, vp, vp]),
    "b2_mesh_destroy": (ci, [vp]),
    "b2_asm_create": (ci, [vp, vp, ci, vp, ci, vp, vp, vp, vp, vp, vp]),
    "b2_asm_destroy": (ci, [vp]),
    "b2_asm_poisson": (ci, [vp, vp, vp, cd, cd]),
    "b2_asm_poisson_galerkin": (ci, [vp, vp, vp, vp, cd, cd]),
    "b2_asm_neumann": (ci, [vp, i64, vp, vp, vp, ci, vp, vp, vp, vp, vp, vp]),
    "b2_asm_neumann_faces": (ci, [vp, i64, vp, vp, vp, ci, ci, vp, vp, vp, vp, vp, vp]),
    "b2_asm_last_kernel_ms": (cd, [vp]),
    "b2_mg_create": (ci, [vp, ci, vp]),
    "b2_mg_set_level": (ci, [vp, ci, vp, vp, vp, i64, ci, ci, cd]),
    "b2_mg_set_coarse": (ci, [vp, cd, ci]),
    "b2_mg_set_smoother": (ci, [vp, ci, ci, cd, cd]),
    "b2_schwarz_create": (ci, [vp, vp, i64, vp, vp, i64, vp, vp, vp]),
    "b2_schwarz_set_subsolver": (ci, [vp, ci]),
    "b2_schwarz_set_row_levels": (ci, [vp, ci]),
    "b2_schwarz_row_levels": (i64, [vp]),
    "b2_schwarz_setup": (ci, [vp]),
    "b2_schwarz_apply": (ci, [vp, vp, vp]),
    "b2_schwarz_bytes": (i64, [vp]),
    "b2_schwarz_groups": (i64, [vp]),
    "b2_schwarz_destroy": (ci, [vp]),
    "b2_mg_set_level_schwarz": (ci, [vp, ci, vp]),
    "b2_mg_set_coarse_schwarz": (ci, [vp, vp]),
    "b2_mg_set_level_ksp": (ci, [vp, ci, ci]),
    "b2_stokes_create": (ci, [vp, vp, vp, ci, ci, ci, vp, vp, vp, vp, vp, vp]),
    "b2_stokes_assemble": (ci, [vp, vp, vp, cd]),
    "b2_stokes_destroy": (ci, [vp]),
    "b2_ns_create": (ci, [vp, vp, vp, ci, ci, ci, vp, vp, vp, vp, vp, vp, vp]),
    "b2_ns_assemble": (ci, [vp, vp, vp, cd]),
    "b2_ns_pressure_faces": (ci, [vp, i64, vp, vp, vp, ci, ci, vp, vp, vp, vp, vp, vp]),
    "b2_mg_level_bounds": (ci, [vp, ci, vp, vp]),
    "b2_mg_set_level_halo": (ci, [vp, ci, vp]),
    "b2_halo_create": (ci, [vp, i64, i64, vp, vp, i64, vp, vp, vp]),
    "b2_halo_destroy": (ci, [vp]),
    "b2_halo_owned_count": (i64, [vp]),
    "b2_halo_interface_count": (i64, [vp]),
    "b2_halo_sum": (ci, [vp, vp]),
    "b2_vec_set_halo": (ci, [vp, vp]),
    "b2_mg_vcycle": (ci, [vp, vp, vp]),
    "b2_mg_solve": (ci, [vp, vp, vp]),
    "b2_mg_coarse_iterations": (ci, [vp]),
    "b2_mg_destroy": (ci, [vp]),
}


def header_symbols():
    """Every function name include/femus_b200.h declares."""
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(b2_[a-z0-9_]+)\s*\(", txt)))


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -m femus_b200.build` "
                               "(femus_b200 has no CPU fallback)")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _PROTOS.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


class B2Error(RuntimeError):
    pass


def check(status):
    if status != 0:
        raise B2Error(lib().b2_last_error().decode())


def _ptr(a):
    return None if a is None else a.ctypes.data_as(vp)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


class Context:
    def __init__(self, device=0):
        self.L = lib()
        h = vp()
        check(self.L.b2_ctx_create(device, ctypes.byref(h)))
        self.h = h

    def close(self):
        if self.h:
            self.L.b2_ctx_destroy(self.h)
            self.h = None

    def sync(self):
        check(self.L.b2_ctx_sync(self.h))

    def stream(self):
        return self.L.b2_ctx_stream(self.h)

    def comm_init(self, nranks, rank, uid_bytes):
        buf = (ctypes.c_char * 128).from_buffer_copy(uid_bytes)
        check(self.L.b2_ctx_comm_init(self.h, nranks, rank, buf))

    def peer_init(self, slot_cells, allgather):
        """Peer-memory exchange over NVLink: export this rank's inbox, gather the IPC handles, open every rank's."""
        buf = (ctypes.c_char * 64)()
        check(self.L.b2_ctx_peer_export(self.h, int(slot_cells), buf))
        handles = b"".join(allgather(bytes(buf)))
        check(self.L.b2_ctx_peer_open(self.h, (ctypes.c_char * len(handles)).from_buffer_copy(handles)))

    def peer_error(self):
        e = ci()
        check(self.L.b2_ctx_peer_error(self.h, ctypes.byref(e)))
        return bool(e.value)

    @staticmethod
    def nccl_unique_id():
        buf = (ctypes.c_char * 128)()
        check(lib().b2_nccl_unique_id(buf))
        return bytes(buf)

    def launches(self, reset=False):
        return int(self.L.b2_ctx_launch_count(self.h, 1 if reset else 0))

    def bytes_in_use(self):
        return int(self.L.b2_ctx_bytes_in_use(self.h))

    def timer_start(self):
        check(self.L.b2_timer_start(self.h))

    def timer_stop_ms(self):
        ms = cd()
        check(self.L.b2_timer_stop_ms(self.h, ctypes.byref(ms)))
        return ms.value

    def open_copies(self):
        check(self.L.b2_ctx_open_copies(self.h))

    def join_copies(self):
        check(self.L.b2_ctx_join_copies(self.h))

    def mark_copies(self):
        check(self.L.b2_ctx_mark_copies(self.h))

    def wait_marked(self):
        check(self.L.b2_ctx_wait_marked(self.h))

    def measure_fp64_tensor(self):
        """Measured fp64 tensor-core (DMMA) peak of this device in TFLOP/s."""
        t = cd()
        check(self.L.b2_ctx_measure_fp64_tensor(self.h, ctypes.byref(t)))
        return t.value

    def measure_fp64_fma(self):
        """Measured fp64 CUDA-core (DFMA) issue-rate peak of this device in TFLOP/s."""
        t = cd()
        check(self.L.b2_ctx_measure_fp64_fma(self.h, ctypes.byref(t)))
        return t.value

    def set_option(self, name, value):
        check(self.L.b2_ctx_set_option(self.h, name.encode(), int(value)))

    def flush_l2(self):
        check(self.L.b2_ctx_flush_l2(self.h))

    def profile(self, on):
        check(self.L.b2_ctx_profile(self.h, 1 if on else 0))

    def profile_only(self, obj):
        check(self.L.b2_ctx_profile_only(self.h, obj.h if obj is not None else None))

    def profile_read(self, obj):
        """(launch count, total ms) of the profiled launches of a Csr / Assembler; consumed."""
        n, ms = ci(), cd()
        check(self.L.b2_ctx_profile_read(self.h, obj.h, ctypes.byref(n), ctypes.byref(ms)))
        return n.value, ms.value

    def profile_clear(self):
        check(self.L.b2_ctx_profile_clear(self.h))

    # factories
    def vector(self, n_or_array):
        return Vector(self, n_or_array)

    def csr(self, nrows, ncols, rowptr, col, vals=None):
        return Csr.from_host(self, nrows, ncols, rowptr, col, vals)

    def csr_from_scipy(self, A):
        A = A.tocsr()
        A.sort_indices()
        return Csr.from_host(self, A.shape[0], A.shape[1], A.indptr, A.indices, A.data)


class Vector:
    def __init__(self, ctx, n_or_array):
        self.ctx, self.L = ctx, ctx.L
        h = vp()
        if np.isscalar(n_or_array):
            n = int(n_or_array)
            check(self.L.b2_vec_create(ctx.h, n, ctypes.byref(h)))
            self.h = h
        else:
            a = _f64(n_or_array)
            check(self.L.b2_vec_create(ctx.h, a.shape[0], ctypes.byref(h)))
            self.h = h
            self.put(a)

    def __del__(self):
        try:
            if self.h and self.ctx.h:
                self.L.b2_vec_destroy(self.h)
        except Exception:
            pass

    @property
    def n(self):
        return int(self.L.b2_vec_size(self.h))

    def put(self, a):
        a = _f64(a)
        check(self.L.b2_vec_put(self.h, _ptr(a), a.shape[0]))

    def get(self):
        out = np.empty(self.n)
        check(self.L.b2_vec_get(self.h, _ptr(out), out.shape[0]))
        return out

    def put_async(self, host_ptr, n):
        check(self.L.b2_vec_put_async(self.h, host_ptr, n))

    def prefetch(self, host_ptr, n):
        check(self.L.b2_vec_prefetch(self.h, host_ptr, n))

    def fetch(self, host_ptr, n):
        """D2H on the copy stream, ordered after the compute enqueued so far."""
        check(self.L.b2_vec_fetch(self.h, host_ptr, n))

    def get_async(self, host_ptr, n):
        check(self.L.b2_vec_get_async(self.h, host_ptr, n))

    def zero(self):
        check(self.L.b2_vec_zero(self.h))

    def fill(self, a):
        check(self.L.b2_vec_fill(self.h, float(a)))

    def copy_from(self, x):
        check(self.L.b2_vec_copy(self.h, x.h))

    def axpy(self, a, x):
        check(self.L.b2_vec_axpy(self.h, float(a), x.h))

    def aypx(self, a, x):
        check(self.L.b2_vec_aypx(self.h, float(a), x.h))

    def scale(self, a):
        check(self.L.b2_vec_scale(self.h, float(a)))

    def add_scalar(self, a):
        check(self.L.b2_vec_add_scalar(self.h, float(a)))

    def pointwise_mult(self, x, y):
        check(self.L.b2_vec_pointwise_mult(self.h, x.h, y.h))

    def dot(self, y):
        out = cd()
        check(self.L.b2_vec_dot(self.h, y.h, ctypes.byref(out)))
        return out.value

    def norm(self, kind=2):
        out = cd()
        check(self.L.b2_vec_norm(self.h, kind, ctypes.byref(out)))
        return out.value

    def sum(self):
        out = cd()
        check(self.L.b2_vec_sum(self.h, ctypes.byref(out)))
        return out.value

    def minmax(self):
        a, b = cd(), cd()
        check(self.L.b2_vec_minmax(self.h, ctypes.byref(a), ctypes.byref(b)))
        return a.value, b.value

    def abs(self):
        check(self.L.b2_vec_abs(self.h))

    def set_indexed(self, idx, vals):
        idx, vals = _i32(idx), _f64(vals)
        check(self.L.b2_vec_set_indexed(self.h, _ptr(idx), _ptr(vals), idx.shape[0]))

    def add_indexed(self, idx, vals):
        idx, vals = _i32(idx), _f64(vals)
        check(self.L.b2_vec_add_indexed(self.h, _ptr(idx), _ptr(vals), idx.shape[0]))

    def fill_indexed(self, idx, a):
        idx = _i32(idx)
        check(self.L.b2_vec_fill_indexed(self.h, _ptr(idx), idx.shape[0], float(a)))

    def get_indexed(self, idx):
        idx = _i32(idx)
        out = np.empty(idx.shape[0])
        check(self.L.b2_vec_get_indexed(self.h, _ptr(idx), _ptr(out), idx.shape[0]))
        return out

    def set_halo(self, halo):
        self._halo = halo
        check(self.L.b2_vec_set_halo(self.h, halo.h if halo is not None else None))

    def copy_masked(self, src, mask, thr):
        check(self.L.b2_vec_copy_masked(self.h, src.h, mask.h, float(thr)))


class Csr:
    def __init__(self, ctx, h):
        self.ctx, self.L, self.h = ctx, ctx.L, h

    @classmethod
    def from_host(cls, ctx, nrows, ncols, rowptr, col, vals=None):
        rowptr, col = _i64(rowptr), _i32(col)
        vals = None if vals is None else _f64(vals)
        h = vp()
        check(ctx.L.b2_csr_create(ctx.h, nrows, ncols, _ptr(rowptr), _ptr(col), _ptr(vals), ctypes.byref(h)))
        return cls(ctx, h)

    @classmethod
    def from_elements(cls, ctx, nrows, dof):
        dof = _i32(dof)
        h = vp()
        check(ctx.L.b2_csr_create_from_elements(ctx.h, nrows, dof.shape[0], dof.shape[1], _ptr(dof), ctypes.byref(h)))
        return cls(ctx, h)

    def __del__(self):
        try:
            if self.h and self.ctx.h:
                self.L.b2_csr_destroy(self.h)
        except Exception:
            pass

    @property
    def shape(self):
        return int(self.L.b2_csr_nrows(self.h)), int(self.L.b2_csr_ncols(self.h))

    @property
    def nnz(self):
        return int(self.L.b2_csr_nnz(self.h))

    def get(self, structure=True, values=True):
        n, nnz = self.shape[0], self.nnz
        rp = np.empty(n + 1, dtype=np.int64) if structure else None
        ci_ = np.empty(nnz, dtype=np.int32) if structure else None
        v = np.empty(nnz) if values else None
        check(self.L.b2_csr_get(self.h, _ptr(rp), _ptr(ci_), _ptr(v)))
        return rp, ci_, v

    def to_scipy(self):
        import scipy.sparse as sp
        rp, ci_, v = self.get()
        return sp.csr_matrix((v, ci_, rp), shape=self.shape)

    def put_vals(self, v):
        v = _f64(v)
        assert v.shape[0] == self.nnz
        check(self.L.b2_csr_put_vals(self.h, _ptr(v)))

    def zero(self):
        check(self.L.b2_csr_zero(self.h))

    def copy_vals_from(self, other):
        check(self.L.b2_csr_copy_vals(self.h, other.h))

    def add_blocks(self, rows, cols, vals):
        rows, cols, vals = _i32(rows), _i32(cols), _f64(vals)
        nblk, nrow = rows.shape
        ncol = cols.shape[1]
        check(self.L.b2_csr_add_blocks(self.h, nblk, nrow, ncol, _ptr(rows), _ptr(cols), _ptr(vals)))

    def set_rows(self, rows, ptr, cols, vals):
        rows, ptr, cols, vals = _i32(rows), _i64(ptr), _i32(cols), _f64(vals)
        check(self.L.b2_csr_set_rows(self.h, rows.shape[0], _ptr(rows), _ptr(ptr), _ptr(cols), _ptr(vals)))

    def zero_rows(self, rows, diag):
        rows = _i32(rows)
        check(self.L.b2_csr_zero_rows(self.h, _ptr(rows), rows.shape[0], float(diag)))

    def zero_cols(self, cols):
        cols = _i32(cols)
        check(self.L.b2_csr_zero_cols(self.h, _ptr(cols), cols.shape[0]))

    def diag(self, d):
        check(self.L.b2_csr_diag(self.h, d.h))

    def transpose(self):
        h = vp()
        check(self.L.b2_csr_transpose(self.h, ctypes.byref(h)))
        return Csr(self.ctx, h)

    def matmat(self, B):
        """self * B as a new matrix (b2_csr_matmat)."""
        h = vp()
        check(self.L.b2_csr_matmat(self.h, B.h, ctypes.byref(h)))
        return Csr(self.ctx, h)

    def axpy(self, a, X):
        """self += a X, the pattern of X inside the pattern of self (b2_csr_axpy)."""
        check(self.L.b2_csr_axpy(self.h, float(a), X.h))

    def pattern_contains(self, X):
        r = ci()
        check(self.L.b2_csr_pattern_contains(self.h, X.h, ctypes.byref(r)))
        return bool(r.value)

    def spmv(self, x, y):
        check(self.L.b2_csr_spmv(self.h, x.h, y.h))

    def spmv_add(self, x, y):
        check(self.L.b2_csr_spmv_add(self.h, x.h, y.h))

    def spmv_t(self, x, y):
        check(self.L.b2_csr_spmv_t(self.h, x.h, y.h))

    def resid(self, b, x, r):
        check(self.L.b2_csr_resid(self.h, b.h, x.h, r.h))

    def jacobi_sweep(self, dinv, b, xin, xout, omega):
        check(self.L.b2_csr_jacobi_sweep(self.h, dinv.h, b.h, xin.h, xout.h, float(omega)))

    def ptap(self, P, A):
        """self = P^T A P (numeric, onto self's pattern)."""
        check(self.L.b2_csr_ptap(P.h, A.h, self.h))


class Halo:
    """Distributed layout of a rank-local vector (b2_halo_*)."""

    def __init__(self, ctx, n_local, local_idx, packed_pos, n_packed, owned, mult):
        self.ctx, self.L = ctx, ctx.L
        local_idx, packed_pos = _i32(local_idx), _i32(packed_pos)
        owned = np.ascontiguousarray(owned, dtype=np.uint8)
        mult = np.ascontiguousarray(mult, dtype=np.uint8)
        assert owned.shape[0] == n_local and mult.shape[0] == n_local and local_idx.shape == packed_pos.shape
        h = vp()
        check(self.L.b2_halo_create(ctx.h, n_local, local_idx.shape[0], _ptr(local_idx), _ptr(packed_pos), n_packed,
                                    _ptr(owned), _ptr(mult), ctypes.byref(h)))
        self.h = h
        self.n_local = n_local

    def sum(self, v):
        check(self.L.b2_halo_sum(self.h, v.h))

    def set_exchange(self, hold_ptr, hold_rank, hold_pos, hold_spos):
        hold_rank, hold_pos, hold_spos = _i32(hold_rank), _i32(hold_pos), _i32(hold_spos)
        hold_ptr = np.ascontiguousarray(hold_ptr, dtype=np.int64)
        check(self.L.b2_halo_set_exchange(self.h, _ptr(hold_ptr), _ptr(hold_rank), _ptr(hold_pos), _ptr(hold_spos)))

    def owned_count(self):
        return int(self.L.b2_halo_owned_count(self.h))

    def __del__(self):
        try:
            if self.h and self.ctx.h:
                self.L.b2_halo_destroy(self.h)
        except Exception:
            pass


class Galerkin:
    """Element-gather Galerkin product Ac = P^T Af P (b2_galerkin_*)."""

    def __init__(self, Af, Ac, fine_dofs, coarse_dofs, ploc, fine_entity, valence, fine_mask=None, coarse_mask=None):
        self.ctx, self.L, self.Af, self.Ac = Af.ctx, Af.ctx.L, Af, Ac
        fine_dofs, coarse_dofs, ploc = _i32(fine_dofs), _i32(coarse_dofs), _f64(ploc)
        u8 = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.uint8)
        fine_entity, valence, fine_mask, coarse_mask = u8(fine_entity), u8(valence), u8(fine_mask), u8(coarse_mask)
        assert fine_dofs.shape[0] == coarse_dofs.shape[0] == valence.shape[0] and valence.shape[1] == 27
        assert ploc.shape == (fine_dofs.shape[1], coarse_dofs.shape[1])
        assert fine_mask is None or fine_mask.shape[0] == Af.shape[0]
        assert coarse_mask is None or coarse_mask.shape[0] == Ac.shape[0]
        h = vp()
        check(self.L.b2_galerkin_create(Af.h, Ac.h, fine_dofs.shape[0], fine_dofs.shape[1], coarse_dofs.shape[1],
                                        _ptr(fine_dofs), _ptr(coarse_dofs), _ptr(ploc), _ptr(fine_entity), _ptr(valence),
                                        _ptr(fine_mask), _ptr(coarse_mask), ctypes.byref(h)))
        self.h = h

    def apply(self):
        check(self.L.b2_galerkin_apply(self.h))

    def record_elements(self, on=True):
        check(self.L.b2_galerkin_record_elements(self.h, 1 if on else 0))

    def apply_from_elements(self, finer):
        """Ac = P^T Af P from the element matrices recorded by `finer` (whose coarse matrix is our Af)."""
        check(self.L.b2_galerkin_apply_from_elements(self.h, finer.h))

    def __del__(self):
        try:
            if self.h and self.ctx.h:
                self.L.b2_galerkin_destroy(self.h)
        except Exception:
            pass


class Mesh:
    def __init__(self, ctx, xyz, conn):
        self.ctx, self.L = ctx, ctx.L
        xyz, conn = _f64(xyz), _i32(conn)
        assert xyz.shape[0] == 3 and conn.shape[1] == 27
        h = vp()
        check(self.L.b2_mesh_create(ctx.h, xyz.shape[1], conn.shape[0], _ptr(xyz), _ptr(conn), ctypes.byref(h)))
        self.h = h
        self.nel, self.nnode = conn.shape[0], xyz.shape[1]

    def prefetch(self, xyz_ptr=None, conn_ptr=None):
        """Upload the next step's mesh into the shadow buffers on the copy stream."""
        check(self.L.b2_mesh_prefetch(self.h, xyz_ptr, conn_ptr))

    def swap(self):
        check(self.L.b2_mesh_swap(self.h))

    def update(self, xyz_ptr=None, conn_ptr=None):
        """Asynchronous re-upload from (pinned) host pointers."""
        check(self.L.b2_mesh_update(self.h, xyz_ptr, conn_ptr))

    def __del__(self):
        try:
            if self.h and self.ctx.h:
                self.L.b2_mesh_destroy(self.h)
        except Exception:
            pass


class Assembler:
    def __init__(self, mesh, A, dof, tables):
        """tables = (phi, dxi, deta, dzeta, w) with phi.shape = [ngauss, nve]."""
        self.ctx, self.L, self.mesh, self.A = mesh.ctx, mesh.ctx.L, mesh, A
        dof = _i32(dof)
        phi, dxi, deta, dzeta, w = [_f64(t) for t in tables]
        h = vp()
        check(self.L.b2_asm_create(mesh.h, A.h, dof.shape[1], _ptr(dof), phi.shape[0], _ptr(phi), _ptr(dxi),
                                   _ptr(deta), _ptr(dzeta), _ptr(w), ctypes.byref(h)))
        self.h = h

    def __del__(self):
        try:
            if self.h and self.ctx.h:
                self.L.b2_asm_destroy(self.h)
        except Exception:
            pass

    def poisson(self, u=None, rhs=None, nu=1.0, fsrc=1.0):
        check(self.L.b2_asm_poisson(self.h, u.h if u is not None else None, rhs.h if rhs is not None else None,
                                    float(nu), float(fsrc)))

    def neumann(self, face_elem, face_local, face_value, face_tables, face_nodes, rhs):
        """rhs += boundary integrals of a constant flux over the listed faces (b2_asm_neumann)."""
        fe, fl, fv = _i32(face_elem), _i32(face_local), _f64(face_value)
        phi, dxi, deta, w = [_f64(t) for t in face_tables]
        fn = _i32(face_nodes)
        check(self.L.b2_asm_neumann(self.h, fe.shape[0], _ptr(fe), _ptr(fl), _ptr(fv), phi.shape[1], _ptr(phi), _ptr(dxi),
                                    _ptr(deta), _ptr(w), _ptr(fn), rhs.h))

    def neumann_faces(self, face_elem, face_local, face_value, face_tables, face_nodes, rhs):
        """The same for the faces of any element type, one face kind per call (b2_asm_neumann_faces): face_tables
        = (phi, dxi, deta [ngf][nvf], w[ngf]) of the triangle or quadrilateral, face_nodes[6][9] of the element type."""
        fe, fl, fv = _i32(face_elem), _i32(face_local), _f64(face_value)
        phi, dxi, deta, w = [_f64(t) for t in face_tables]
        fn = _i32(face_nodes)
        check(self.L.b2_asm_neumann_faces(self.h, fe.shape[0], _ptr(fe), _ptr(fl), _ptr(fv), phi.shape[1], phi.shape[0],
                                          _ptr(phi), _ptr(dxi), _ptr(deta), _ptr(w), _ptr(fn), rhs.h))

    def poisson_galerkin(self, gal, u=None, rhs=None, nu=1.0, fsrc=1.0):
        """Assembly fused with the Galerkin product of `gal` (Ac = P^T A P from the element matrices)."""
        check(self.L.b2_asm_poisson_galerkin(self.h, gal.h, u.h if u is not None else None,
                                             rhs.h if rhs is not None else None, float(nu), float(fsrc)))


class StokesAssembler:
    """Steady Stokes assembly plan (b2_stokes_*): three velocity components of one family + a pressure of another on a
    mesh of one element type; elem_dofs [nel][4][27] system dofs, tables of the two families (hostapi.elem_tables)."""

    def __init__(self, mesh, A, elem_dofs, tables_v, tables_p, navier_stokes=False):
        """navier_stokes: the plan also carries the velocity phi table (b2_ns_create) and offers assemble_ns."""
        self.ctx, self.L, self.mesh, self.A = mesh.ctx, mesh.ctx.L, mesh, A
        ed = _i32(elem_dofs)
        phiv, dxi, deta, dzeta, w = [_f64(t) for t in tables_v]
        phip = _f64(tables_p[0])
        h = vp()
        if navier_stokes:
            check(self.L.b2_ns_create(mesh.h, A.h, _ptr(ed), dxi.shape[1], phip.shape[1], dxi.shape[0], _ptr(phiv), _ptr(dxi), _ptr(deta),
                                      _ptr(dzeta), _ptr(w), _ptr(phip), ctypes.byref(h)))
        else:
            check(self.L.b2_stokes_create(mesh.h, A.h, _ptr(ed), dxi.shape[1], phip.shape[1], dxi.shape[0], _ptr(dxi), _ptr(deta), _ptr(dzeta),
                                          _ptr(w), _ptr(phip), ctypes.byref(h)))
        self.h = h

    def __del__(self):
        try:
            if self.h and self.ctx.h:
                self.L.b2_stokes_destroy(self.h)
        except Exception:
            pass

    def assemble(self, sol=None, rhs=None, IRe=1.0):
        check(self.L.b2_stokes_assemble(self.h, sol.h if sol is not None else None, rhs.h if rhs is not None else None, float(IRe)))

    def pressure_faces(self, face_elem, face_local, face_value, face_tables, face_nodes, rhs):
        """Boundary pressure term of the Navier-Stokes residual over faces of one kind (b2_ns_pressure_faces)."""
        fe, fl, fv = _i32(face_elem), _i32(face_local), _f64(face_value)
        phi, dxi, deta, w = [_f64(t) for t in face_tables]
        fn = _i32(face_nodes)
        check(self.L.b2_ns_pressure_faces(self.h, fe.shape[0], _ptr(fe), _ptr(fl), _ptr(fv), phi.shape[1], phi.shape[0], _ptr(phi), _ptr(dxi),
                                          _ptr(deta), _ptr(w), _ptr(fn), rhs.h))

    def assemble_ns(self, sol=None, rhs=None, nu=1.0):
        """Navier-Stokes residual RES = -aRes and exact Newton Jacobian (b2_ns_assemble)."""
        check(self.L.b2_ns_assemble(self.h, sol.h if sol is not None else None, rhs.h if rhs is not None else None, float(nu)))


class Schwarz:
    """Element-block multiplicative Schwarz preconditioner on a device operator (b2_schwarz_*): blocks as
    (blk_ptr, blk_dofs) with sorted dofs, schedule as (group_ptr, group_blocks) -- hostapi.AsmIndex / asm_schedule."""

    def __init__(self, ctx, A, blk_ptr, blk_dofs, group_ptr, group_blocks):
        self.ctx, self.L, self.A = ctx, ctx.L, A
        bp, bd = np.ascontiguousarray(blk_ptr, dtype=np.int64), _i32(blk_dofs)
        gp, gb = np.ascontiguousarray(group_ptr, dtype=np.int64), _i32(group_blocks)
        h = vp()
        check(self.L.b2_schwarz_create(ctx.h, A.h, bp.shape[0] - 1, _ptr(bp), _ptr(bd), gp.shape[0] - 1, _ptr(gp), _ptr(gb),
                                       ctypes.byref(h)))
        self.h = h

    def __del__(self):
        try:
            if self.h and self.ctx.h:
                self.L.b2_schwarz_destroy(self.h)
        except Exception:
            pass

    def set_subsolver(self, kind):
        """"lu": exact block solves (dense inverses); "ssor": one SSOR iteration per block (PCSOR default);
        "ilu": ILU(0) of every block in its sorted dofs (PCILU default)."""
        check(self.L.b2_schwarz_set_subsolver(self.h, {"lu": 0, "ssor": 1, "ilu": 2}[kind]))

    def set_row_levels(self, on=True):
        """SSOR / ILU(0): sort every block's rows into dependency levels and let all warps of the CTA work inside a level
        (same result bit for bit; for large blocks).  row_levels: the longest chain found by the setup."""
        check(self.L.b2_schwarz_set_row_levels(self.h, 1 if on else 0))

    @property
    def row_levels(self):
        return int(self.L.b2_schwarz_row_levels(self.h))

    def setup(self):
        check(self.L.b2_schwarz_setup(self.h))

    def apply(self, r, y):
        check(self.L.b2_schwarz_apply(self.h, r.h, y.h))

    @property
    def nbytes(self):
        return int(self.L.b2_schwarz_bytes(self.h))

    @property
    def ngroups(self):
        return int(self.L.b2_schwarz_groups(self.h))


class Multigrid:
    def __init__(self, ctx, nlevels):
        self.ctx, self.L = ctx, ctx.L
        h = vp()
        check(self.L.b2_mg_create(ctx.h, nlevels, ctypes.byref(h)))
        self.h = h
        self.nlevels = nlevels
        self._keep = []

    def __del__(self):
        try:
            if self.h and self.ctx.h:
                self.L.b2_mg_destroy(self.h)
        except Exception:
            pass

    def set_level(self, level, A, P, bdc_idx, npre=1, npost=1, omega=0.5):
        bdc_idx = _i32(bdc_idx)
        self._keep.append((A, P))
        check(self.L.b2_mg_set_level(self.h, level, A.h, P.h if P is not None else None, _ptr(bdc_idx),
                                     bdc_idx.shape[0], npre, npost, float(omega)))

    def set_level_halo(self, level, halo):
        self._keep.append(halo)
        check(self.L.b2_mg_set_level_halo(self.h, level, halo.h if halo is not None else None))

    def set_smoother(self, level, kind, emin=0.0, emax=0.0):
        """kind: "richardson" (Richardson+Jacobi) or "chebyshev" (Chebyshev+Jacobi; emax <= 0: estimated)."""
        check(self.L.b2_mg_set_smoother(self.h, level, {"richardson": 0, "chebyshev": 1}[kind], float(emin), float(emax)))

    def set_level_schwarz(self, level, schwarz):
        """Level smoother = Richardson(omega) + the element-block preconditioner (b2_mg_set_level_schwarz)."""
        self._keep.append(schwarz)
        check(self.L.b2_mg_set_level_schwarz(self.h, level, schwarz.h if schwarz is not None else None))

    def set_level_ksp(self, level, kind):
        """Level solver around the level's preconditioner: "richardson" or "gmres" (b2_mg_set_level_ksp)."""
        check(self.L.b2_mg_set_level_ksp(self.h, level, {"richardson": 0, "gmres": 1}[kind]))

    def set_coarse_schwarz(self, schwarz):
        """Direct coarse solve through a one-block exact Schwarz object (b2_mg_set_coarse_schwarz)."""
        self._keep.append(schwarz)
        check(self.L.b2_mg_set_coarse_schwarz(self.h, schwarz.h if schwarz is not None else None))

    def level_bounds(self, level):
        a, b = cd(), cd()
        check(self.L.b2_mg_level_bounds(self.h, level, ctypes.byref(a), ctypes.byref(b)))
        return a.value, b.value

    def set_coarse(self, rtol=1e-14, maxit=5000):
        check(self.L.b2_mg_set_coarse(self.h, float(rtol), int(maxit)))

    def vcycle(self, rhs, x):
        check(self.L.b2_mg_vcycle(self.h, rhs.h, x.h))

    def solve(self, res, eps):
        check(self.L.b2_mg_solve(self.h, res.h, eps.h))

    def coarse_iterations(self):
        return int(self.L.b2_mg_coarse_iterations(self.h))

    def set_level_nullspace(self, level, nvec):
        """nvec: Vector spanning the null space of the level operator (copied and normalised), or None."""
        check(self.L.b2_mg_set_level_nullspace(self.h, level, nvec.h if nvec is not None else None))

    def set_timing(self, on=True):
        check(self.L.b2_mg_set_timing(self.h, 1 if on else 0))

    def get_timing(self, nlevels):
        """ms[level][phase] since the last call; phases: pre-smoothing, residual, restriction, coarse solve, prolongation,
        post-smoothing (b2_mg_get_timing)."""
        out = np.zeros((nlevels, 6))
        check(self.L.b2_mg_get_timing(self.h, _ptr(out)))
        return out
