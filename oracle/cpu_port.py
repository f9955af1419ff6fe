"""Oracle / CPU baseline (TEST INFRASTRUCTURE ONLY): the multigrid solve of oracle/mg.py with
its CSR kernels running in the OpenMP C port (oracle/cpu_port/spmv_port.c) instead of scipy, so
that the CPU baseline uses all host cores the way `mpirun -n <cores>` PETSc would.  Same operator
definitions as mg.Hierarchy (and therefore as SURVEY.md section 3.5); the coarse solve is the same
Jacobi-PCG the device path uses."""
import ctypes
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_port", "libfemus_port.so")
_lib = None
vp, ci, cd, i64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_int64


def build():
    subprocess.run(["make", "-C", os.path.join(_HERE, "cpu_port")], check=True, stdout=subprocess.DEVNULL)
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = ctypes.CDLL(LIB_PATH)
        L.port_spmv.argtypes = [ci, i64, vp, vp, vp, vp, vp, vp, vp, cd, ci]
        L.port_dot.restype = cd
        L.port_dot.argtypes = [i64, vp, vp, ci]
        L.port_axpby.argtypes = [i64, cd, vp, cd, vp, ci]
        L.port_pmult.argtypes = [i64, vp, vp, vp, ci]
        L.port_max_threads.restype = ci
        L.port_ptap.restype = ci
        L.port_ptap.argtypes = [i64, i64] + [vp] * 12 + [ci]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(vp)


class PortCsr:
    def __init__(self, A, nthreads):
        A = A.tocsr()
        A.sort_indices()
        self.n, self.m = A.shape
        self.rp = np.ascontiguousarray(A.indptr, dtype=np.int64)
        self.ci = np.ascontiguousarray(A.indices, dtype=np.int32)
        self.v = np.ascontiguousarray(A.data, dtype=np.float64)
        self.nt = nthreads
        self.L = lib()

    def spmv(self, x, y, mode=0, b=None, dinv=None, omega=0.0):
        self.L.port_spmv(mode, self.n, _p(self.rp), _p(self.ci), _p(self.v), _p(x), _p(b), _p(dinv), _p(y), omega, self.nt)


class PortMG:
    """V-cycle on the operators of an oracle mg.Hierarchy, kernels in OpenMP C."""

    def __init__(self, H, nthreads=None):
        self.L = lib()
        self.nt = nthreads or os.cpu_count()
        self.H = H
        nl = len(H.A)
        self.A = [PortCsr(A, self.nt) for A in H.A]
        self.P = [None] + [PortCsr(H.P[l], self.nt) for l in range(1, nl)]
        self.R = [None] + [PortCsr(H.P[l].T.tocsr(), self.nt) for l in range(1, nl)]
        self.dinv = [np.ascontiguousarray(d) for d in H.dinv]
        self.coarse_its = 0

    def dot(self, x, y):
        return self.L.port_dot(x.shape[0], _p(x), _p(y), self.nt)

    def coarse(self, b, rtol=1e-14, maxit=10000):
        A, dinv, bdc = self.A[0], self.dinv[0], self.H.bdc_idx[0]
        n = A.n
        x = np.zeros(n)
        x[bdc] = b[bdc]
        r = np.empty(n)
        A.spmv(x, r, 2, b=b)
        bb = self.dot(b, b)
        if bb == 0.0:
            return x
        z = dinv * r
        p = z.copy()
        q = np.empty(n)
        rz = self.dot(r, z)
        it = 0
        while it < maxit and self.dot(r, r) > rtol * rtol * bb:
            A.spmv(p, q, 0)
            alpha = rz / self.dot(p, q)
            self.L.port_axpby(n, alpha, _p(p), 1.0, _p(x), self.nt)
            self.L.port_axpby(n, -alpha, _p(q), 1.0, _p(r), self.nt)
            self.L.port_pmult(n, _p(dinv), _p(r), _p(z), self.nt)
            rzn = self.dot(r, z)
            self.L.port_axpby(n, 1.0, _p(z), rzn / rz, _p(p), self.nt)
            rz = rzn
            it += 1
        self.coarse_its = it
        return x

    def vcycle(self, l, b, npre=1, npost=1, omega=0.5):
        if l == 0:
            return self.coarse(b)
        A, dinv = self.A[l], self.dinv[l]
        x = omega * dinv * b                       # first sweep from a zero guess
        t = np.empty_like(x)
        for _ in range(npre - 1):
            A.spmv(x, t, 3, b=b, dinv=dinv, omega=omega)
            x, t = t, x
        r = np.empty_like(x)
        A.spmv(x, r, 2, b=b)
        bc = np.empty(self.R[l].n)
        self.R[l].spmv(r, bc, 0)
        xc = self.vcycle(l - 1, bc, npre, npost, omega)
        self.P[l].spmv(xc, x, 1)
        for _ in range(npost):
            A.spmv(x, t, 3, b=b, dinv=dinv, omega=omega)
            x, t = t, x
        return x

    def mg_solve(self, res, eps, npre=1, npost=1, omega=0.5):
        top = len(self.A) - 1
        res[self.H.bdc_idx[top]] = 0.0
        epsc = self.vcycle(top, res, npre, npost, omega)
        self.A[top].spmv(epsc, res, 2, b=res.copy())
        eps += epsc
        return res, eps


def ptap(P, A, rp, ci_, nthreads):
    """C = P^T A P on the pattern (rp, ci_) with the OpenMP port (scipy CSR in, scipy CSR out)."""
    import scipy.sparse as sp
    L = lib()
    P = P.tocsr(); A = A.tocsr(); R = P.T.tocsr()
    for M in (P, A, R):
        M.sort_indices()
    f = lambda M: (np.ascontiguousarray(M.indptr, dtype=np.int64), np.ascontiguousarray(M.indices, dtype=np.int32),
                   np.ascontiguousarray(M.data, dtype=np.float64))
    Ap, Ai, Ax = f(A); Pp, Pi, Px = f(P); Rp, Ri, Rx = f(R)
    rp = np.ascontiguousarray(rp, dtype=np.int64); ci_ = np.ascontiguousarray(ci_, dtype=np.int32)
    Cx = np.zeros(ci_.shape[0])
    err = L.port_ptap(P.shape[0], P.shape[1], _p(Ap), _p(Ai), _p(Ax), _p(Pp), _p(Pi), _p(Px), _p(Rp), _p(Ri), _p(Rx),
                      _p(rp), _p(ci_), _p(Cx), int(nthreads))
    if err:
        raise RuntimeError(f"port_ptap failed ({err}): product leaves the coarse pattern or out of memory")
    return sp.csr_matrix((Cx, ci_, rp), shape=(P.shape[1], P.shape[1]))
