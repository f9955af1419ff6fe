"""Oracle (TEST INFRASTRUCTURE ONLY): ctypes access to oracle/_ref/libfemus_fe_ref.so, the
reference's own FE kernel compiled in place (recipe: oracle/ref_capi/Makefile).  The library is
built only where /root/reference exists; on the GPU box the prebuilt file travels with the repo."""
import ctypes
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libfemus_fe_ref.so")
_lib = None


def build(reference="/root/reference"):
    """Compile oracle/_ref from the reference sources if they are present.  Returns True if the
    library exists afterwards."""
    if os.path.isdir(os.path.join(reference, "src", "02_reference_geom_elements")):
        subprocess.run(["make", "-C", os.path.join(_HERE, "ref_capi"), f"REF={reference}"],
                       check=True, stdout=subprocess.DEVNULL)
    return os.path.exists(LIB_PATH)


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(LIB_PATH)
        vp, ci, cd = ctypes.c_void_p, ctypes.c_int, ctypes.c_double
        L.fref_create.restype = vp
        L.fref_create.argtypes = [ctypes.c_char_p] * 3
        L.fref_destroy.argtypes = [vp]
        for f in ("fref_ndofs", "fref_ngauss", "fref_ndofs_fine"):
            getattr(L, f).argtypes = [vp]
        L.fref_gauss.argtypes = [vp, vp, vp]
        L.fref_tables.argtypes = [vp] * 5
        L.fref_jacobian.argtypes = [vp, vp, ci, ci, vp, vp, vp, vp]
        L.fref_prol_row.argtypes = [vp, ci, vp, vp, vp, vp]
        L.fref_poisson_element.argtypes = [vp, vp, vp, cd, vp, vp]
        L.fref_poisson_assemble_csr.restype = cd
        L.fref_poisson_assemble_csr.argtypes = [vp, ctypes.c_long, ctypes.c_long, vp, vp, vp, ctypes.c_long,
                                                vp, vp, vp, vp, vp, cd, ci]
        L.fref2_create.restype = vp
        L.fref2_create.argtypes = [ctypes.c_char_p] * 3
        L.fref2_destroy.argtypes = [vp]
        L.fref2_ndofs.argtypes = [vp]
        L.fref2_ngauss.argtypes = [vp]
        L.fref2_gauss.argtypes = [vp, vp, vp]
        L.fref2_tables.argtypes = [vp] * 4
        L.fref2_jacobian_sur.argtypes = [vp, vp, ci, ci, vp, vp, vp]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class RefElem:
    """elem_type_3D(geom, order, gauss) of the reference; geom = "hex" | "tet" | "wedge"."""

    def __init__(self, geom="hex", order="biquadratic", gauss="seventh"):
        self.L = lib()
        self.h = ctypes.c_void_p(self.L.fref_create(geom.encode(), order.encode(), gauss.encode()))
        self.n = self.L.fref_ndofs(self.h)
        self.ng = self.L.fref_ngauss(self.h)
        self.nf = self.L.fref_ndofs_fine(self.h)

    def gauss(self):
        w = np.zeros(self.ng)
        xi = np.zeros((3, self.ng))
        self.L.fref_gauss(self.h, _p(w), _p(xi))
        return w, xi.T.copy()

    def tables(self):
        t = [np.zeros((self.ng, self.n)) for _ in range(4)]
        self.L.fref_tables(self.h, *[_p(a) for a in t])
        return t

    def jacobian(self, X, ig):
        X = np.ascontiguousarray(X, dtype=np.float64)          # [3][n]
        w = ctypes.c_double()
        phi = np.zeros(self.n)
        g = np.zeros((self.n, 3))
        nb = np.zeros((self.n, 6))
        self.L.fref_jacobian(self.h, _p(X), X.shape[1], ig, ctypes.byref(w), _p(phi), _p(g), _p(nb))
        return w.value, phi, g

    def prolongator(self):
        """rows: list of (child, node, idx[], val[]) for the nf fine dofs."""
        rows = []
        idx = np.zeros(64, dtype=np.int32)
        val = np.zeros(64)
        for i in range(self.nf):
            ch, nd = ctypes.c_int(), ctypes.c_int()
            nc = self.L.fref_prol_row(self.h, i, _p(idx), _p(val), ctypes.byref(ch), ctypes.byref(nd))
            rows.append((ch.value, nd.value, idx[:nc].copy(), val[:nc].copy()))
        return rows

    def poisson_element(self, X, U, fsrc=1.0):
        X = np.ascontiguousarray(X, dtype=np.float64)
        U = np.ascontiguousarray(U, dtype=np.float64)
        F = np.zeros(self.n)
        B = np.zeros((self.n, self.n))
        self.L.fref_poisson_element(self.h, _p(X), _p(U), float(fsrc), _p(F), _p(B))
        return F, B

    def assemble_csr(self, conn, dof, xyz, sol, rowptr, col, fsrc=1.0, nthreads=1, e0=0, e1=None):
        """Assemble the Poisson matrix/residual over elements [e0,e1) into a CSR with the given
        pattern.  Returns (vals, rhs, seconds)."""
        conn = np.ascontiguousarray(conn, dtype=np.int32)
        dof = np.ascontiguousarray(dof, dtype=np.int32)
        xyz = np.ascontiguousarray(xyz, dtype=np.float64)       # [3][nnode]
        sol = np.ascontiguousarray(sol, dtype=np.float64)
        rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
        col = np.ascontiguousarray(col, dtype=np.int32)
        vals = np.zeros(col.shape[0])
        rhs = np.zeros(rowptr.shape[0] - 1)
        if e1 is None:
            e1 = conn.shape[0]
        sec = self.L.fref_poisson_assemble_csr(self.h, e0, e1, _p(conn), _p(dof), _p(xyz), xyz.shape[1], _p(sol),
                                               _p(rowptr), _p(col), _p(vals), _p(rhs), float(fsrc), int(nthreads))
        return vals, rhs, sec


class RefHex(RefElem):
    """elem_type_3D("hex", order, gauss) of the reference."""

    def __init__(self, order="biquadratic", gauss="seventh"):
        super().__init__("hex", order, gauss)


class RefTet(RefElem):
    """elem_type_3D("tet", order, gauss) of the reference (order: linear 4, quadratic 10, biquadratic 15)."""

    def __init__(self, order="quadratic", gauss="seventh"):
        super().__init__("tet", order, gauss)


class RefFace:
    """elem_type_2D(geom, order, gauss) of the reference, geom = "quad" | "tri": the face elements of the 3-D
    elements (quadrilaterals with 4 / 8 / 9 dofs, triangles with 3 / 6 / 7)."""

    def __init__(self, geom="quad", order="biquadratic", gauss="seventh"):
        self.L = lib()
        self.h = ctypes.c_void_p(self.L.fref2_create(geom.encode(), order.encode(), gauss.encode()))
        self.n = self.L.fref2_ndofs(self.h)
        self.ng = self.L.fref2_ngauss(self.h)

    def gauss(self):
        w = np.zeros(self.ng)
        xi = np.zeros((2, self.ng))
        self.L.fref2_gauss(self.h, _p(w), _p(xi))
        return w, xi.T.copy()

    def tables(self):
        t = [np.zeros((self.ng, self.n)) for _ in range(3)]
        self.L.fref2_tables(self.h, *[_p(a) for a in t])
        return t

    def jacobian_sur(self, X, ig):
        X = np.ascontiguousarray(X, dtype=np.float64)          # [3][n]
        w = ctypes.c_double()
        phi = np.zeros(self.n)
        nrm = np.zeros(3)
        self.L.fref2_jacobian_sur(self.h, _p(X), X.shape[1], ig, ctypes.byref(w), _p(phi), _p(nrm))
        return w.value, phi, nrm

    def __del__(self):
        try:
            self.L.fref2_destroy(self.h)
        except Exception:
            pass


class RefQuad(RefFace):
    """elem_type_2D("quad", order, gauss) of the reference: the face element of a hexahedron."""

    def __init__(self, order="biquadratic", gauss="seventh"):
        super().__init__("quad", order, gauss)
