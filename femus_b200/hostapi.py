"""ctypes binding of include/femus_b200_host.h: the C++ host layer (mesh hierarchy, dof maps,
Dirichlet flags, prolongators, FE tables) as numpy arrays.  Harness only."""
import ctypes
import numpy as np

from .capi import lib

vp, ci, i64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64
LINEAR, SERENDIPITY, BIQUADRATIC = 0, 1, 2
HEX, TET, WEDGE = 0, 1, 2
FAMILY = {"linear": 0, "quadratic": 1, "biquadratic": 2}

_ready = False


def _L():
    global _ready
    L = lib()
    if not _ready:
        P = {
            "b2h_hier_create": (vp, [ci, ci, ci, ci, vp, ci]),
            "b2h_hier_create_local": (vp, [ci, ci, ci, ci, vp, ci, ci]),
            "b2h_hier_create_from_neu": (vp, [ctypes.c_char_p, ci, ctypes.c_double]),
            "b2h_hier_destroy": (None, [vp]),
            "b2h_level_ijk": (vp, [vp, ci]),
            "b2h_level_interface_nodes": (i64, [vp, ci, vp]),
            "b2h_hier_nlevels": (ci, [vp]),
            "b2h_hier_nprocs": (ci, [vp]),
            "b2h_level_nel": (i64, [vp, ci]),
            "b2h_level_nnode": (i64, [vp, ci]),
            "b2h_level_conn": (vp, [vp, ci]),
            "b2h_level_face": (vp, [vp, ci]),
            "b2h_level_part": (vp, [vp, ci]),
            "b2h_level_xyz": (vp, [vp, ci]),
            "b2h_level_child_el": (vp, [vp, ci]),
            "b2h_level_offsets": (None, [vp, ci, vp, vp]),
            "b2h_level_ndofs": (i64, [vp, ci, ci]),
            "b2h_level_system_dofs": (None, [vp, ci, ci, vp]),
            "b2h_level_bdc": (None, [vp, ci, ci, vp, vp]),
            "b2h_prolongator_create": (vp, [vp, ci, ci]),
            "b2h_csr_destroy": (None, [vp]),
            "b2h_csr_nrows": (i64, [vp]),
            "b2h_csr_ncols": (i64, [vp]),
            "b2h_csr_nnz": (i64, [vp]),
            "b2h_csr_rowptr": (vp, [vp]),
            "b2h_csr_col": (vp, [vp]),
            "b2h_csr_val": (vp, [vp]),
            "b2h_galerkin_nf": (ci, [ci]),
            "b2h_galerkin_element": (None, [ci, vp, vp]),
            "b2h_galerkin_maps": (ci, [vp, ci, ci, i64, i64, vp, vp]),
            "b2h_level_elem_type": (ci, [vp, ci]),
            "b2h_elem_nve": (ci, [ci, ci]),
            "b2h_elem_ngauss": (ci, [ci]),
            "b2h_elem_tables": (None, [ci, ci, vp, vp, vp, vp, vp]),
            "b2h_elem_prolongator_row": (ci, [ci, ci, ci, ci, vp, vp]),
            "b2h_elem_child_face": (ci, [ci, ci, ci]),
            "b2h_level_elem_types": (None, [vp, ci, vp]),
            "b2h_level_system_dofs27": (None, [vp, ci, ci, vp]),
            "b2h_sparsity_create": (vp, [vp, ci, ci]),
            "b2h_hier_create_general": (vp, [ci, ci, ci, ci]),
            "b2h_hex_nve": (ci, [ci]),
            "b2h_hex_tables": (None, [ci, vp, vp, vp, vp, vp]),
            "b2h_hex_prolongator_row": (ci, [ci, ci, ci, ci, vp, vp]),
            "b2h_face_nvf": (ci, [ci]),
            "b2h_face_tables": (None, [ci, vp, vp, vp, vp]),
            "b2h_hex_face_nodes": (None, [vp]),
            "b2h_level_boundary_faces": (i64, [vp, ci, vp, vp, vp]),
            "b2h_asm_create": (vp, [vp, ci, ci, ci, ci]),
            "b2h_asm_create_system": (vp, [vp, ci, ci, vp, ci, ci, ci]),
            "b2h_system_offsets": (ci, [vp, ci, ci, vp, vp]),
            "b2h_system_elem_dofs": (None, [vp, ci, ci, vp, vp]),
            "b2h_system_sparsity_create": (vp, [vp, ci, ci, vp, vp]),
            "b2h_system_prolongator_create": (vp, [vp, ci, ci, vp]),
            "b2h_system_bdc": (ci, [vp, ci, ci, vp, vp, vp]),
            "b2h_asm_destroy": (None, [vp]),
            "b2h_asm_nblocks": (i64, [vp]),
            "b2h_asm_block_type_range": (None, [vp, vp]),
            "b2h_asm_elem_ptr": (vp, [vp]),
            "b2h_asm_elems": (vp, [vp]),
            "b2h_asm_local_ptr": (vp, [vp]),
            "b2h_asm_local": (vp, [vp]),
            "b2h_asm_overlap_ptr": (vp, [vp]),
            "b2h_asm_overlap": (vp, [vp]),
            "b2h_asm_schedule": (i64, [i64, vp, vp, i64, vp, vp, ci, vp]),
            "b2h_last_error": (ctypes.c_char_p, []),
            "b2h_face_kind_ngauss": (ci, [ci]),
            "b2h_face_kind_ndofs": (ci, [ci, ci]),
            "b2h_face_kind_tables": (None, [ci, ci, vp, vp, vp, vp]),
            "b2h_elem_face_nodes": (None, [ci, vp]),
            "b2h_elem_face_kind": (ci, [ci, ci]),
        }
        for n, (r, a) in P.items():
            f = getattr(L, n)
            f.restype, f.argtypes = r, a
        _ready = True
    return L


def _view(ptr, shape, dtype):
    n = int(np.prod(shape))
    if n == 0 or not ptr:
        return np.zeros(shape, dtype=dtype)
    buf = (ctypes.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype).reshape(shape)


def _fam(f):
    return FAMILY[f] if isinstance(f, str) else int(f)


class HostLevel:
    def __init__(self, hier, l):
        L, h = hier.L, hier.h
        self.hier, self.l = hier, l
        self.nel = int(L.b2h_level_nel(h, l))
        self.nnode = int(L.b2h_level_nnode(h, l))
        self.elem_type = int(L.b2h_level_elem_type(h, l))      # -1: several element types (see elem_types)
        self.elem_types = np.zeros(self.nel, dtype=np.uint8)
        L.b2h_level_elem_types(h, l, self.elem_types.ctypes.data_as(vp))
        self.conn = _view(L.b2h_level_conn(h, l), (self.nel, 27), np.int32)
        self.face = _view(L.b2h_level_face(h, l), (self.nel, 6), np.int32)
        self.part = _view(L.b2h_level_part(h, l), (self.nel,), np.int32)
        self.xyz = _view(L.b2h_level_xyz(h, l), (3, self.nnode), np.float64)
        p_ijk = L.b2h_level_ijk(h, l)      # lattice names of the nodes: box meshes only
        self.ijk = _view(p_ijk, (3, self.nnode), np.int32) if p_ijk else None
        np1 = hier.nprocs + 1
        eo = np.zeros(np1, dtype=np.int64)
        do = np.zeros((3, np1), dtype=np.int64)
        L.b2h_level_offsets(h, l, eo.ctypes.data_as(vp), do.ctypes.data_as(vp))
        self.elem_offset, self.dof_offset = eo, do

    @property
    def child_el(self):
        p = self.hier.L.b2h_level_child_el(self.hier.h, self.l)
        return _view(p, (self.nel, 8), np.int32) if p else None

    def ndofs(self, family):
        return int(self.hier.L.b2h_level_ndofs(self.hier.h, self.l, _fam(family)))

    def interface_nodes(self):
        """Sorted nodes on faces shared with the sub-meshes of other ranks (empty on a complete mesh)."""
        n = int(self.hier.L.b2h_level_interface_nodes(self.hier.h, self.l, None))
        out = np.zeros(n, dtype=np.int32)
        if n:
            self.hier.L.b2h_level_interface_nodes(self.hier.h, self.l, out.ctypes.data_as(vp))
        return out

    def lattice_key(self, nodes=None):
        """Rank-independent int64 name of the nodes: i + SX (j + SY k) on the level's global lattice."""
        sx, sy, _ = self.hier.lattice_dims(self.l)
        ijk = self.ijk if nodes is None else self.ijk[:, nodes]
        return ijk[0].astype(np.int64) + sx * (ijk[1].astype(np.int64) + sy * ijk[2].astype(np.int64))

    def system_dofs27(self, family):
        """GetSystemDof of every element in rows of 27, padded with -1 (meshes of several element types)."""
        out = np.zeros((self.nel, 27), dtype=np.int32)
        self.hier.L.b2h_level_system_dofs27(self.hier.h, self.l, _fam(family), out.ctypes.data_as(vp))
        return out

    def sparsity(self, family):
        """(rowptr, col) of the system matrix, built on the host (GetSparsityPatternSize)."""
        L = self.hier.L
        p = L.b2h_sparsity_create(self.hier.h, self.l, _fam(family))
        n, nnz = int(L.b2h_csr_nrows(p)), int(L.b2h_csr_nnz(p))
        rp = _view(L.b2h_csr_rowptr(p), (n + 1,), np.int64).copy()
        ci_ = _view(L.b2h_csr_col(p), (nnz,), np.int32).copy()
        L.b2h_csr_destroy(p)
        return rp, ci_

    def system_dofs(self, family):
        f = _fam(family)
        if self.elem_type < 0:
            raise ValueError("mesh of several element types: use system_dofs27")
        out = np.zeros((self.nel, self.hier.L.b2h_elem_nve(self.elem_type, f)), dtype=np.int32)
        self.hier.L.b2h_level_system_dofs(self.hier.h, self.l, f, out.ctypes.data_as(vp))
        return out

    def boundary_faces(self):
        """(element, local face, boundary index) of every boundary face of the level."""
        L, h = self.hier.L, self.hier.h
        n = L.b2h_level_boundary_faces(h, self.l, None, None, None)
        e, f, b = (np.zeros(n, dtype=np.int32) for _ in range(3))
        L.b2h_level_boundary_faces(h, self.l, e.ctypes.data_as(vp), f.ctypes.data_as(vp), b.ctypes.data_as(vp))
        return e, f, b

    def bdc(self, family, dirichlet_faces=(1, 2, 3, 4, 5, 6)):
        flags = np.zeros(7, dtype=np.int32)
        flags[list(dirichlet_faces)] = 1
        out = np.zeros(self.ndofs(family))
        self.hier.L.b2h_level_bdc(self.hier.h, self.l, _fam(family), flags.ctypes.data_as(vp), out.ctypes.data_as(vp))
        return out


class HostHierarchy:
    """MultiLevelMesh of the host layer: GenerateCoarseBoxMesh + RefineMesh."""

    def __init__(self, nx, ny, nz, nlevels, bounds=None, nprocs=1, local_rank=None):
        """local_rank=None: the complete mesh (numbered for `nprocs` ranks, z-slabs).
        local_rank=r: only rank r's sub-mesh of that partition, locally numbered and refined."""
        self.L = _L()
        b = None if bounds is None else np.ascontiguousarray(bounds, dtype=np.float64)
        bp = None if b is None else b.ctypes.data_as(vp)
        self.box = (nx, ny, nz)
        if local_rank is None:
            self.h = self.L.b2h_hier_create(nx, ny, nz, nlevels, bp, nprocs)
            self.nprocs = nprocs
        else:
            self.h = self.L.b2h_hier_create_local(nx, ny, nz, nlevels, bp, nprocs, local_rank)
            self.nprocs = 1
        if not self.h:
            raise ValueError("b2h_hier_create failed")
        self.nlevels = nlevels
        self.levels = [HostLevel(self, l) for l in range(nlevels)]

    @classmethod
    def from_neu(cls, path, nlevels, Lref=1.0):
        """MultiLevelMesh::ReadCoarseMesh on a Gambit .neu file (27-node hexahedra, 10-node tetrahedra, 18-node
        wedges, also mixed) + RefineMesh."""
        self = cls.__new__(cls)
        self.L = _L()
        self.box = None
        self.nprocs = 1
        self.h = self.L.b2h_hier_create_from_neu(str(path).encode(), nlevels, float(Lref))
        if not self.h:
            raise ValueError("b2h_hier_create_from_neu failed")
        self.nlevels = nlevels
        self.levels = [HostLevel(self, l) for l in range(nlevels)]
        return self

    @classmethod
    def box_general(cls, nx, ny, nz, nlevels):
        """Test hook: a generated box of hexahedra refined by the general (any element type) code path."""
        self = cls.__new__(cls)
        self.L = _L()
        self.box = (nx, ny, nz)
        self.nprocs = 1
        self.h = self.L.b2h_hier_create_general(nx, ny, nz, nlevels)
        if not self.h:
            raise ValueError("b2h_hier_create_general failed")
        self.nlevels = nlevels
        self.levels = [HostLevel(self, l) for l in range(nlevels)]
        return self

    def __del__(self):
        try:
            if self.h:
                self.levels = []
                self.L.b2h_hier_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def lattice_dims(self, l):
        nx, ny, nz = self.box
        f = 2 ** (l + 1)
        return nx * f + 1, ny * f + 1, nz * f + 1

    def galerkin_maps(self, lcoarse, family, e0=0, e1=None):
        """(fine_dofs[ne][nf], valence[ne][27]) of the coarse elements [e0, e1) of level lcoarse."""
        f = _fam(family)
        nelc = self.levels[lcoarse].nel
        e1 = nelc if e1 is None else e1
        nf = self.L.b2h_galerkin_nf(f)
        fd = np.zeros((e1 - e0, nf), dtype=np.int32)
        val = np.zeros((e1 - e0, 27), dtype=np.uint8)
        if self.L.b2h_galerkin_maps(self.h, lcoarse, f, e0, e1, fd.ctypes.data_as(vp), val.ctypes.data_as(vp)):
            raise ValueError("b2h_galerkin_maps: bad level or element range")
        return fd, val

    def prolongator(self, lfine, family):
        """(rowptr, col, val, shape) of P from level lfine-1 to lfine; arrays are copies."""
        p = self.L.b2h_prolongator_create(self.h, lfine, _fam(family))
        if not p:
            raise ValueError("bad level")
        n, m, nnz = int(self.L.b2h_csr_nrows(p)), int(self.L.b2h_csr_ncols(p)), int(self.L.b2h_csr_nnz(p))
        rp = _view(self.L.b2h_csr_rowptr(p), (n + 1,), np.int64).copy()
        ci_ = _view(self.L.b2h_csr_col(p), (nnz,), np.int32).copy()
        v = _view(self.L.b2h_csr_val(p), (nnz,), np.float64).copy()
        self.L.b2h_csr_destroy(p)
        return rp, ci_, v, (n, m)


def galerkin_element(family):
    """(ploc[nf][nc], fine_entity[nf]) of the element-gather Galerkin product."""
    L = _L()
    f = _fam(family)
    nf, nc = L.b2h_galerkin_nf(f), L.b2h_hex_nve(f)
    ploc = np.zeros((nf, nc))
    ent = np.zeros(nf, dtype=np.uint8)
    L.b2h_galerkin_element(f, ploc.ctypes.data_as(vp), ent.ctypes.data_as(vp))
    return ploc, ent


def elem_nve(elem_type, family):
    return int(_L().b2h_elem_nve(elem_type, _fam(family)))


def elem_tables(elem_type, family):
    """(phi, dxi, deta, dzeta, w) of elem_type_3D(type, family, "seventh"): [ngauss][nve] and [ngauss]."""
    L = _L()
    f = _fam(family)
    nve, ng = L.b2h_elem_nve(elem_type, f), L.b2h_elem_ngauss(elem_type)
    t = [np.zeros((ng, nve)) for _ in range(4)] + [np.zeros(ng)]
    L.b2h_elem_tables(elem_type, f, *[a.ctypes.data_as(vp) for a in t])
    return tuple(t)


def elem_prolongator_row(elem_type, family, child, node):
    L = _L()
    idx = np.zeros(27, dtype=np.int32)
    val = np.zeros(27)
    n = L.b2h_elem_prolongator_row(elem_type, _fam(family), child, node, idx.ctypes.data_as(vp), val.ctypes.data_as(vp))
    return idx[:n].copy(), val[:n].copy()


def elem_child_face(elem_type, child, child_face):
    return int(_L().b2h_elem_child_face(elem_type, child, child_face))


def tet_prolongator_row(family, child, node):
    return elem_prolongator_row(TET, family, child, node)


def tet_child_face(child, child_face):
    return elem_child_face(TET, child, child_face)


def hex_tables(family):
    L = _L()
    f = _fam(family)
    nve = L.b2h_hex_nve(f)
    t = [np.zeros((64, nve)) for _ in range(4)] + [np.zeros(64)]
    L.b2h_hex_tables(f, *[a.ctypes.data_as(vp) for a in t])
    return tuple(t)


def face_tables(family):
    """(phi, dxi, deta, w) of the face element elem_type_2D("quad", family, "seventh"): [16][nvf], [16]."""
    L = _L()
    f = _fam(family)
    nvf = L.b2h_face_nvf(f)
    t = [np.zeros((16, nvf)) for _ in range(3)] + [np.zeros(16)]
    L.b2h_face_tables(f, *[a.ctypes.data_as(vp) for a in t])
    return tuple(t)


def hex_face_nodes():
    out = np.zeros((6, 9), dtype=np.int32)
    _L().b2h_hex_face_nodes(out.ctypes.data_as(vp))
    return out


QUAD_FACE, TRI_FACE = 0, 1


class AsmIndex:
    """Element blocks and index sets of the ASM / Vanka smoother on one level for one Lagrange variable without
    Schur variables (MeshASMPartitioning::DoPartition + LinearEquationSolverPetscAsm::BuildASMIndex): per block
    its elements, the sorted local and overlapping dof sets, as (ptr[nblocks+1], entries) pairs."""

    def __init__(self, level, family, block_elems, iproc=0, nschur=0):
        """family: one family name, or a list of them for a system of several variables (rows [rank][variable][dof]),
        the last `nschur` of which are Schur variables (Vanka blocks of velocity-pressure systems)."""
        L = level.hier.L
        if isinstance(family, (list, tuple)):
            fams = np.array([_fam(f) for f in family], dtype=np.int32)
            h = L.b2h_asm_create_system(level.hier.h, level.l, len(fams), fams.ctypes.data_as(vp), int(nschur), int(block_elems), int(iproc))
        else:
            h = L.b2h_asm_create(level.hier.h, level.l, _fam(family), int(block_elems), int(iproc))
        if not h:
            raise ValueError(L.b2h_last_error().decode())
        try:
            nb = int(L.b2h_asm_nblocks(h))
            self.nblocks = nb
            rng = np.zeros(3, dtype=np.int64)
            L.b2h_asm_block_type_range(h, rng.ctypes.data_as(vp))
            self.block_type_range = rng
            for name in ("elem", "local", "overlap"):
                ptr = _view(getattr(L, f"b2h_asm_{name}_ptr")(h), (nb + 1,), np.int64).copy()
                ent = getattr(L, "b2h_asm_" + {"elem": "elems", "local": "local", "overlap": "overlap"}[name])(h)
                setattr(self, name + "_ptr", ptr)
                setattr(self, name, _view(ent, (int(ptr[-1]),), np.int32).copy() if ptr[-1] else np.zeros(0, dtype=np.int32))
        finally:
            L.b2h_asm_destroy(h)

    def blocks(self, which="overlap"):
        ptr, ent = getattr(self, which + "_ptr"), getattr(self, which)
        return [ent[ptr[b]:ptr[b + 1]] for b in range(self.nblocks)]


def system_offsets(level, families):
    """KKoffset[nvars+1][nprocs] of a system of several variables on a level (LinearEquation::InitPde)."""
    fams = np.array([_fam(f) for f in families], dtype=np.int32)
    out = np.zeros((len(fams) + 1, level.hier.nprocs), dtype=np.int64)
    if level.hier.L.b2h_system_offsets(level.hier.h, level.l, len(fams), fams.ctypes.data_as(vp), out.ctypes.data_as(vp)):
        raise ValueError(level.hier.L.b2h_last_error().decode())
    return out


class SystemOnLevel:
    """A system of several Lagrange variables on a level, rows [rank][variable][dof] (LinearEquation::InitPde):
    element dof lists, sparsity pattern, prolongator from the level below, Dirichlet flags -- host side of SURVEY 8f row 3."""

    def __init__(self, level, families):
        self.level, self.families = level, list(families)
        self._f = np.array([_fam(f) for f in families], dtype=np.int32)
        self.offsets = system_offsets(level, families)
        self.n = int(self.offsets[-1, -1])

    def _csr(self, p, with_values):
        L = self.level.hier.L
        if not p:
            raise ValueError(L.b2h_last_error().decode())
        n, nnz = int(L.b2h_csr_nrows(p)), int(L.b2h_csr_nnz(p))
        rp = _view(L.b2h_csr_rowptr(p), (n + 1,), np.int64).copy()
        col = _view(L.b2h_csr_col(p), (nnz,), np.int32).copy()
        val = _view(L.b2h_csr_val(p), (nnz,), np.float64).copy() if with_values else None
        shape = (n, int(L.b2h_csr_ncols(p)))
        L.b2h_csr_destroy(p)
        return rp, col, val, shape

    def elem_dofs(self):
        """[nel][nvars][27] system dofs, -1 padded."""
        out = np.zeros((self.level.nel, len(self._f), 27), dtype=np.int32)
        self.level.hier.L.b2h_system_elem_dofs(self.level.hier.h, self.level.l, len(self._f), self._f.ctypes.data_as(vp), out.ctypes.data_as(vp))
        return out

    def sparsity(self, pattern=None):
        pat = None if pattern is None else np.ascontiguousarray(pattern, dtype=np.uint8)
        p = self.level.hier.L.b2h_system_sparsity_create(self.level.hier.h, self.level.l, len(self._f), self._f.ctypes.data_as(vp),
                                                         pat.ctypes.data_as(vp) if pat is not None else None)
        rp, col, _, _ = self._csr(p, False)
        return rp, col

    def prolongator(self):
        p = self.level.hier.L.b2h_system_prolongator_create(self.level.hier.h, self.level.l, len(self._f), self._f.ctypes.data_as(vp))
        return self._csr(p, True)

    def bdc(self, dirichlet_faces_per_var):
        """dirichlet_faces_per_var[k] = boundary sets (1..6) on which variable k is Dirichlet."""
        flags = np.zeros((len(self._f), 7), dtype=np.uint8)
        for k, faces in enumerate(dirichlet_faces_per_var):
            flags[k, list(faces)] = 1
        out = np.zeros(self.n)
        if self.level.hier.L.b2h_system_bdc(self.level.hier.h, self.level.l, len(self._f), self._f.ctypes.data_as(vp), flags.ctypes.data_as(vp),
                                            out.ctypes.data_as(vp)):
            raise ValueError(self.level.hier.L.b2h_last_error().decode())
        return out


def asm_schedule(rowptr, col, blk_ptr, blk_dofs, mode="colours"):
    """Groups of mutually independent blocks for the multiplicative sweep (b2h_asm_schedule): mode "levels" = the
    dependency levels of the given block order (the reference's sequential sweep exactly), "colours" = greedy
    colouring.  Returns (group_of_block, group_ptr, group_blocks): blocks in sweep order, cut into groups."""
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
    col = np.ascontiguousarray(col, dtype=np.int32)
    blk_ptr = np.ascontiguousarray(blk_ptr, dtype=np.int64)
    blk_dofs = np.ascontiguousarray(blk_dofs, dtype=np.int32)
    nb = blk_ptr.shape[0] - 1
    grp = np.zeros(nb, dtype=np.int32)
    L = _L()
    ng = L.b2h_asm_schedule(rowptr.shape[0] - 1, rowptr.ctypes.data_as(vp), col.ctypes.data_as(vp), nb, blk_ptr.ctypes.data_as(vp),
                            blk_dofs.ctypes.data_as(vp), {"levels": 0, "colours": 1}[mode], grp.ctypes.data_as(vp))
    if ng < 0:
        raise ValueError(L.b2h_last_error().decode())
    order = np.argsort(grp, kind="stable").astype(np.int32)
    gptr = np.zeros(ng + 1, dtype=np.int64)
    np.cumsum(np.bincount(grp, minlength=ng), out=gptr[1:])
    return grp, gptr, order


def face_kind_tables(kind, family):
    """(phi, dxi, deta, w) of elem_type_2D("quad" | "tri", family, "seventh"): [ngauss][ndofs], [ngauss]
    (kind 0: quadrilateral, 4 / 8 / 9 dofs, 16 points; 1: triangle, 3 / 6 / 7 dofs, 13 points)."""
    L = _L()
    f = _fam(family)
    ng, nvf = L.b2h_face_kind_ngauss(kind), L.b2h_face_kind_ndofs(kind, f)
    t = [np.zeros((ng, nvf)) for _ in range(3)] + [np.zeros(ng)]
    L.b2h_face_kind_tables(kind, f, *[a.ctypes.data_as(vp) for a in t])
    return tuple(t)


def elem_face_nodes(elem_type):
    """[6][9] element-local nodes of the faces of an element type, -1 where no face / entry exists."""
    out = np.zeros((6, 9), dtype=np.int32)
    _L().b2h_elem_face_nodes(elem_type, out.ctypes.data_as(vp))
    return out


def elem_face_kinds(elem_type):
    """face kind (0 quadrilateral, 1 triangle) of the 6 local faces, -1 past the last one."""
    L = _L()
    return np.array([L.b2h_elem_face_kind(elem_type, f) for f in range(6)], dtype=np.int32)


def hex_prolongator_row(family, a, b, c):
    L = _L()
    idx = np.zeros(27, dtype=np.int32)
    val = np.zeros(27)
    n = L.b2h_hex_prolongator_row(_fam(family), a, b, c, idx.ctypes.data_as(vp), val.ctypes.data_as(vp))
    return idx[:n].copy(), val[:n].copy()
