"""Oracle (TEST INFRASTRUCTURE ONLY): meshes of hexahedra, tetrahedra and wedges -- also mixed -- read from
Gambit neutral files: FEMuS node/dof numbering, the face and centre nodes the file lacks, uniform 1 -> 8
refinement, Dirichlet flags, sparsity, Poisson assembly and prolongators, restated with numpy/scipy and plain
Python loops over ragged element lists (small meshes only).

PARITY PINNED TO REFERENCE OUTPUT (round 2): the reference's own sources, compiled unmodified on the single-process host
backend of oracle/ref_build and run here (applications/001_Poisson/main.cpp), produced tests/golden/ref_poisson_*.npz
(tests/golden/make_ref_golden.py); tests/test_reference_pin.py compares this module with them -- integers bit-exact,
values and printed residual norms to the stated tolerances.  (The FE arithmetic, fe_hex / fe_tet / fe_wedge, is pinned to the compiled
reference as well.)
Restates (paths relative to /root/reference/src):
  06_mesh/00_single_level/01_input/01_from_external_file/GambitIO.cpp:56-85, 92-352   file sections, permutations
  06_mesh/00_single_level/00_definition/Mesh.cpp:105-125, 1207-1333   AddBiquadraticNodesNotInMeshFile
  06_mesh/00_single_level/00_definition/Mesh.cpp:517-559              node renumbering by first visit
  06_mesh/00_single_level/03_refinement/MeshRefinement.cpp:188-507, 513-621; MeshRefinement.hpp:79-135
  06_solution/01_multiple_levels/00_definition/MultiLevelSolution.cpp:725-840        GenerateBdc
  08_algebra.../LinearEquation.cpp:407-548; 08_equations/00_stationary/LinearImplicitSystem.cpp:761-909, 1032-1120
The tables below are typed in from the reference's own (coarse2FineFaceMapping, Gambit permutations), NOT
derived geometrically as the product's host layer derives them, so that the two are independent.
"""
import numpy as np
import scipy.sparse as sp

from . import fe_hex, fe_tet, fe_wedge, mesh_box as mb

HEX, TET, WEDGE = 0, 1, 2
FAMILY = mb.FAMILY
NVE = {HEX: (8, 20, 27), TET: (4, 10, 15), WEDGE: (6, 15, 21)}
NFACES = {HEX: 6, TET: 4, WEDGE: 5}
FE = {HEX: fe_hex, TET: fe_tet, WEDGE: fe_wedge}
FACE_NODES = {HEX: [list(r) for r in fe_hex.FACE_NODES], TET: [list(r) for r in fe_tet.FACE_NODES], WEDGE: fe_wedge.FACE_NODES}
FACE_NVERT = {HEX: [4] * 6, TET: [3] * 4, WEDGE: fe_wedge.FACE_NVERT}
HEX_EDGES = [(0, 1), (1, 2), (2, 3), (3, 0), (4, 5), (5, 6), (6, 7), (7, 4), (0, 4), (1, 5), (2, 6), (3, 7)]
EDGES = {HEX: HEX_EDGES, TET: fe_tet.EDGES, WEDGE: fe_wedge.EDGES}
CHILD_VERTICES = {HEX: fe_hex.child_vertex_map(), TET: fe_tet.CHILD_VERTICES, WEDGE: fe_wedge.CHILD_VERTICES}
FILE_NVE = {27: HEX, 10: TET, 18: WEDGE}
GAMBIT_TO_FEMUS_VERTEX = {HEX: [4, 16, 0, 15, 23, 11, 7, 19, 3, 12, 20, 8, 25, 26, 24, 14, 22, 10, 5, 17, 1, 13, 21, 9, 6, 18, 2],
                          TET: [0, 4, 1, 6, 5, 2, 7, 8, 9, 3],
                          WEDGE: [3, 11, 5, 9, 10, 4, 12, 17, 14, 15, 16, 13, 0, 8, 2, 6, 7, 1]}
GAMBIT_TO_FEMUS_FACE = {HEX: [0, 4, 2, 5, 3, 1], TET: [0, 1, 2, 3], WEDGE: [2, 1, 0, 4, 3]}
# MeshRefinement.hpp:79-100: parent face -> its 4 (child, child face) pairs
COARSE_TO_FINE_FACE = {
    HEX: [[(0, 0), (1, 0), (4, 0), (5, 0)], [(1, 1), (2, 1), (5, 1), (6, 1)], [(2, 2), (3, 2), (6, 2), (7, 2)],
          [(3, 3), (0, 3), (7, 3), (4, 3)], [(0, 4), (1, 4), (2, 4), (3, 4)], [(4, 5), (5, 5), (6, 5), (7, 5)]],
    TET: [[(0, 0), (1, 0), (2, 0), (4, 0)], [(0, 1), (1, 1), (3, 1), (5, 1)], [(1, 2), (2, 2), (3, 2), (6, 2)], [(2, 3), (0, 3), (3, 3), (7, 3)]],
    WEDGE: [[(0, 0), (1, 0), (4, 0), (5, 0)], [(1, 1), (2, 1), (5, 1), (6, 1)], [(2, 2), (0, 2), (6, 2), (4, 2)],
            [(0, 3), (1, 3), (2, 3), (3, 3)], [(4, 4), (5, 4), (6, 4), (7, 4)]]}
# Mesh.cpp:105-125: weights of the file nodes for the nodes tetrahedra (10-14) and wedges (18-20) lack
_WT = np.zeros((5, 10))
for _f in range(4):
    _WT[_f, fe_tet.FACE_NODES[_f, :3]] = -1. / 9.
    _WT[_f, fe_tet.FACE_NODES[_f, 3:6]] = 4. / 9.
_WT[4, :4], _WT[4, 4:] = -1. / 8., 1. / 4.
_WW = np.zeros((3, 18))
_WW[0, [0, 1, 2]], _WW[0, [6, 7, 8]] = -1. / 9., 4. / 9.
_WW[1, [3, 4, 5]], _WW[1, [9, 10, 11]] = -1. / 9., 4. / 9.
_WW[2, [12, 13, 14]], _WW[2, [15, 16, 17]] = -1. / 9., 4. / 9.
MISSING = {HEX: (27, None), TET: (10, _WT), WEDGE: (18, _WW)}


class Level:
    pass


def _renumber(L, conn, etype, part, nprocs, material=None, group=None):
    """Mesh.cpp:517-559 (+ :589-616, :621-702): stable element reorder by rank, then the reference's bubble sort
    by (material, group, index) inside each rank; nodes by first visit over (rank, family, element, local node
    of that family)."""
    order = np.argsort(part, kind="stable")
    nel = conn.shape[0]
    elem_offset = np.concatenate([[0], np.cumsum(np.bincount(part, minlength=nprocs))])
    if material is not None:
        inv = list(order)                                   # inverse_element_mapping
        for p in range(nprocs):
            n = elem_offset[p + 1] - elem_offset[p]
            while n > 1:
                new_n = 0
                for j in range(elem_offset[p] + 1, elem_offset[p] + n):
                    jel, iel = inv[j], inv[j - 1]
                    if material[jel] < material[iel] or (material[jel] == material[iel] and
                                                         (group[jel] < group[iel] or (group[jel] == group[iel] and jel < iel))):
                        inv[j - 1], inv[j] = jel, iel
                        new_n = j - elem_offset[p]
                n = new_n
        order = np.array(inv, dtype=np.int64)
        L.material, L.group = material[order], group[order]
    conn, etype, part = conn[order], etype[order], part[order]
    new = {}
    own = np.zeros((3, nprocs), dtype=np.int64)
    for p in range(nprocs):
        for k in range(3):
            for e in range(elem_offset[p], elem_offset[p + 1]):
                lo = 0 if k == 0 else NVE[etype[e]][k - 1]
                for i in range(lo, NVE[etype[e]][k]):
                    ii = int(conn[e, i])
                    if ii not in new:
                        new[ii] = len(new)
                        own[k:, p] += 1
    L.order_el, L.etype, L.part, L.nprocs, L.nel = order, etype, part, nprocs, nel
    L.conn = np.full_like(conn, -1)
    for e in range(nel):
        n = NVE[etype[e]][2]
        L.conn[e, :n] = [new[int(v)] for v in conn[e, :n]]
    L.old_of_new = np.array(sorted(new, key=new.get), dtype=np.int64)
    L.nnode = len(new)
    L.elem_offset = elem_offset
    L.dof_offset = np.zeros((3, nprocs + 1), dtype=np.int64)
    L.dof_offset[:, 1:] = np.cumsum(own, axis=1)
    return L


def _add_biquadratic_nodes(conn, etype, xyz):
    nel = conn.shape[0]
    nn = xyz.shape[1]
    for iel in range(nel):
        t = etype[iel]
        if t == HEX:
            continue
        for iface in range(NFACES[t]):
            if FACE_NVERT[t][iface] != 3:
                continue
            inode = NVE[t][1] + iface
            if conn[iel, inode] >= 0:
                continue
            conn[iel, inode] = nn
            mine = set(conn[iel, FACE_NODES[t][iface][:3]].tolist())
            found = False
            for jel in range(iel + 1, nel):
                tj = etype[jel]
                if tj == HEX:
                    continue
                for jface in range(NFACES[tj]):
                    if FACE_NVERT[tj][jface] == 3 and conn[jel, NVE[tj][1] + jface] < 0 and \
                            set(conn[jel, FACE_NODES[tj][jface][:3]].tolist()) == mine:
                        conn[jel, NVE[tj][1] + jface] = nn
                        found = True
                        break
                if found:
                    break
            nn += 1
    for iel in range(nel):
        if etype[iel] != HEX:
            conn[iel, NVE[etype[iel]][2] - 1] = nn
            nn += 1
    out = np.zeros((3, nn))
    out[:, :xyz.shape[1]] = xyz
    for iel in range(nel):
        t = etype[iel]
        jstart, W = MISSING[t]
        for j in range(jstart, NVE[t][2]):
            acc = np.zeros(3)
            for i in range(jstart):
                acc = acc + out[:, conn[iel, i]] * W[j - jstart, i]
            out[:, conn[iel, j]] = acc
    return conn, out


def read_neu(path, Lref=1.0):
    lines = open(path).read().split("\n")

    def section(title):
        i = next(k for k, l in enumerate(lines) if l.strip().startswith(title))
        j = next(k for k in range(i, len(lines)) if lines[k].strip() == "ENDOFSECTION")
        return lines[i + 1:j]

    hdr = next(k for k, l in enumerate(lines) if "NUMNP" in l)
    nvt, nel, ngroup, nbcd, dim, dimn = [int(t) for t in lines[hdr + 1].split()]
    assert dim == 3 and dimn == 3 and ngroup >= 1
    xyz = np.array([[float(t) for t in l.split()[1:4]] for l in section("NODAL COORDINATES")]).T / Lref
    toks = " ".join(section("ELEMENTS/CELLS")).split()
    # ELEMENT GROUP sections (GambitIO.cpp:290-313): header, integer group name, flags line, element ids
    material, group = np.zeros(nel, dtype=np.int64), np.ones(nel, dtype=np.int64)
    gstarts = [k for k, l in enumerate(lines) if l.strip().startswith("GROUP:")]
    assert len(gstarts) == ngroup
    for i in gstarts:
        head = lines[i].split()
        ngel, mat = int(head[head.index("ELEMENTS:") + 1]), int(head[head.index("MATERIAL:") + 1])
        name = int(lines[i + 1].split()[0])
        j = next(k for k in range(i, len(lines)) if lines[k].strip() == "ENDOFSECTION")
        ids = np.array(" ".join(lines[i + 3:j]).split(), dtype=np.int64) - 1
        assert ids.shape[0] == ngel
        group[ids], material[ids] = name, mat
    conn = np.full((nel, 27), -1, dtype=np.int64)
    etype = np.zeros(nel, dtype=np.int64)
    p = 0
    for e in range(nel):
        nve = int(toks[p + 2])
        t = FILE_NVE[nve]
        etype[e] = t
        conn[e, GAMBIT_TO_FEMUS_VERTEX[t]] = np.array(toks[p + 3:p + 3 + nve], dtype=np.int64) - 1
        p += 3 + nve
    face = np.full((nel, 6), -1, dtype=np.int64)
    starts = [k for k, l in enumerate(lines) if l.strip().startswith("BOUNDARY CONDITIONS")]
    assert len(starts) == nbcd
    for i in starts:
        head = lines[i + 1].split()
        value, nface = int(head[0]), int(head[2])
        for l in lines[i + 2:i + 2 + nface]:
            iel, _, iface = [int(t) for t in l.split()]
            face[iel - 1, GAMBIT_TO_FEMUS_FACE[etype[iel - 1]][iface - 1]] = -value - 1
    conn, xyz_file = _add_biquadratic_nodes(conn, etype, xyz)
    L = _renumber(Level(), conn, etype, np.zeros(nel, dtype=np.int64), 1, material, group)
    L.face = face[L.order_el]
    L.xyz = xyz_file[:, L.old_of_new]
    L.level = 0
    return L


def refine(C):
    nelc = C.nel
    conn = np.full((nelc * 8, 27), -1, dtype=np.int64)
    face = np.full((nelc * 8, 6), -1, dtype=np.int64)
    etype = np.repeat(C.etype, 8)
    for iel in range(nelc):
        t = C.etype[iel]
        nv = NVE[t][0]
        for j in range(8):
            conn[iel * 8 + j, :nv] = C.conn[iel, CHILD_VERTICES[t][j][:nv]]
        for f in range(NFACES[t]):
            if C.face[iel, f] < -1:
                for (j, jf) in COARSE_TO_FINE_FACE[t][f]:
                    face[iel * 8 + j, jf] = C.face[iel, f]
    nn = C.nnode
    edges, faces = {}, {}
    for e in range(nelc * 8):
        t = etype[e]
        for k, (a, b) in enumerate(EDGES[t]):
            key = frozenset((int(conn[e, a]), int(conn[e, b])))
            if key not in edges:
                edges[key] = nn
                nn += 1
            conn[e, NVE[t][0] + k] = edges[key]
    for e in range(nelc * 8):
        t = etype[e]
        for f in range(NFACES[t]):
            key = frozenset(int(conn[e, v]) for v in FACE_NODES[t][f][:FACE_NVERT[t][f]])
            if key not in faces:
                faces[key] = nn
                nn += 1
            conn[e, NVE[t][1] + f] = faces[key]
    for e in range(nelc * 8):
        conn[e, NVE[etype[e]][2] - 1] = nn
        nn += 1
    F = _renumber(Level(), conn, etype, np.repeat(C.part, 8), C.nprocs, np.repeat(C.material, 8), np.repeat(C.group, 8))
    F.face = face[F.order_el]
    F.level = C.level + 1
    inv = np.empty(nelc * 8, dtype=np.int64)
    inv[F.order_el] = np.arange(nelc * 8)
    F.child_el = inv.reshape(nelc, 8)
    P = prolongator(C, F, "biquadratic")
    F.xyz = np.stack([P @ C.xyz[d] for d in range(3)])
    return F


def build_hierarchy(path, nlevels, Lref=1.0):
    lv = [read_neu(path, Lref)]
    for _ in range(1, nlevels):
        lv.append(refine(lv[-1]))
    return lv


def node_dof(L, order, nodes):
    k = FAMILY[order]
    nodes = np.asarray(nodes)
    if k == 2:
        return nodes.copy()
    p = np.searchsorted(L.dof_offset[2], nodes, side="right") - 1
    return (nodes - L.dof_offset[2][p]) + L.dof_offset[k][p]


def element_dofs(L, order):
    """list over elements of the dofs of family `order` (ragged)."""
    k = FAMILY[order]
    return [node_dof(L, order, L.conn[e, :NVE[L.etype[e]][k]]) for e in range(L.nel)]


def system_dofs27(L, order):
    out = np.full((L.nel, 27), -1, dtype=np.int64)
    for e, d in enumerate(element_dofs(L, order)):
        out[e, :len(d)] = d
    return out


def ndofs(L, order):
    return int(L.dof_offset[FAMILY[order]][-1])


def bdc_flags(L, order, dirichlet_faces=(1, 2, 3, 4, 5, 6)):
    bdc = np.full(ndofs(L, order), 2.0)
    for e in range(L.nel):
        t = L.etype[e]
        for f in range(NFACES[t]):
            bidx = -(L.face[e, f] + 1)
            if bidx > 0 and bidx in dirichlet_faces:
                nfd = FE[t].FACE_NDOFS[order]
                nfd = nfd[f] if isinstance(nfd, list) else nfd
                bdc[node_dof(L, order, L.conn[e, FACE_NODES[t][f][:nfd]])] = 0.0
    return bdc


def sparsity(L, order):
    n = ndofs(L, order)
    rows, cols = [], []
    for d in element_dofs(L, order):
        rows.append(np.repeat(d, len(d)))
        cols.append(np.tile(d, len(d)))
    rows, cols = np.concatenate(rows), np.concatenate(cols)
    A = sp.csr_matrix((np.ones(rows.shape[0], dtype=np.int8), (rows, cols)), shape=(n, n))
    A.sum_duplicates()
    A.sort_indices()
    return A.indptr.astype(np.int64), A.indices.astype(np.int32)


def assemble(L, order, U=None, fsrc=1.0):
    n = ndofs(L, order)
    if U is None:
        U = np.zeros(n)
    dofs = element_dofs(L, order)
    rows, cols, vals = [], [], []
    rhs = np.zeros(n)
    for t in (HEX, TET, WEDGE):
        sel = [e for e in range(L.nel) if L.etype[e] == t]
        if not sel:
            continue
        nve = NVE[t][FAMILY[order]]
        d = np.array([dofs[e] for e in sel])
        X = L.xyz[:, L.conn[sel][:, :nve]].transpose(1, 0, 2)
        F, B = fe_hex.poisson_elements(order, X, U[d], fsrc, FE[t].tables(order))
        rows.append(np.repeat(d, nve, axis=1).ravel())
        cols.append(np.tile(d, (1, nve)).ravel())
        vals.append(B.ravel())
        np.add.at(rhs, d.ravel(), F.ravel())
    A = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n))
    A.sum_duplicates()
    A.sort_indices()
    return A, rhs


def neumann_rhs(L, order, neumann):
    """Boundary part of the residual on a mesh of any element types: for every boundary face whose index is a key
    of `neumann` (value = constant flux), F[local node] += sum_g phi_i value weight_g with the face element
    _finiteElement[GetElementFaceType][order] -- triangle or quadrilateral -- on the face's first
    GetElementFaceDofNumber nodes (applications/001_Poisson/main.cpp:495-548; face -> local nodes: Elem.hpp `ig`)."""
    from . import fe_face
    rhs = np.zeros(ndofs(L, order))
    tabs = {k: fe_face.tables(k, order) for k in ("tri", "quad")}
    for e in range(L.nel):
        t = L.etype[e]
        for f in range(NFACES[t]):
            b = -(int(L.face[e, f]) + 1)
            if b <= 0 or b not in neumann:
                continue
            kind = fe_face.face_kind(FACE_NVERT[t][f])
            loc = FACE_NODES[t][f][:fe_face.ndofs(kind, order)]
            Ff = fe_face.neumann_face(L.xyz[:, L.conn[e, loc]], float(neumann[b]), tabs[kind])
            np.add.at(rhs, node_dof(L, order, L.conn[e, loc]), Ff)
    return rhs


def _local_prolongator(t, order):
    if t != HEX:
        return FE[t].local_prolongator(order)
    n = fe_hex.NDOFS[order]
    P = np.zeros((8, n, n))
    for j in range(8):
        phi, _ = fe_hex.shape(order, (fe_hex.XC[j][None, :] + fe_hex.XC[:n]) / 2.0)
        phi[np.abs(phi) < 1.0e-14] = 0.0
        P[j] = phi
    return P


def prolongator(C, F, order):
    k = FAMILY[order]
    Ploc = {t: _local_prolongator(t, order) for t in set(C.etype.tolist())}
    dc, df = element_dofs(C, order), element_dofs(F, order)
    rows, cols, vals = [], [], []
    done = np.zeros(ndofs(F, order), dtype=bool)
    for E in range(C.nel):
        t = C.etype[E]
        for j in range(8):
            fe = int(F.child_el[E, j])
            for a in range(NVE[t][k]):
                r = df[fe][a]
                if done[r]:
                    continue
                done[r] = True
                nz = np.nonzero(Ploc[t][j, a])[0]
                rows += [r] * len(nz)
                cols += dc[E][nz].tolist()
                vals += Ploc[t][j, a, nz].tolist()
    P = sp.csr_matrix((vals, (rows, cols)), shape=(ndofs(F, order), ndofs(C, order)))
    P.sort_indices()
    return P


zero_dirichlet = mb.zero_dirichlet
