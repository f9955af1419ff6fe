"""Oracle (TEST INFRASTRUCTURE ONLY): tetrahedral meshes read from Gambit neutral files -- FEMuS node/dof
numbering, the face and centre nodes the file lacks, uniform 1 -> 8 refinement, Dirichlet flags, sparsity,
Poisson assembly and prolongators, restated with numpy/scipy and plain loops (small meshes).

PARITY: equal to oracle/mesh_mixed.py on tetrahedral files (tests/test_host_mesh.py), which tests/test_reference_pin.py pins
to the reference's own output (tests/golden/ref_poisson_cube_tet_*.npz); FE arithmetic pinned to the compiled reference.
Restates (paths relative to /root/reference/src):
  06_mesh/00_single_level/01_input/01_from_external_file/GambitIO.cpp:56-85, 92-352   file sections, permutations
  06_mesh/00_single_level/00_definition/Mesh.cpp:105-125, 1207-1333   AddBiquadraticNodesNotInMeshFile
  06_mesh/00_single_level/00_definition/Mesh.cpp:517-559              renumbering by first visit (mesh_box._renumber)
  06_mesh/00_single_level/03_refinement/MeshRefinement.cpp:188-507, 513-621; MeshRefinement.hpp:88-93, 124-127
  06_solution/01_multiple_levels/00_definition/MultiLevelSolution.cpp:725-840        GenerateBdc
  08_equations/00_stationary/LinearImplicitSystem.cpp:761-909, 1032-1120             prolongators
The loops follow the reference's own (quadratic-time face matching included), NOT the hash-map scheme of
the product's host layer, so that the two implementations are independent.
"""
import numpy as np
import scipy.sparse as sp

from . import fe_hex, fe_tet, mesh_box as mb

NVE = (4, 10, 15)
FAMILY = mb.FAMILY
GAMBIT_TO_FEMUS_VERTEX = np.array([0, 4, 1, 6, 5, 2, 7, 8, 9, 3])
GAMBIT_TO_FEMUS_FACE = np.array([0, 1, 2, 3])
# MeshRefinement.hpp:88-93: parent face -> its 4 (child, child face) pairs
COARSE_TO_FINE_FACE = [[(0, 0), (1, 0), (2, 0), (4, 0)], [(0, 1), (1, 1), (3, 1), (5, 1)],
                       [(1, 2), (2, 2), (3, 2), (6, 2)], [(2, 3), (0, 3), (3, 3), (7, 3)]]
# Mesh.cpp:105-113: weights of the 10 file nodes for the 4 face nodes and the centre
_W = np.zeros((5, 10))
for _f in range(4):
    _W[_f, fe_tet.FACE_NODES[_f, :3]] = -1. / 9.
    _W[_f, fe_tet.FACE_NODES[_f, 3:6]] = 4. / 9.
_W[4, :4] = -1. / 8.
_W[4, 4:] = 1. / 4.


def _add_biquadratic_nodes(conn10, xyz):
    """conn10[nel,10], xyz[3,nvt] (file numbering) -> conn15, xyz with the new nodes appended."""
    nel = conn10.shape[0]
    conn = np.full((nel, 15), -1, dtype=np.int64)
    conn[:, :10] = conn10
    nn = xyz.shape[1]
    for iel in range(nel):
        for iface in range(4):
            inode = 10 + iface
            if conn[iel, inode] >= 0:
                continue
            conn[iel, inode] = nn
            mine = set(conn[iel, fe_tet.FACE_NODES[iface, :3]].tolist())
            found = False
            for jel in range(iel + 1, nel):
                for jface in range(4):
                    if conn[jel, 10 + jface] < 0 and set(conn[jel, fe_tet.FACE_NODES[jface, :3]].tolist()) == mine:
                        conn[jel, 10 + jface] = nn
                        found = True
                        break
                if found:
                    break
            nn += 1
    for iel in range(nel):
        conn[iel, 14] = nn
        nn += 1
    out = np.zeros((3, nn))
    out[:, :xyz.shape[1]] = xyz
    for iel in range(nel):
        for j in range(10, 15):
            acc = np.zeros(3)
            for i in range(10):
                acc = acc + out[:, conn[iel, i]] * _W[j - 10, i]
            out[:, conn[iel, j]] = acc
    return conn, out


def read_tet10(path, Lref=1.0):
    """Level 0 read from a .neu file of 10-node tetrahedra (one element group)."""
    lines = open(path).read().split("\n")

    def section(title):
        i = next(k for k, l in enumerate(lines) if l.strip().startswith(title))
        j = next(k for k in range(i, len(lines)) if lines[k].strip() == "ENDOFSECTION")
        return lines[i + 1:j]

    hdr = next(k for k, l in enumerate(lines) if "NUMNP" in l)
    nvt, nel, ngroup, nbcd, dim, dimn = [int(t) for t in lines[hdr + 1].split()]
    assert dim == 3 and dimn == 3 and ngroup == 1
    xyz = np.array([[float(t) for t in l.split()[1:4]] for l in section("NODAL COORDINATES")]).T / Lref
    toks = " ".join(section("ELEMENTS/CELLS")).split()
    conn10 = np.zeros((nel, 10), dtype=np.int64)
    p = 0
    for e in range(nel):
        assert int(toks[p + 2]) == 10
        conn10[e, GAMBIT_TO_FEMUS_VERTEX] = np.array(toks[p + 3:p + 13], dtype=np.int64) - 1
        p += 13
    face = np.full((nel, 4), -1, dtype=np.int64)
    starts = [k for k, l in enumerate(lines) if l.strip().startswith("BOUNDARY CONDITIONS")]
    assert len(starts) == nbcd
    for i in starts:
        head = lines[i + 1].split()
        value, nface = int(head[0]), int(head[2])
        for l in lines[i + 2:i + 2 + nface]:
            iel, _, iface = [int(t) for t in l.split()]
            face[iel - 1, GAMBIT_TO_FEMUS_FACE[iface - 1]] = -value - 1
    conn_file, xyz_file = _add_biquadratic_nodes(conn10, xyz)
    L = mb.Level()
    mb._finish_level(L, conn_file, np.zeros(nel, dtype=np.int64), 1, NVE)
    L.face = face[L.order_el]
    L.xyz = xyz_file[:, L.lat_of_node]
    L.level = 0
    return L


def refine(C):
    """MeshRefinement::RefineMesh for tetrahedra: children 8*iel+j, vertices through CHILD_VERTICES,
    mid-edge / face / centre nodes shared by vertex sets, boundary flags through COARSE_TO_FINE_FACE."""
    nelc = C.nel
    conn = np.full((nelc * 8, 15), -1, dtype=np.int64)
    face = np.full((nelc * 8, 4), -1, dtype=np.int64)
    for iel in range(nelc):
        for j in range(8):
            conn[iel * 8 + j, :4] = C.conn[iel, fe_tet.CHILD_VERTICES[j]]
        for f in range(4):
            if C.face[iel, f] < -1:
                for (j, jf) in COARSE_TO_FINE_FACE[f]:
                    face[iel * 8 + j, jf] = C.face[iel, f]
    nn = C.nnode
    edges, faces = {}, {}
    for e in range(nelc * 8):                       # mid-edge nodes first, for all elements (MeshRefinement.cpp:365-417)
        for k, (a, b) in enumerate(fe_tet.EDGES):
            key = frozenset((int(conn[e, a]), int(conn[e, b])))
            if key not in edges:
                edges[key] = nn
                nn += 1
            conn[e, 4 + k] = edges[key]
    for e in range(nelc * 8):                       # then the face nodes, then the centres (AddFaceDofAndElementDof)
        for k, f in enumerate(fe_tet.FACES):
            key = frozenset(int(conn[e, v]) for v in f)
            if key not in faces:
                faces[key] = nn
                nn += 1
            conn[e, 10 + k] = faces[key]
    for e in range(nelc * 8):
        conn[e, 14] = nn
        nn += 1
    F = mb.Level()
    mb._finish_level(F, conn, np.repeat(C.part, 8), C.nprocs, NVE)
    F.face = face[F.order_el]
    F.level = C.level + 1
    inv = np.empty(nelc * 8, dtype=np.int64)
    inv[F.order_el] = np.arange(nelc * 8)
    F.child_el = inv.reshape(nelc, 8)
    P = prolongator(C, F, "biquadratic")
    F.xyz = np.stack([P @ C.xyz[d] for d in range(3)])
    return F


def build_hierarchy(path, nlevels, Lref=1.0):
    lv = [read_tet10(path, Lref)]
    for _ in range(1, nlevels):
        lv.append(refine(lv[-1]))
    return lv


def solution_dof(L, order):
    k = FAMILY[order]
    nodes = L.conn[:, :NVE[k]]
    if k == 2:
        return nodes.copy()
    p = mb.node_owner(L, nodes)
    return (nodes - L.dof_offset[2][p]) + L.dof_offset[k][p]


system_dof = solution_dof


def ndofs(L, order):
    return int(L.dof_offset[FAMILY[order]][-1])


def bdc_flags(L, order, dirichlet_faces=(1, 2, 3, 4, 5, 6)):
    bdc = np.full(ndofs(L, order), 2.0)
    dofs = solution_dof(L, order)
    nfd = fe_tet.FACE_NDOFS[order]
    for f in range(4):
        bidx = -(L.face[:, f] + 1)
        sel = np.isin(bidx, dirichlet_faces) & (bidx > 0)
        bdc[dofs[sel][:, fe_tet.FACE_NODES[f, :nfd]].ravel()] = 0.0
    return bdc


def sparsity(L, order):
    d = system_dof(L, order)
    n = ndofs(L, order)
    nve = d.shape[1]
    A = sp.csr_matrix((np.ones(d.size * nve, dtype=np.int8), (np.repeat(d, nve, axis=1).ravel(), np.tile(d, (1, nve)).ravel())),
                      shape=(n, n))
    A.sum_duplicates()
    A.sort_indices()
    return A.indptr.astype(np.int64), A.indices.astype(np.int32)


def assemble(L, order, U=None, fsrc=1.0):
    """Poisson matrix (CSR, full pattern) and residual on level L (applications/001_Poisson/main.cpp:346-605;
    the geometry map uses the unknown's own nodes, ElemType.hpp:1462)."""
    d = system_dof(L, order)
    n = ndofs(L, order)
    nve = d.shape[1]
    if U is None:
        U = np.zeros(n)
    X = L.xyz[:, L.conn[:, :nve]].transpose(1, 0, 2)
    F, B = fe_hex.poisson_elements(order, X, U[d], fsrc, fe_tet.tables(order))
    A = sp.csr_matrix((B.ravel(), (np.repeat(d, nve, axis=1).ravel(), np.tile(d, (1, nve)).ravel())), shape=(n, n))
    A.sum_duplicates()
    A.sort_indices()
    rhs = np.zeros(n)
    np.add.at(rhs, d.ravel(), F.ravel())
    return A, rhs


def prolongator(C, F, order):
    """P from C to its refinement F (BuildProlongatorMatrix): row of the fine dof at (child j, node a) =
    coarse functions there; rows are inserted, the first visit defines them."""
    nve = NVE[FAMILY[order]]
    Ploc = fe_tet.local_prolongator(order)
    dc, df = solution_dof(C, order), solution_dof(F, order)
    rows, cols, vals = [], [], []
    done = np.zeros(ndofs(F, order), dtype=bool)
    for E in range(C.nel):
        for j in range(8):
            fe = C_child(C, F, E, j)
            for a in range(nve):
                r = df[fe, a]
                if done[r]:
                    continue
                done[r] = True
                nz = np.nonzero(Ploc[j, a])[0]
                rows += [r] * len(nz)
                cols += dc[E, nz].tolist()
                vals += Ploc[j, a, nz].tolist()
    P = sp.csr_matrix((vals, (rows, cols)), shape=(ndofs(F, order), ndofs(C, order)))
    P.sort_indices()
    return P


zero_dirichlet = mb.zero_dirichlet


def C_child(C, F, E, j):
    return int(F.child_el[E, j])
