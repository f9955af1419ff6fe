"""Oracle (TEST INFRASTRUCTURE ONLY): structured HEX27 box meshes, FEMuS node/dof numbering,
uniform refinement, Dirichlet flags, system sparsity and prolongators, restated with numpy.

PARITY PINNED TO REFERENCE OUTPUT (round 2): the reference's own sources, compiled unmodified on the single-process host
backend of oracle/ref_build and run here (applications/001_Poisson/main.cpp), produced tests/golden/ref_poisson_*.npz
(tests/golden/make_ref_golden.py); tests/test_reference_pin.py compares this module with them -- integers bit-exact,
values and printed residual norms to the stated tolerances.  Each function cites what it restates
(paths relative to /root/reference/src):

  06_mesh/00_single_level/01_input/02_from_implemented_code/MeshGeneration.cpp:790-849   box nodes
  .../MeshGeneration.cpp:976-1071, MeshGeneration.hpp:113-139     element connectivity, face flags
  06_mesh/00_single_level/00_definition/Mesh.cpp:517-559          node renumbering (first visit)
  .../Mesh.cpp:589-616, 621-702                                   element reorder by rank / material
  .../Mesh.cpp:706-853, 1021-1074                                 dof offsets, GetSolutionDof
  06_mesh/00_single_level/03_refinement/MeshRefinement.cpp:188-507  uniform 1->8 refinement
  06_mesh/00_single_level/02_partitioning/MeshMetisPartitioning.cpp:143-155  children inherit rank
  06_solution/01_multiple_levels/00_definition/MultiLevelSolution.cpp:725-840  GenerateBdc
  08_algebra.../LinearEquation.cpp:76-85, 213-237, 407-548        system dofs, sparsity
  08_equations/00_stationary/LinearImplicitSystem.cpp:761-909, 1032-1120  prolongator, Dirichlet zeroing

Unlike the product (which refines topologically, like the reference), this oracle exploits that a
box mesh lives on an integer lattice: level l has (2*nx_l+1)(2*ny_l+1)(2*nz_l+1) lattice nodes and
every entity is identified by its lattice coordinates.  Only the ELEMENT ORDER and the shared-node
TOPOLOGY matter for the final numbering (the reference renumbers after every refinement).
"""
import numpy as np
import scipy.sparse as sp

from . import fe_hex

FAMILY = {"linear": 0, "quadratic": 1, "biquadratic": 2}
NVE = (8, 20, 27)
# BuildBox: face -> faceElementIndex written by the generator (MeshGeneration.cpp:1038-1071);
# boundary index = -(value+1) (Elem.cpp:361-364): front(y-)=2 right(x+)=3 behind(y+)=4 left(x-)=5
# bottom(z-)=1 top(z+)=6.
FACE_BOUNDARY_INDEX = {0: 2, 1: 3, 2: 4, 3: 5, 4: 1, 5: 6}


class Level:
    pass


def _renumber(conn_lat, part_el, nprocs, NVE=NVE):
    """Element reorder by rank (stable) + node renumbering by first visit over
    (rank, family k, element, local node in [NVE[k-1],NVE[k])) -- Mesh.cpp:517-559, 589-616.
    NVE: dofs per family of the (single) element type, hexahedra by default."""
    order_el = np.argsort(part_el, kind="stable")
    conn_lat = conn_lat[order_el]
    part_el = part_el[order_el]
    elem_offset = np.concatenate([[0], np.cumsum(np.bincount(part_el, minlength=nprocs))])
    seqs, tags = [], []
    for p in range(nprocs):
        c = conn_lat[elem_offset[p]:elem_offset[p + 1]]
        lo = 0
        for k in range(3):
            s = c[:, lo:NVE[k]].ravel()
            seqs.append(s)
            tags.append(np.full(s.shape[0], p * 3 + k, dtype=np.int64))
            lo = NVE[k]
    seq = np.concatenate(seqs)
    tag = np.concatenate(tags)
    uniq, first = np.unique(seq, return_index=True)
    visit = np.argsort(first, kind="stable")              # lattice ids in first-visit order
    new_of_lat = {}
    lat_of_new = uniq[visit]
    # own sizes: a node first met at (p,k) counts for families j>=k on rank p
    t = tag[first[visit]]
    own = np.zeros((3, nprocs), dtype=np.int64)
    for p in range(nprocs):
        for k in range(3):
            cnt = int(np.count_nonzero(t == p * 3 + k))
            own[k:, p] += cnt
    dof_offset = np.zeros((3, nprocs + 1), dtype=np.int64)
    dof_offset[:, 1:] = np.cumsum(own, axis=1)
    # map lattice id -> new id
    sorter = np.argsort(lat_of_new)
    pos = np.searchsorted(lat_of_new, conn_lat.ravel(), sorter=sorter)
    conn = sorter[pos].reshape(conn_lat.shape)
    return order_el, conn, lat_of_new, elem_offset, dof_offset, own, part_el


def _finish_level(L, conn_lat, part_el, nprocs, NVE=NVE):
    order_el, conn, lat_of_new, elem_offset, dof_offset, own, part_el = _renumber(conn_lat, part_el, nprocs, NVE)
    L.order_el = order_el                 # position -> pre-reorder element index
    L.conn = conn                         # [nel,27] node ids (FEMuS numbering)
    L.lat_of_node = lat_of_new            # node id -> lattice id
    L.elem_offset = elem_offset
    L.dof_offset = dof_offset             # [3][nprocs+1]
    L.own_size = own
    L.part = part_el
    L.nprocs = nprocs
    L.nel = conn.shape[0]
    L.nnode = lat_of_new.shape[0]
    return L


def build_box(nx, ny, nz, bounds=(0., 1., 0., 1., 0., 1.), partition=None, nprocs=1):
    """Level-0 HEX27 box (MeshGeneration.cpp:790-849, 976-1071)."""
    L = Level()
    L.n = (nx, ny, nz)
    L.bounds = bounds
    sx, sy = 2 * nx + 1, 2 * ny + 1
    k, j, i = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    i, j, k = 2 * i.ravel(), 2 * j.ravel(), 2 * k.ravel()       # element order: k, j, i (i fastest)
    off = fe_hex.IND                                            # local node -> lattice offset
    conn_lat = ((i[:, None] + off[None, :, 0]) + sx * ((j[:, None] + off[None, :, 1]) +
                                                      sy * (k[:, None] + off[None, :, 2])))
    nel = nx * ny * nz
    face = np.full((nel, 6), -1, dtype=np.int64)               # faceElementIndex; -1 = unset/interior
    face[k == 0, 4] = -2
    face[k == 2 * (nz - 1), 5] = -7
    face[j == 0, 0] = -3
    face[j == 2 * (ny - 1), 2] = -5
    face[i == 0, 3] = -6
    face[i == 2 * (nx - 1), 1] = -4
    part = np.zeros(nel, dtype=np.int64) if partition is None else np.asarray(partition, dtype=np.int64)
    _finish_level(L, conn_lat, part, nprocs)
    L.face = face[L.order_el]
    # coordinates (MeshGeneration.cpp:842-844), permuted to the new numbering
    lat = L.lat_of_node
    li, lj, lk = lat % sx, (lat // sx) % sy, lat // (sx * sy)
    xmin, xmax, ymin, ymax, zmin, zmax = bounds
    L.xyz = np.stack([(li.astype(np.float64) / float(2 * nx)) * (xmax - xmin) + xmin,
                      (lj.astype(np.float64) / float(2 * ny)) * (ymax - ymin) + ymin,
                      (lk.astype(np.float64) / float(2 * nz)) * (zmax - zmin) + zmin])
    L.level = 0
    L.parent = None
    return L


def slab_partition(nx, ny, nz, nprocs):
    """z-slab partition vector for the level-0 box (elements in k,j,i order)."""
    k = np.repeat(np.arange(nz), nx * ny)
    return (k * nprocs) // nz


def refine(C):
    """Uniform refinement of level C (MeshRefinement.cpp:188-507): children 8*iel+j in coarse
    element order, child j = octant at parent vertex j, children inherit the parent's rank."""
    F = Level()
    nx, ny, nz = C.n
    F.n = (2 * nx, 2 * ny, 2 * nz)
    F.bounds = C.bounds
    sxc, syc = 2 * nx + 1, 2 * ny + 1
    sxf, syf = 4 * nx + 1, 4 * ny + 1
    # lattice origin (vertex 0) of each coarse element on the FINE lattice
    lat0 = C.lat_of_node[C.conn[:, 0]]
    oi, oj, ok = 2 * (lat0 % sxc), 2 * ((lat0 // sxc) % syc), 2 * (lat0 // (sxc * syc))
    cvm = fe_hex.child_vertex_map()
    nelc = C.nel
    conn_lat = np.zeros((nelc, 8, 27), dtype=np.int64)
    face = np.full((nelc, 8, 6), -1, dtype=np.int64)
    for jc in range(8):
        # child origin inside the parent's 5x5x5 block: position of child vertex 0
        o = 2 * fe_hex.IND[cvm[jc, 0]]
        ci, cj, ck = oi + o[0], oj + o[1], ok + o[2]
        conn_lat[:, jc, :] = ((ci[:, None] + fe_hex.IND[None, :, 0]) +
                              sxf * ((cj[:, None] + fe_hex.IND[None, :, 1]) +
                                     syf * (ck[:, None] + fe_hex.IND[None, :, 2])))
        for f in range(6):                    # child jc touches parent face f iff vertex jc is on it
            if jc in fe_hex.FACE_NODES[f, :4]:
                sel = C.face[:, f] < -1
                face[sel, jc, f] = C.face[sel, f]
    conn_lat = conn_lat.reshape(nelc * 8, 27)
    part = np.repeat(C.part, 8)
    _finish_level(F, conn_lat, part, C.nprocs)
    F.face = face.reshape(nelc * 8, 6)[F.order_el]
    F.level = C.level + 1
    F.parent = C
    # (coarse element, child) -> fine element index after reordering
    inv = np.empty(nelc * 8, dtype=np.int64)
    inv[F.order_el] = np.arange(nelc * 8)
    F.child_el = inv.reshape(nelc, 8)
    # coordinates = P_biquadratic * coarse coordinates (MeshRefinement.cpp:470-472)
    P = prolongator(C, F, "biquadratic")
    F.xyz = np.stack([P @ C.xyz[d] for d in range(3)])
    return F


def build_hierarchy(nx, ny, nz, nlevels, bounds=(0., 1., 0., 1., 0., 1.), nprocs=1, partition=None):
    if partition is None and nprocs > 1:
        partition = slab_partition(nx, ny, nz, nprocs)
    levels = [build_box(nx, ny, nz, bounds, partition, nprocs)]
    for _ in range(1, nlevels):
        levels.append(refine(levels[-1]))
    return levels


# ---------------------------------------------------------------------------------------------
def node_owner(L, nodes):
    return np.searchsorted(L.dof_offset[2], nodes, side="right") - 1


def solution_dof(L, order):
    """[nel, nve] dof of family `order` for every (element, local node) -- Mesh.cpp:1021-1074
    (uniform meshes: no 'owned ghost' nodes)."""
    k = FAMILY[order]
    nodes = L.conn[:, :NVE[k]]
    if k == 2:
        return nodes.copy()
    p = node_owner(L, nodes)
    return (nodes - L.dof_offset[2][p]) + L.dof_offset[k][p]


def ndofs(L, order):
    return int(L.dof_offset[FAMILY[order]][-1])


def system_dof(L, order):
    """Single-variable system: row = KKoffset[0][p] + dof - dofOffset[k][p] with
    KKoffset[0][p] = dofOffset[k][p]  (LinearEquation.cpp:76-85, 213-237) => identity."""
    return solution_dof(L, order)


def bdc_flags(L, order, dirichlet_faces=(1, 2, 3, 4, 5, 6)):
    """_Bdc per dof of family `order`: 2 free, 0 Dirichlet (MultiLevelSolution.cpp:725-840);
    `dirichlet_faces` = boundary indices (1..6) on which the BC function returns true."""
    bdc = np.full(ndofs(L, order), 2.0)
    dofs = solution_dof(L, order)
    nfd = fe_hex.FACE_NDOFS[order]
    for f in range(6):
        bidx = -(L.face[:, f] + 1)                 # Elem.cpp:361-364
        sel = np.isin(bidx, dirichlet_faces) & (bidx > 0)
        loc = fe_hex.FACE_NODES[f, :nfd]
        bdc[dofs[sel][:, loc].ravel()] = 0.0
    return bdc


def sparsity(L, order):
    """CSR pattern (rowptr int64, col int32, sorted) of the system matrix: every (i,j) pair of
    every element, zeros included (LinearEquation.cpp:407-548)."""
    d = system_dof(L, order)
    n = ndofs(L, order)
    nve = d.shape[1]
    rows = np.repeat(d, nve, axis=1).ravel()
    cols = np.tile(d, (1, nve)).ravel()
    A = sp.csr_matrix((np.ones(rows.shape[0], dtype=np.int8), (rows, cols)), shape=(n, n))
    A.sum_duplicates()
    A.sort_indices()
    return A.indptr.astype(np.int64), A.indices.astype(np.int32)


def assemble(L, order, U=None, fsrc=1.0):
    """Poisson matrix (CSR, full pattern) and residual vector on level L
    (applications/001_Poisson/main.cpp:346-605)."""
    d = system_dof(L, order)
    n = ndofs(L, order)
    nve = d.shape[1]
    if U is None:
        U = np.zeros(n)
    X = L.xyz[:, L.conn[:, :nve]].transpose(1, 0, 2)           # [nel,3,nve]: geometry from the unknown's nodes
    F, B = fe_hex.poisson_elements(order, X, U[d], fsrc)
    rows = np.repeat(d, nve, axis=1).ravel()
    cols = np.tile(d, (1, nve)).ravel()
    A = sp.csr_matrix((B.ravel(), (rows, cols)), shape=(n, n))
    A.sum_duplicates()
    A.sort_indices()
    rhs = np.zeros(n)
    np.add.at(rhs, d.ravel(), F.ravel())
    return A, rhs


def neumann_rhs(L, order, neumann):
    """Boundary part of the residual: for every element face on a boundary whose index is a key of
    `neumann` (value = constant flux), F[local node] += sum_g phi_i value weight with the face element of
    the unknown's family on the first nve_face face nodes (applications/001_Poisson/main.cpp:495-548;
    face -> local nodes: Elem.hpp `ig` table = fe_hex.FACE_NODES; boundary index = -(faceElementIndex+1),
    Elem.cpp:361-364)."""
    from . import fe_quad
    d = system_dof(L, order)
    n = ndofs(L, order)
    nvf = fe_quad.NDOFS2[order]
    tabs = fe_quad.tables2(order)
    rhs = np.zeros(n)
    for f in range(6):
        bidx = -(L.face[:, f] + 1)
        for e in np.nonzero(L.face[:, f] < -1)[0]:
            b = int(bidx[e])
            if b not in neumann:
                continue
            loc = fe_hex.FACE_NODES[f][:nvf]
            X = L.xyz[:, L.conn[e, loc]]
            Ff = fe_quad.neumann_face(order, X, float(neumann[b]), tabs)
            # the face's local nodes are element-local nodes < nve for both families (vertices first)
            np.add.at(rhs, d[e, loc], Ff)
    return rhs


def prolongator(C, F, order):
    """P (fine dofs x coarse dofs) of family `order` from level C to its refinement F
    (LinearImplicitSystem.cpp:761-909; rows inserted, identical from every neighbour)."""
    k = FAMILY[order]
    pts2 = fe_hex.fine_points(order)                          # doubled reference positions in [-2,2]
    Pl = fe_hex.local_prolongator(order, pts2)                # [nf, nve]
    nve = NVE[k]
    # fine dof ids of the nf fine points of every coarse element, through one (child, node) each
    child_of, node_of = [], []
    for pt in pts2:
        for jc in range(8):
            hit = np.nonzero(np.all(fe_hex.XC[jc] + fe_hex.XC[:nve] == pt, axis=1))[0]
            if hit.size:
                child_of.append(jc)
                node_of.append(int(hit[0]))
                break
    child_of, node_of = np.array(child_of), np.array(node_of)
    fdofs_all = solution_dof(F, order)                        # [nelf, nve]
    cdofs = solution_dof(C, order)                            # [nelc, nve]
    fel = F.child_el[:, child_of]                             # [nelc, nf]
    frow = fdofs_all[fel, node_of[None, :]]                   # [nelc, nf]
    li, lj = np.nonzero(Pl)
    rows = frow[:, li].ravel()
    cols = cdofs[:, lj].ravel()
    vals = np.tile(Pl[li, lj], C.nel)
    key = rows * ndofs(C, order) + cols
    _, first = np.unique(key, return_index=True)              # insert semantics: identical duplicates
    P = sp.csr_matrix((vals[first], (rows[first], cols[first])), shape=(ndofs(F, order), ndofs(C, order)))
    P.sort_indices()
    return P


def zero_dirichlet(P, bdc_f, bdc_c):
    """ZeroInterpolatorDirichletNodes (LinearImplicitSystem.cpp:1032-1120): rows of fine Dirichlet
    dofs and columns of coarse Dirichlet dofs set to zero; the pattern is kept."""
    P = P.tocoo()
    keep = (bdc_f[P.row] >= 1.5) & (bdc_c[P.col] >= 1.5)
    v = np.where(keep, P.data, 0.0)
    Q = sp.csr_matrix((v, (P.row, P.col)), shape=P.shape)
    Q.sort_indices()
    return Q
