"""CPU: the product's C++ host layer (topological refinement, femus_b200/host/BoxMesh.hpp) against
the independent lattice-based numpy oracle: integer results bit-exact, coordinates bit-exact."""
import os
import numpy as np
import pytest

from femus_b200 import hostapi
from oracle import mesh_box as mb, fe_hex


@pytest.mark.parametrize("shape,nl,nprocs", [((2, 3, 2), 3, 1), ((2, 2, 4), 3, 2), ((3, 2, 4), 2, 4), ((1, 1, 1), 3, 1),
                                             ((4, 4, 4), 1, 1), ((2, 2, 8), 2, 8)])
def test_hierarchy_bit_exact(shape, nl, nprocs):
    H = hostapi.HostHierarchy(*shape, nl, nprocs=nprocs)
    lv = mb.build_hierarchy(*shape, nl, nprocs=nprocs)
    for l in range(nl):
        a, b = H.levels[l], lv[l]
        assert np.array_equal(a.conn, b.conn)
        assert np.array_equal(a.face, b.face)
        assert np.array_equal(a.part, b.part)
        assert np.array_equal(a.elem_offset, b.elem_offset)
        assert np.array_equal(a.dof_offset, b.dof_offset)
        assert np.array_equal(a.xyz, b.xyz)
        if l < nl - 1:
            assert np.array_equal(a.child_el, lv[l + 1].child_el)
        for fam in ("linear", "quadratic", "biquadratic"):
            assert np.array_equal(a.system_dofs(fam), mb.system_dof(b, fam))
            assert np.array_equal(a.bdc(fam), mb.bdc_flags(b, fam))
            assert np.array_equal(a.bdc(fam, (3,)), mb.bdc_flags(b, fam, (3,)))
            if l > 0:
                rp, ci, v, shp = H.prolongator(l, fam)
                P = mb.prolongator(lv[l - 1], lv[l], fam)
                assert shp == P.shape
                assert np.array_equal(rp, P.indptr) and np.array_equal(ci, P.indices) and np.array_equal(v, P.data)


def test_non_unit_bounds_and_tables():
    H = hostapi.HostHierarchy(2, 2, 2, 2, bounds=(-1., 2., 0., 0.5, 3., 4.))
    lv = mb.build_hierarchy(2, 2, 2, 2, bounds=(-1., 2., 0., 0.5, 3., 4.))
    assert np.array_equal(H.levels[1].xyz, lv[1].xyz)
    for fam in ("linear", "quadratic", "biquadratic"):
        for x, y in zip(hostapi.hex_tables(fam), fe_hex.tables(fam)):
            assert np.array_equal(x, y)
        Pl = fe_hex.local_prolongator(fam)
        pts = fe_hex.fine_points(fam)
        for r, pt in enumerate(pts):
            idx, val = hostapi.hex_prolongator_row(fam, *(pt + 2))
            d = np.zeros(Pl.shape[1])
            d[idx] = val
            assert np.array_equal(d, Pl[r])


def test_bad_arguments():
    with pytest.raises(ValueError):
        hostapi.HostHierarchy(0, 1, 1, 1)


def test_face_element_tables_and_boundary_faces():
    """Host face element (Neumann integrals) bit-exact against the oracle restatement (itself pinned to the
    reference's elem_type_2D), face-node table, and the boundary-face list against the oracle's face flags."""
    from oracle import fe_quad, fe_hex
    for order in ("linear", "biquadratic"):
        for a, b in zip(hostapi.face_tables(order), fe_quad.tables2(order)):
            assert np.array_equal(a, b)
    assert np.array_equal(hostapi.hex_face_nodes(), fe_hex.FACE_NODES)
    H = hostapi.HostHierarchy(2, 3, 2, 2)
    lv = mb.build_hierarchy(2, 3, 2, 2)
    e, f, b = H.levels[-1].boundary_faces()
    ref = [(int(el), int(fa), int(-(lv[-1].face[el, fa] + 1))) for el in range(lv[-1].nel) for fa in range(6) if lv[-1].face[el, fa] < -1]
    assert list(zip(e.tolist(), f.tolist(), b.tolist())) == ref


NEU = os.path.join(os.path.dirname(__file__), "golden", "cube_hex27_2x2x2.neu")


def test_gambit_reader_matches_oracle_reader():
    """Host .neu reader (GambitIO.cpp:92-352) against the independent numpy reader: connectivity in FEMuS
    local order and first-visit numbering, boundary flags, coordinates, dof maps and Dirichlet flags,
    bit-exact; the refined levels stay conforming (node counts of a 2x2x2 cube: 5^3, 9^3, 17^3)."""
    from oracle import gambit
    H = hostapi.HostHierarchy.from_neu(NEU, 3)
    L = gambit.read_hex27(NEU)
    h = H.levels[0]
    assert np.array_equal(h.conn, L.conn) and np.array_equal(h.face, L.face) and np.array_equal(h.xyz, L.xyz)
    for order in ("linear", "biquadratic"):
        assert np.array_equal(h.system_dofs(order), mb.system_dof(L, order))
        assert np.array_equal(h.bdc(order), mb.bdc_flags(L, order))
    assert [lv.nnode for lv in H.levels] == [125, 729, 4913] and [lv.nel for lv in H.levels] == [8, 64, 512]
    for lv in H.levels:                      # six boundary sets with the same number of faces each
        b = -(lv.face[lv.face < -1] + 1)
        assert np.array_equal(np.bincount(b)[1:], np.full(6, lv.nel ** (2 / 3) + 0.5, dtype=int))
        assert lv.xyz.min() == 0.0 and lv.xyz.max() == 1.0


# ------------------------------------------------------------------------------ tetrahedra
NEU_TET = os.path.join(os.path.dirname(__file__), "golden", "cube_tet10.neu")
TET_ORDERS = ("linear", "quadratic", "biquadratic")


@pytest.mark.parametrize("geom", ["tet", "wedge"])
def test_simplex_element_tables_and_prolongator_rows(geom):
    """Host TetElement.hpp / WedgeElement.hpp against the committed values of the compiled reference
    (tests/golden/fe_{tet,wedge}_ref.npz): Gauss rule bit-exact, shape tables to a few ulp, element prolongator
    rows with the reference's non-zero structure."""
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", f"fe_{geom}_ref.npz"))
    et = {"tet": hostapi.TET, "wedge": hostapi.WEDGE}[geom]
    for order in TET_ORDERS:
        phi, dxi, deta, dzeta, w = hostapi.elem_tables(et, order)
        assert np.array_equal(w, G[f"{order}_gauss_w"])
        for a, k in ((phi, "phi"), (dxi, "dxi"), (deta, "deta"), (dzeta, "dzeta")):
            assert a.shape == G[f"{order}_{k}"].shape and np.abs(a - G[f"{order}_{k}"]).max() <= 3e-15, (order, k)
        Pg, kv = G[f"{order}_prol"], G[f"{order}_prol_kvert"]
        for i in range(Pg.shape[0]):
            idx, val = hostapi.elem_prolongator_row(et, order, int(kv[i, 0]), int(kv[i, 1]))
            assert np.array_equal(idx, np.nonzero(Pg[i])[0])
            assert np.abs(val - Pg[i, idx]).max() <= 4e-16


def test_child_faces_and_hex_through_the_general_entry_points():
    """The parent face of every child face, derived geometrically by the host, against coarse2FineFaceMapping
    (MeshRefinement.hpp:79-100) for the three element types; hexahedral tables / prolongator rows through the
    per-type entry points equal the specialised ones."""
    from oracle import mesh_mixed as mm
    for et in (hostapi.HEX, hostapi.TET, hostapi.WEDGE):
        nf = mm.NFACES[et]
        got = {(f, j, cf) for j in range(8) for cf in range(nf) for f in [hostapi.elem_child_face(et, j, cf)] if f >= 0}
        assert got == {(f, j, cf) for f in range(nf) for (j, cf) in mm.COARSE_TO_FINE_FACE[et][f]}
    for a, b in zip(hostapi.elem_tables(hostapi.HEX, "biquadratic"), hostapi.hex_tables("biquadratic")):
        assert np.array_equal(a, b)
    for order in ("linear", "quadratic", "biquadratic"):
        P = mm._local_prolongator(mm.HEX, order)
        for j in range(8):
            for a in range(P.shape[1]):
                idx, val = hostapi.elem_prolongator_row(hostapi.HEX, order, j, a)
                assert np.array_equal(idx, np.nonzero(P[j, a])[0]) and np.array_equal(val, P[j, a, idx])


@pytest.mark.parametrize("name", ["cube_wedge18", "cube_mixed", "cube_tet10", "cube_hex27_2x2x2", "cube_mixed_3groups"])
def test_unstructured_hierarchies_match_mixed_oracle(name):
    """The reference's shipped 3-D coarse meshes (re-serialised): 16 wedges; 4 hexahedra + 10 tetrahedra + 6 wedges
    sharing triangular and quadrilateral faces; 105 tetrahedra; 8 hexahedra.  Reader (+ nodes the file lacks),
    numbering, 1 -> 8 refinement, boundary flags, dof maps, Dirichlet flags, host sparsity: integers bit-exact
    against the independent oracle (oracle/mesh_mixed.py) on 3 levels; prolongators identical in structure."""
    from oracle import mesh_mixed as mm
    path = os.path.join(os.path.dirname(__file__), "golden", name + ".neu")
    H = hostapi.HostHierarchy.from_neu(path, 3)
    lv = mm.build_hierarchy(path, 3)
    for l, (h, L) in enumerate(zip(H.levels, lv)):
        assert h.nnode == L.nnode and np.array_equal(h.elem_types, L.etype)
        assert np.array_equal(h.conn, L.conn) and np.array_equal(h.face, L.face) and np.array_equal(h.dof_offset, L.dof_offset)
        assert np.abs(h.xyz - L.xyz).max() <= (0.0 if l == 0 else 1e-15)
        for order in TET_ORDERS:
            assert np.array_equal(h.system_dofs27(order), mm.system_dofs27(L, order))
            assert np.array_equal(h.bdc(order), mm.bdc_flags(L, order))
            assert np.array_equal(h.bdc(order, (2, 5)), mm.bdc_flags(L, order, (2, 5)))
            rp, ci = h.sparsity(order)
            rpo, cio = mm.sparsity(L, order)
            assert np.array_equal(rp, rpo) and np.array_equal(ci, cio)
        if l + 1 < len(lv):
            assert np.array_equal(h.child_el, lv[l + 1].child_el)
        v, e = h.dof_offset[0, -1], h.dof_offset[1, -1] - h.dof_offset[0, -1]
        nfaces = {0: 6, 1: 4, 2: 5}
        ncentre = h.nel
        f = h.nnode - h.dof_offset[1, -1] - ncentre
        assert v - e + f - h.nel == 1                                   # Euler characteristic of a ball
    for l in (1, 2):
        for order in TET_ORDERS:
            rp, ci, v, shp = H.prolongator(l, order)
            P = mm.prolongator(lv[l - 1], lv[l], order)
            assert shp == P.shape and np.array_equal(rp, P.indptr) and np.array_equal(ci, P.indices)
            assert np.abs(v - P.data).max() <= 1e-15
            assert np.abs(P @ np.ones(P.shape[1]) - 1.0).max() < 1e-14


def test_general_refinement_equals_hexahedral_refinement():
    """A generated box refined by the general (any element type) code and by the specialised hexahedral code:
    identical connectivity, numbering, flags, coordinates and prolongators."""
    A = hostapi.HostHierarchy.box_general(2, 3, 2, 3)
    B = hostapi.HostHierarchy(2, 3, 2, 3)
    for l in range(3):
        a, b = A.levels[l], B.levels[l]
        assert np.array_equal(a.conn, b.conn) and np.array_equal(a.face, b.face) and np.array_equal(a.dof_offset, b.dof_offset)
        assert np.array_equal(a.xyz, b.xyz)
        if l:
            for order in ("linear", "quadratic", "biquadratic"):
                pa, pb = A.prolongator(l, order), B.prolongator(l, order)
                assert all(np.array_equal(x, y) for x, y in zip(pa[:3], pb[:3]))


def test_tet_mesh_hierarchy_matches_oracle():
    """cube_tet10.neu (the reference's cube_Tet.neu re-serialised: 105 ten-node tetrahedra): reader with the
    face / centre nodes of AddBiquadraticNodesNotInMeshFile, first-visit numbering, 1 -> 8 refinement, boundary
    flags, dof maps, Dirichlet flags: integers bit-exact against the independent numpy oracle on 3 levels;
    coordinates exact on level 0 (same operation order) and to 1e-15 on refined levels; prolongators with
    identical structure and values to an ulp."""
    from oracle import mesh_tet as mt
    H = hostapi.HostHierarchy.from_neu(NEU_TET, 3)
    lv = mt.build_hierarchy(NEU_TET, 3)
    assert [h.elem_type for h in H.levels] == [hostapi.TET] * 3
    assert [h.nel for h in H.levels] == [105, 840, 6720]
    for l, (h, L) in enumerate(zip(H.levels, lv)):
        assert h.nnode == L.nnode
        assert np.array_equal(h.conn[:, :15], L.conn) and np.all(h.conn[:, 15:] == -1)
        assert np.array_equal(h.face[:, :4], L.face) and np.all(h.face[:, 4:] == -1)
        assert np.array_equal(h.dof_offset, L.dof_offset)
        if l == 0:
            assert np.array_equal(h.xyz, L.xyz)
        else:
            assert np.abs(h.xyz - L.xyz).max() <= 1e-15
        assert h.xyz.min() >= -1e-15 and h.xyz.max() <= 1 + 1e-15
        for order in TET_ORDERS:
            assert np.array_equal(h.system_dofs(order), mt.system_dof(L, order))
            assert np.array_equal(h.bdc(order), mt.bdc_flags(L, order))
            assert np.array_equal(h.bdc(order, (1, 4)), mt.bdc_flags(L, order, (1, 4)))
        if l + 1 < len(lv):
            assert np.array_equal(h.child_el, lv[l + 1].child_el)
    # Euler characteristic of a ball: V - E + F - C = 1 on every level (vertices, edges, faces, cells)
    for h in H.levels:
        v, e, f = h.dof_offset[0, -1], h.dof_offset[1, -1] - h.dof_offset[0, -1], h.nnode - h.dof_offset[1, -1] - h.nel
        assert v - e + f - h.nel == 1
    for l in (1, 2):
        for order in TET_ORDERS:
            rp, ci, v, shp = H.prolongator(l, order)
            P = mt.prolongator(lv[l - 1], lv[l], order)
            assert shp == P.shape and np.array_equal(rp, P.indptr) and np.array_equal(ci, P.indices)
            assert np.abs(v - P.data).max() <= 1e-15
            assert np.abs(P @ np.ones(P.shape[1]) - 1.0).max() < 1e-14          # constants are reproduced


def test_tet_refinement_is_nested():
    """Refined coordinates are the coarse FE geometry evaluated at the child nodes: on this affine mesh every
    fine vertex that is a coarse node keeps its coordinates and every element volume is 1/8 of its parent's."""
    H = hostapi.HostHierarchy.from_neu(NEU_TET, 2)
    C, F = H.levels
    def vol(L):
        X = L.xyz[:, L.conn[:, :4]]
        M = np.stack([X[:, :, k] - X[:, :, 0] for k in (1, 2, 3)], axis=-1)       # [3, nel, 3]
        return np.abs(np.linalg.det(M.transpose(1, 0, 2))) / 6
    vc, vf = vol(C), vol(F)
    assert abs(vc.sum() - 1.0) < 1e-13 and abs(vf.sum() - 1.0) < 1e-13
    # the file's mid-edge nodes are the edge midpoints only to its 12 printed digits
    assert np.abs(vf[C.child_el] - vc[:, None] / 8).max() < 1e-10 * vc.max()


# ---- face elements of every element type / Neumann face groups ---------------------------------------------
def test_face_kind_tables_and_face_nodes_match_oracle():
    """Host FaceElement (triangles 3/6/7 with 13 points, quadrilaterals 4/8/9 with 16) against oracle/fe_face.py
    (pinned to the compiled reference): weights bit-exact, 4/9-node quadrilateral tables bit-exact, the others to
    2e-15; face -> local node tables and face kinds of hexahedra, tetrahedra and wedges."""
    from oracle import fe_face, mesh_mixed as mm
    for kind, geom in ((hostapi.QUAD_FACE, "quad"), (hostapi.TRI_FACE, "tri")):
        for order in ("linear", "quadratic", "biquadratic"):
            got, want = hostapi.face_kind_tables(kind, order), fe_face.tables(geom, order)
            assert got[0].shape == want[0].shape and np.array_equal(got[3], want[3])
            for a, b in zip(got[:3], want[:3]):
                assert np.abs(a - b).max() <= (0.0 if geom == "quad" and order != "quadratic" else 2e-15)
            assert np.abs(got[0].sum(axis=1) - 1.0).max() < 1e-14          # partition of unity
    for t in (mm.HEX, mm.TET, mm.WEDGE):
        fn, fk = hostapi.elem_face_nodes(t), hostapi.elem_face_kinds(t)
        for f in range(6):
            if f >= mm.NFACES[t]:
                assert fk[f] == -1 and (fn[f] == -1).all()
                continue
            assert fk[f] == (hostapi.TRI_FACE if mm.FACE_NVERT[t][f] == 3 else hostapi.QUAD_FACE)
            nodes = list(mm.FACE_NODES[t][f])
            n = 7 if mm.FACE_NVERT[t][f] == 3 else 9
            assert fn[f, :n].tolist() == [int(x) for x in nodes[:n]] and (fn[f, n:] == -1).all()


def _neumann_groups_numpy(level, order, neumann):
    """What the driver hands to b2_asm_neumann_faces, evaluated with numpy in the kernel's operation order."""
    from femus_b200.poisson import neumann_face_groups
    mixed = level.elem_type < 0
    dofs = level.system_dofs27(order)
    rhs = np.zeros(level.ndofs(order))
    for t in (sorted(set(level.elem_types.tolist())) if mixed else [level.elem_type]):
        sel = np.nonzero(level.elem_types == t)[0] if mixed else slice(None)
        conn, dof = level.conn[sel], dofs[sel]
        nve = hostapi.elem_nve(t, order)
        for (fe, fl, fv), (phi, dxi, deta, w), fnodes in neumann_face_groups(level, order, neumann, t, sel):
            nvf = phi.shape[1]
            for e, f, v in zip(fe, fl, fv):
                loc = fnodes[f, :nvf]
                assert (loc >= 0).all() and (loc < nve).all()
                X = level.xyz[:, conn[e, loc]]
                J = np.stack([X @ dxi.T, X @ deta.T], axis=-1)             # [3][ng][2]
                n = np.cross(J[:, :, 0].T, J[:, :, 1].T)
                area = np.sqrt((n * n).sum(axis=1))
                np.add.at(rhs, dof[e, loc], (phi * (v * area * w)[:, None]).sum(axis=0))
    return rhs


@pytest.mark.parametrize("name", ["cube_tet10", "cube_wedge18", "cube_mixed", "cube_hex27_2x2x2"])
def test_neumann_face_groups_match_oracle(name):
    """The (plan, face kind) groups of Neumann faces the driver builds for b2_asm_neumann_faces -- plan-local
    element rows, local faces, host face tables, face-node tables -- reproduce the oracle's boundary vector
    (oracle/mesh_mixed.neumann_rhs, main.cpp:495-548) on tetrahedra, wedges, the mixed mesh and hexahedra."""
    from oracle import mesh_mixed as mm
    path = os.path.join(os.path.dirname(__file__), "golden", name + ".neu")
    H = hostapi.HostHierarchy.from_neu(path, 2)
    L = mm.build_hierarchy(path, 2)[-1]
    neumann = {1: 0.2, 4: -1.5, 6: 0.7}
    for order in ("linear", "quadratic", "biquadratic"):
        got = _neumann_groups_numpy(H.levels[-1], order, neumann)
        want = mm.neumann_rhs(L, order, neumann)
        assert np.abs(got - want).max() <= 1e-13 * np.abs(want).max()
        assert abs(got.sum() - (0.2 - 1.5 + 0.7)) < 1e-13
