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
        for fam in ("linear", "biquadratic"):
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
    for fam in ("linear", "biquadratic"):
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
