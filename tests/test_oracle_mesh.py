"""CPU: consistency of the restated mesh / dof / multigrid oracle (parity unpinned by the
reference, so these are structural properties the reference guarantees by construction)."""
import numpy as np
import pytest

from oracle import mesh_box as mb, mg, fe_hex


def test_box_numbering_first_visit():
    L = mb.build_box(3, 2, 2)
    # element 0 sees vertices 0..7 in local order, all vertices come first, then edges, then faces
    assert L.conn[0, :8].tolist() == list(range(8))
    nv = 4 * 3 * 3
    assert L.dof_offset[0].tolist() == [0, nv]
    assert L.conn[:, :8].max() == nv - 1
    assert L.conn[:, 8:20].min() >= nv
    assert L.dof_offset[2][-1] == 7 * 5 * 5
    # connectivity is a bijection onto the nodes
    assert np.unique(L.conn).shape[0] == L.nnode


@pytest.mark.parametrize("order,nnz", [("linear", lambda n: (3 * (n + 1) - 2) ** 3), ("biquadratic", lambda n: (8 * n + 1) ** 3)])
def test_sizes_match_survey_formulas(order, nnz):
    lv = mb.build_hierarchy(2, 2, 2, 3)
    F = lv[-1]
    n = F.n[0]
    rp, ci = mb.sparsity(F, order)
    assert rp[-1] == nnz(n)
    P = mb.prolongator(lv[1], F, order)
    if order == "biquadratic":
        assert P.nnz == (4 * n + 1) ** 3
    assert np.allclose(P.sum(axis=1), 1.0)


def test_refined_coordinates_are_lattice_points():
    lv = mb.build_hierarchy(2, 3, 2, 3, bounds=(0., 1., 0., 1., 0., 1.))
    F = lv[-1]
    sx, sy = 2 * F.n[0] + 1, 2 * F.n[1] + 1
    lat = F.lat_of_node
    exact = np.stack([(lat % sx) / (2 * F.n[0]), ((lat // sx) % sy) / (2 * F.n[1]), (lat // (sx * sy)) / (2 * F.n[2])])
    assert np.abs(F.xyz - exact).max() < 1e-15


def test_galerkin_equals_rediscretisation_on_nested_affine_meshes():
    lv = mb.build_hierarchy(2, 2, 2, 2)
    for order in ("linear", "biquadratic"):
        H = mg.Hierarchy(lv, order)
        A0, _ = mb.assemble(lv[0], order)
        free = H.bdc[0] > 1.5
        D = (H.A_raw[0] - A0).tocsr()[free][:, free]
        assert abs(D).max() < 1e-13


def test_vcycle_converges_and_partition_is_a_renumbering():
    lv = mb.build_hierarchy(2, 2, 2, 3)
    H = mg.Hierarchy(lv, "biquadratic")
    tr, eps = H.mg_solve_trace(6)
    assert all(b < a for a, b in zip(tr, tr[1:]))
    # two-rank slab partition: same operator up to a permutation of dofs
    lv2 = mb.build_hierarchy(2, 2, 2, 3, nprocs=2)
    A1, r1 = mb.assemble(lv[-1], "biquadratic")
    A2, r2 = mb.assemble(lv2[-1], "biquadratic")
    # match nodes through their lattice ids
    o1 = np.argsort(lv[-1].lat_of_node)
    o2 = np.argsort(lv2[-1].lat_of_node)
    perm = np.empty_like(o1)
    perm[o1] = o2                       # node id in numbering 1 -> node id in numbering 2
    assert np.allclose(r1, r2[perm], atol=1e-18)
    B = A2.tocsr()[perm][:, perm]
    assert abs(A1 - B).max() < 1e-15
    # rank 0 owns its nodes first: ownership offsets are monotone and cover all nodes
    assert lv2[-1].dof_offset[2].tolist()[0] == 0 and lv2[-1].dof_offset[2][-1] == lv2[-1].nnode


def test_dirichlet_flags_all_faces():
    lv = mb.build_hierarchy(2, 2, 2, 2)
    F = lv[-1]
    for order, cnt in (("linear", 5 ** 3 - 3 ** 3), ("biquadratic", 9 ** 3 - 7 ** 3)):
        b = mb.bdc_flags(F, order)
        assert int((b < 1.5).sum()) == cnt
    b = mb.bdc_flags(F, "biquadratic", dirichlet_faces=(1,))        # bottom only
    assert int((b < 1.5).sum()) == 9 * 9
    z = F.xyz[2][np.nonzero(b < 1.5)[0]]
    assert np.all(z == 0.0)
