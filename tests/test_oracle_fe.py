"""CPU: pin the numpy FE oracle (oracle/fe_hex.py) against the reference's own FE kernel --
the committed golden fixture (generated from oracle/_ref by tests/golden/make_fe_golden.py) and,
when the compiled reference is present, oracle/_ref itself."""
import os
import numpy as np
import pytest

from oracle import fe_hex, ref

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "fe_hex_ref.npz"))
ORDERS = ("linear", "biquadratic")
HEX_ORDERS = ("linear", "quadratic", "biquadratic")      # 8, 20 (serendipity) and 27 dofs


@pytest.mark.parametrize("order", HEX_ORDERS)
def test_gauss_and_tables_bit_exact(order):
    w, xi = fe_hex.gauss_hex("seventh")
    assert np.array_equal(w, GOLD[f"{order}_gauss_w"])
    assert np.array_equal(xi, GOLD[f"{order}_gauss_xi"])
    phi, dxi, deta, dzeta, _ = fe_hex.tables(order)
    for a, k in ((phi, "phi"), (dxi, "dxi"), (deta, "deta"), (dzeta, "dzeta")):
        assert np.array_equal(a, GOLD[f"{order}_{k}"]), k


@pytest.mark.parametrize("order", HEX_ORDERS)
def test_jacobian_bit_exact(order):
    X = GOLD[f"{order}_X"]
    tabs = fe_hex.tables(order)
    for ig in range(64):
        w, _, g = fe_hex.jacobian(order, X, ig, tabs)
        assert np.array_equal(w, GOLD[f"{order}_weight"][:, ig])
        assert np.array_equal(g, GOLD[f"{order}_gradphi"][:, ig])


@pytest.mark.parametrize("order", HEX_ORDERS)
def test_poisson_element_bit_exact(order):
    F, B = fe_hex.poisson_elements(order, GOLD[f"{order}_X"], GOLD[f"{order}_U"], 1.0)
    assert np.array_equal(B, GOLD[f"{order}_B"])
    assert np.array_equal(F, GOLD[f"{order}_F"])


@pytest.mark.parametrize("order", HEX_ORDERS)
def test_local_prolongator(order):
    P = fe_hex.local_prolongator(order)
    pts = fe_hex.fine_points(order)
    lut = {tuple(p): i for i, p in enumerate(pts)}
    Pg, pos = GOLD[f"{order}_prol"], GOLD[f"{order}_prol_pos2"]
    assert Pg.shape == P.shape
    for i in range(Pg.shape[0]):
        assert np.array_equal(P[lut[tuple(pos[i])]], Pg[i])
    assert int((P != 0).sum()) == {"linear": 64, "quadratic": 472, "biquadratic": 729}[order]


def test_survey_known_answers():
    """Numbers measured from the compiled reference and recorded in SURVEY.md section 8c."""
    n = 27
    for N, vol, b00, b01, b2626, fro in ((128, 4.7683715820312267e-07, 0.00097222222222222198,
                                          -0.00011574074074074252, 0.035555555555555, 0.048678499916971554),
                                         (256, None, 0.00048611111111111099, None, None, 0.024339249958485777)):
        X = (fe_hex.XC[:n].T + 1.0) * (0.5 / N)
        F, B = fe_hex.poisson_elements("biquadratic", X[None], np.zeros((1, n)), 1.0)
        assert abs(B[0, 0, 0] - b00) < 1e-17
        assert abs(np.linalg.norm(B[0]) - fro) < 1e-15
        if vol is not None:
            assert abs(F[0].sum() - vol) < 1e-20
            assert abs(B[0, 0, 1] - b01) < 1e-17
            assert abs(B[0, 26, 26] - b2626) < 1e-14
    X = (fe_hex.XC[:8].T + 1.0) * (0.5 / 32)
    F, B = fe_hex.poisson_elements("linear", X[None], np.zeros((1, 8)), 1.0)
    assert abs(F[0].sum() - 3.0517578124999851e-05) < 1e-19
    assert abs(B[0, 0, 0] - 0.010416666666666642) < 1e-16
    assert abs(np.linalg.norm(B[0]) - 0.032940392293420523) < 1e-15


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (no /root/reference here)")
@pytest.mark.parametrize("order", HEX_ORDERS)
def test_against_compiled_reference(order):
    R = ref.RefHex(order)
    rng = np.random.default_rng(7)
    n = R.n
    X = (fe_hex.XC[:n].T * 0.1 + 0.5) + rng.standard_normal((3, n)) * 0.01
    U = rng.standard_normal(n)
    Fr, Br = R.poisson_element(X, U, 2.5)
    Fo, Bo = fe_hex.poisson_elements(order, X[None], U[None], 2.5)
    assert np.array_equal(Br, Bo[0]) and np.array_equal(Fr, Fo[0])


# ---- face element / Neumann boundary integral -----------------------------------------------
GQ = np.load(os.path.join(os.path.dirname(__file__), "golden", "fe_quad_ref.npz"))


@pytest.mark.parametrize("order", ORDERS)
def test_face_element_bit_exact_vs_golden(order):
    """oracle/fe_quad.py against the committed values of the reference's elem_type_2D: Gauss rule,
    tables, JacobianSur weight/normal and the Neumann face vector with bdc_func = 0.2."""
    from oracle import fe_quad
    w, xi = fe_quad.gauss_quad()
    assert np.array_equal(w, GQ[f"{order}_gauss_w"]) and np.array_equal(xi, GQ[f"{order}_gauss_xi"])
    phi, dxi, deta, _ = fe_quad.tables2(order)
    assert np.array_equal(phi, GQ[f"{order}_phi"]) and np.array_equal(dxi, GQ[f"{order}_dxi"]) and np.array_equal(deta, GQ[f"{order}_deta"])
    for k, X in enumerate(GQ[f"{order}_X"]):
        for ig in range(16):
            wt, _, nrm = fe_quad.jacobian_sur(order, X, ig)
            assert wt == GQ[f"{order}_weight"][k, ig] and np.array_equal(nrm, GQ[f"{order}_normal"][k, ig])
        assert np.array_equal(fe_quad.neumann_face(order, X, 0.2), GQ[f"{order}_F02"][k])


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (no /root/reference here)")
@pytest.mark.parametrize("order", ORDERS)
def test_face_element_against_compiled_reference(order):
    from oracle import fe_quad
    Q = ref.RefQuad(order)
    rng = np.random.default_rng(2)
    base = np.array([[0, 1, 1, 0, 0.5, 1, 0.5, 0, 0.5], [0, 0, 1, 1, 0, 0.5, 1, 0.5, 0.5], [0.3] * 9])
    for _ in range(4):
        X = base + 0.05 * rng.standard_normal(base.shape)
        for ig in range(16):
            a = Q.jacobian_sur(X[:, :Q.n].copy(), ig)
            b = fe_quad.jacobian_sur(order, X, ig)
            assert a[0] == b[0] and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])


def test_neumann_rhs_integrates_the_flux():
    """Sum of the Neumann vector = flux x area of the face set (partition of unity), any family."""
    from oracle import mesh_box as mb
    lv = mb.build_hierarchy(2, 3, 2, 2)
    for order in ORDERS:
        r = mb.neumann_rhs(lv[-1], order, {3: 0.2, 6: -1.5})
        assert abs(r.sum() - (0.2 * 1.0 - 1.5 * 1.0)) < 1e-13


# ---- face elements of tetrahedra and wedges (triangles 3 / 6 / 7, quadrilaterals 4 / 8 / 9) ----------------
GF = np.load(os.path.join(os.path.dirname(__file__), "golden", "fe_face_ref.npz"))
FACE_CASES = [(g, o) for g in ("tri", "quad") for o in ("linear", "quadratic", "biquadratic")]


@pytest.mark.parametrize("geom,order", FACE_CASES)
def test_any_face_element_vs_golden(geom, order):
    """oracle/fe_face.py against the committed values of the reference's elem_type_2D(geom, order, "seventh"):
    Gauss rule bit-exact; tables, JacobianSur weight / normal and the Neumann face vector (bdc_func = 0.2) to
    2e-15 (the triangle functions are restated as products of barycentrics, not the reference's expanded
    polynomials); the 4 / 9-node quadrilaterals stay bit-exact."""
    from oracle import fe_face
    k = f"{geom}_{order}"
    phi, dxi, deta, w = fe_face.tables(geom, order)
    assert np.array_equal(w, GF[f"{k}_gauss_w"])
    assert phi.shape == GF[f"{k}_phi"].shape == (13 if geom == "tri" else 16, fe_face.ndofs(geom, order))
    exact = geom == "quad" and order != "quadratic"
    tol = 0.0 if exact else 2e-15
    for a, b in ((phi, GF[f"{k}_phi"]), (dxi, GF[f"{k}_dxi"]), (deta, GF[f"{k}_deta"])):
        assert np.abs(a - b).max() <= tol
    for j, X in enumerate(GF[f"{k}_X"]):
        for ig in range(w.shape[0]):
            wt, ph, nrm = fe_face.jacobian_sur(X, ig, (phi, dxi, deta, w))
            assert abs(wt - GF[f"{k}_weight"][j, ig]) <= 1e-14 * abs(GF[f"{k}_weight"][j, ig]) + tol
            assert np.abs(nrm - GF[f"{k}_normal"][j, ig]).max() <= 1e-14
        F = fe_face.neumann_face(X, 0.2, (phi, dxi, deta, w))
        assert np.abs(F - GF[f"{k}_F02"][j]).max() <= 1e-14 * np.abs(GF[f"{k}_F02"][j]).max()
        if exact:
            assert np.array_equal(F, GF[f"{k}_F02"][j])


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (no /root/reference here)")
@pytest.mark.parametrize("geom,order", FACE_CASES)
def test_any_face_element_against_compiled_reference(geom, order):
    from oracle import fe_face
    Q = ref.RefFace(geom, order)
    tabs = fe_face.tables(geom, order)
    assert (Q.n, Q.ng) == (fe_face.ndofs(geom, order), tabs[3].shape[0])
    rng = np.random.default_rng(5)
    base = GF[f"{geom}_{order}_X"][0] * 64.0
    for _ in range(4):
        X = base + 0.05 * rng.standard_normal(base.shape)
        for ig in range(Q.ng):
            a = Q.jacobian_sur(X.copy(), ig)
            b = fe_face.jacobian_sur(X, ig, tabs)
            assert abs(a[0] - b[0]) <= 1e-14 * abs(a[0]) and np.abs(a[1] - b[1]).max() <= 2e-15 and np.abs(a[2] - b[2]).max() <= 1e-14


@pytest.mark.parametrize("mesh", ["cube_tet10", "cube_wedge18", "cube_mixed", "cube_hex27_2x2x2"])
def test_neumann_rhs_on_any_element_type_integrates_the_flux(mesh):
    """Sum of the Neumann vector = flux x area of the face sets (partition of unity of the face bases) on
    tetrahedra (triangular faces), wedges (both kinds) and the mixed mesh, every family; on hexahedra the
    general loop reproduces mesh_box.neumann_rhs's face vectors."""
    from oracle import mesh_mixed as mm
    L = mm.build_hierarchy(os.path.join(os.path.dirname(__file__), "golden", mesh + ".neu"), 2)[-1]
    for order in ("linear", "quadratic", "biquadratic"):
        r = mm.neumann_rhs(L, order, {1: 0.2, 3: -1.5})
        assert abs(r.sum() - (0.2 - 1.5)) < 1e-13
        assert np.count_nonzero(r) > 0


# ------------------------------------------------------------------------------ tetrahedra and wedges
from oracle import fe_tet, fe_wedge  # noqa: E402

GOLD_G = {"tet": np.load(os.path.join(os.path.dirname(__file__), "golden", "fe_tet_ref.npz")),
          "wedge": np.load(os.path.join(os.path.dirname(__file__), "golden", "fe_wedge_ref.npz"))}
FE_G = {"tet": fe_tet, "wedge": fe_wedge}
GEOMS = ("tet", "wedge")
TET_ORDERS = ("linear", "quadratic", "biquadratic")
TET_ULP = 3e-15      # fe_tet.py / fe_wedge.py restate the mathematics, not the reference's operation order


@pytest.mark.parametrize("geom", GEOMS)
@pytest.mark.parametrize("order", TET_ORDERS)
def test_simplex_gauss_and_tables(geom, order):
    """31-point (tet) / 52-point (wedge) rule bit-exact (data), shape tables to a few ulp; partition of unity;
    Kronecker property at the element's own nodes."""
    fe, G = FE_G[geom], GOLD_G[geom]
    w, xi = (fe.gauss_tet if geom == "tet" else fe.gauss_wedge)("seventh")
    assert np.array_equal(w, G[f"{order}_gauss_w"]) and np.array_equal(xi, G[f"{order}_gauss_xi"])
    phi, dxi, deta, dzeta, _ = fe.tables(order)
    for a, k in ((phi, "phi"), (dxi, "dxi"), (deta, "deta"), (dzeta, "dzeta")):
        assert np.abs(a - G[f"{order}_{k}"]).max() <= TET_ULP, k
    assert np.abs(phi.sum(axis=1) - 1.0).max() < 1e-14
    n = fe.NDOFS[order]
    assert np.abs(fe.shape(order, fe.XC[:n])[0] - np.eye(n)).max() < 1e-14


@pytest.mark.parametrize("geom", GEOMS)
@pytest.mark.parametrize("order", TET_ORDERS)
def test_simplex_jacobian_and_poisson_element(geom, order):
    fe, G = FE_G[geom], GOLD_G[geom]
    X, U = G[f"{order}_X"], G[f"{order}_U"]
    tabs = fe.tables(order)
    for ig in range(tabs[4].shape[0]):
        w, _, g = fe_hex.jacobian(order, X, ig, tabs)
        assert np.abs(w - G[f"{order}_weight"][:, ig]).max() <= 1e-14 * np.abs(G[f"{order}_weight"]).max()
        gr = G[f"{order}_gradphi"][:, ig]
        assert np.abs(g - gr).max() <= 1e-13 * np.abs(gr).max()
    F, B = fe_hex.poisson_elements(order, X, U, 1.0, tabs)
    Br, Fr = G[f"{order}_B"], G[f"{order}_F"]
    for k in range(X.shape[0]):
        assert np.abs(B[k] - Br[k]).max() <= 1e-13 * np.abs(Br[k]).max()
        assert np.abs(F[k] - Fr[k]).max() <= 1e-13 * (np.abs(Br[k]) @ np.abs(U[k])).max()


@pytest.mark.parametrize("geom", GEOMS)
@pytest.mark.parametrize("order", TET_ORDERS)
def test_simplex_local_prolongator(geom, order):
    """Element prolongator of the 8 children: same non-zero structure as the reference's, values to an ulp;
    fine-dof counts and nnz as measured from the compiled reference (tet 10/35/67 and 16/116/447,
    wedge 18/57/95 and 36/264/603)."""
    fe, G = FE_G[geom], GOLD_G[geom]
    P = fe.local_prolongator(order)
    Pg, kv = G[f"{order}_prol"], G[f"{order}_prol_kvert"]
    for i in range(Pg.shape[0]):
        mine = P[kv[i, 0], kv[i, 1]]
        assert np.array_equal(mine != 0, Pg[i] != 0)
        assert np.abs(mine - Pg[i]).max() <= 4e-16
    expect = {"tet": {"linear": (10, 16), "quadratic": (35, 116), "biquadratic": (67, 447)},
              "wedge": {"linear": (18, 36), "quadratic": (57, 264), "biquadratic": (95, 603)}}
    assert (Pg.shape[0], int((Pg != 0).sum())) == expect[geom][order]
    # every (child, node) pair maps to one of the nf distinct fine points
    pts = fe.child_points(order).reshape(-1, 3)
    assert len({tuple(np.round(p * 24).astype(int)) for p in pts}) == Pg.shape[0]


@pytest.mark.skipif(not ref.available(), reason="compiled reference (oracle/_ref) not present")
@pytest.mark.parametrize("geom", GEOMS)
@pytest.mark.parametrize("order", TET_ORDERS)
def test_simplex_against_compiled_reference(geom, order):
    fe = FE_G[geom]
    R = ref.RefElem(geom, order)
    rng = np.random.default_rng(5)
    n = R.n
    X = (np.eye(3) + 0.2 * rng.standard_normal((3, 3))) @ fe.XC[:n].T + 0.01 * rng.standard_normal((3, n))
    U = rng.standard_normal(n)
    Fr, Br = R.poisson_element(X, U, 2.5)
    F, B = fe_hex.poisson_elements(order, X[None], U[None], 2.5, fe.tables(order))
    assert np.abs(B[0] - Br).max() <= 1e-13 * np.abs(Br).max()
    assert np.abs(F[0] - Fr).max() <= 1e-13 * (np.abs(Br) @ np.abs(U) + 2.5).max()
