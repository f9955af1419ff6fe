"""GPU parity of the table-driven assembly kernel (b2_assemble.cu: assemble_general_kernel) -- the path of
every element family that is not a hexahedron with the 64-point rule (tetrahedra) --
against the CPU oracle on the same seeded inputs, through the C ABI.  Bar: CSR structure bit-exact,
values / residuals to 1e-12 relative."""
import os
import numpy as np
import pytest
import scipy.sparse as sp

from femus_b200 import capi
from oracle import fe_hex, fe_tet, mesh_box as mb

pytestmark = pytest.mark.gpu
RTOL = 1e-12
GOLD_TET = np.load(os.path.join(os.path.dirname(__file__), "golden", "fe_tet_ref.npz"))


def oracle_assemble(xyz, conn, dof, n, U, fsrc, tabs):
    nve = dof.shape[1]
    X = xyz[:, conn[:, :nve]].transpose(1, 0, 2)
    F, B = fe_hex.poisson_elements(None, X, U[dof], fsrc, tabs)
    rows = np.repeat(dof, nve, axis=1).ravel()
    cols = np.tile(dof, (1, nve)).ravel()
    A = sp.csr_matrix((B.ravel(), (rows, cols)), shape=(n, n))
    A.sum_duplicates()
    A.sort_indices()
    rhs = np.zeros(n)
    np.add.at(rhs, dof.ravel(), F.ravel())
    return A, rhs


def tet_soup(rng, nel, nnode, order):
    """Random conforming-agnostic 'soup': every element is a distorted copy of the reference tetrahedron,
    elements share dofs at random (what the kernel sees is only conn / dof / xyz)."""
    nve = fe_tet.NDOFS[order]
    conn = np.full((nel, 27), -1, dtype=np.int32)
    xyz = np.zeros((3, nnode + 15 * nel))
    dof = np.zeros((nel, nve), dtype=np.int32)
    ndof = max(nve, nnode)
    for e in range(nel):
        M = np.eye(3) + 0.2 * rng.standard_normal((3, 3))
        if np.linalg.det(M) < 0:
            M[:, 0] = -M[:, 0]
        ids = nnode + 15 * e + np.arange(15)
        xyz[:, ids] = M @ fe_tet.XC.T * 0.1 + rng.uniform(0, 1, (3, 1)) + 0.0005 * rng.standard_normal((3, 15))
        conn[e, :15] = ids
        dof[e] = rng.choice(ndof, nve, replace=False)
    conn[conn < 0] = 0           # padding entries of the 27-wide rows are never read
    return xyz, conn, dof, ndof


@pytest.mark.parametrize("order", ["linear", "quadratic", "biquadratic"])
@pytest.mark.parametrize("nel", [1, 7, 400])
def test_tet_assembly_matches_oracle(ctx, order, nel):
    rng = np.random.default_rng(100 + nel)
    xyz, conn, dof, n = tet_soup(rng, nel, 60 if nel < 100 else 800, order)      # <= ~10 elements per dof
    tabs = fe_tet.tables(order)
    u = rng.standard_normal(n)
    Aref, rhs_ref = oracle_assemble(xyz, conn, dof, n, u, 1.5, tabs)
    A = capi.Csr.from_elements(ctx, n, dof)
    asm = capi.Assembler(capi.Mesh(ctx, xyz, conn), A, dof, tabs)
    U, R = ctx.vector(u), ctx.vector(n)
    asm.poisson(U, R, nu=1.0, fsrc=1.5)
    got = A.to_scipy()
    assert np.array_equal(got.indptr, Aref.indptr) and np.array_equal(got.indices, Aref.indices)
    assert np.abs(got.data - Aref.data).max() <= RTOL * np.abs(Aref.data).max()
    mag = np.abs(Aref) @ np.abs(u) + np.abs(rhs_ref)
    assert np.abs(R.get() - rhs_ref).max() <= RTOL * mag.max()
    # accumulate semantics and nu scaling: a second pass with nu = 2 adds twice the matrix
    asm.poisson(U, None, nu=2.0, fsrc=1.5)
    assert np.abs(A.to_scipy().data - 3 * Aref.data).max() <= 3 * RTOL * np.abs(Aref.data).max()


def test_tet_golden_elements(ctx):
    """Single tetrahedra of the committed fixture: element matrices / residuals of the compiled reference."""
    for order in ("linear", "quadratic", "biquadratic"):
        nve = fe_tet.NDOFS[order]
        tabs = fe_tet.tables(order)
        for k in range(GOLD_TET[f"{order}_X"].shape[0]):
            xyz = np.zeros((3, 27))
            xyz[:, :nve] = GOLD_TET[f"{order}_X"][k]
            conn = np.arange(27, dtype=np.int32)[None, :]
            d = np.arange(nve, dtype=np.int32)[None, :]
            A = capi.Csr.from_elements(ctx, nve, d)
            asm = capi.Assembler(capi.Mesh(ctx, xyz, conn), A, d, tabs)
            Uk = GOLD_TET[f"{order}_U"][k]
            U, R = ctx.vector(Uk), ctx.vector(nve)
            asm.poisson(U, R, 1.0, 1.0)
            B = A.to_scipy().toarray()
            Bref, Fref = GOLD_TET[f"{order}_B"][k], GOLD_TET[f"{order}_F"][k]
            assert np.abs(B - Bref).max() <= RTOL * np.abs(Bref).max()
            assert np.abs(R.get() - Fref).max() <= RTOL * (np.abs(Bref) @ np.abs(Uk)).max()


@pytest.mark.parametrize("order", ["linear", "biquadratic"])
def test_general_kernel_on_hexahedra(ctx, order):
    """asm_variant 2 sends hexahedra through the table-driven kernel too: same oracle, same bar as the
    specialised kernels (tests/test_gpu_parity.py::test_assembly_matches_oracle)."""
    lv = mb.build_hierarchy(2, 3, 2, 2)
    L = lv[-1]
    x, y, z = L.xyz
    L.xyz = L.xyz + 0.05 * np.stack([np.sin(np.pi * y) * x * (1 - x), np.sin(np.pi * z) * y * (1 - y), np.sin(np.pi * x) * z * (1 - z)])
    n = mb.ndofs(L, order)
    d = mb.system_dof(L, order)
    u = np.random.default_rng(3).standard_normal(n)
    Aref, rhs_ref = mb.assemble(L, order, u, fsrc=0.7)
    ctx.set_option("asm_variant", 2)
    try:
        A = capi.Csr.from_elements(ctx, n, d)
        asm = capi.Assembler(capi.Mesh(ctx, L.xyz, L.conn), A, d, fe_hex.tables(order))
        U, R = ctx.vector(u), ctx.vector(n)
        asm.poisson(U, R, nu=1.0, fsrc=0.7)
    finally:
        ctx.set_option("asm_variant", 3)
    got = A.to_scipy()
    assert np.array_equal(got.indices, Aref.indices)
    assert np.abs(got.data - Aref.data).max() <= RTOL * np.abs(Aref.data).max()
    assert np.abs(R.get() - rhs_ref).max() <= RTOL * (np.abs(Aref) @ np.abs(u) + np.abs(rhs_ref)).max()


def test_general_plan_rejects_bad_sizes(ctx):
    """Error behaviour: tables beyond 64 Gauss points are refused with a status (no silent fallback)."""
    rng = np.random.default_rng(0)
    xyz, conn, dof, n = tet_soup(rng, 4, 30, "quadratic")
    A = capi.Csr.from_elements(ctx, n, dof)
    mesh = capi.Mesh(ctx, xyz, conn)
    phi, dxi, deta, dzeta, w = fe_tet.tables("quadratic")
    big = np.zeros((65, 10))
    with pytest.raises(capi.B2Error):
        capi.Assembler(mesh, A, dof, (big, big, big, big, np.zeros(65)))


# ------------------------------------------------------------------------------ tetrahedral meshes end to end
NEU_TET = os.path.join(os.path.dirname(__file__), "golden", "cube_tet10.neu")


@pytest.mark.parametrize("order", ["linear", "quadratic", "biquadratic"])
def test_tet_mesh_single_level_matches_oracle(ctx, order):
    """cube_tet10.neu (the reference's cube_Tet.neu, 105 Tet10 elements, re-serialised) through the host reader
    and the driver sequence of 001_Poisson: pattern bit-exact, matrix and residual to 1e-12 against the oracle,
    single-level solve (the reference: LU) against a sparse direct solve."""
    import scipy.sparse.linalg as spla
    from femus_b200 import hostapi
    from femus_b200.poisson import PoissonMG
    from oracle import mesh_tet as mt, mg
    H = hostapi.HostHierarchy.from_neu(NEU_TET, 1)
    pb = PoissonMG(ctx, 0, 0, 0, 1, order, hier=H, coarse_rtol=1e-15)
    pb.assemble()
    L = mt.read_tet10(NEU_TET)
    Aref, rhs = mt.assemble(L, order)
    A = pb.KK[-1].to_scipy()
    assert np.array_equal(A.indptr, Aref.indptr) and np.array_equal(A.indices, Aref.indices)
    assert np.abs(A.data - Aref.data).max() <= RTOL * np.abs(Aref.data).max()
    assert np.abs(pb.RES.get() - rhs).max() <= RTOL * np.abs(rhs).max()
    pb.galerkin(); pb.mg_set_levels(); pb.mg_solve()
    idx = np.nonzero(mt.bdc_flags(L, order) < 1.5)[0]
    Ap = mg.penalty_fast(Aref, idx)
    b = rhs.copy(); b[idx] = 0.0
    x = spla.spsolve(Ap.tocsc(), b)
    assert np.abs(pb.EPS.get() - x).max() <= 1e-10 * np.abs(x).max()
    del pb


@pytest.mark.parametrize("order,nl", [("linear", 3), ("quadratic", 3), ("biquadratic", 2)])
def test_tet_mesh_vcycle_trace(ctx, order, nl):
    """Multigrid on refined tetrahedra: assembly on the finest level (table-driven kernel), Galerkin chain by
    the general triple product (what MatPtAP does), penalised level operators and the residual trace of six
    V-cycles against the oracle (exact coarse solve), 1e-12 relative."""
    from femus_b200 import hostapi
    from femus_b200.poisson import PoissonMG
    from oracle import mesh_tet as mt, mg
    H = hostapi.HostHierarchy.from_neu(NEU_TET, nl)
    pb = PoissonMG(ctx, 0, 0, 0, nl, order, hier=H, coarse_rtol=1e-15)
    assert not pb.fused and pb.nve == fe_tet.NDOFS[order]
    pb.assemble(); pb.galerkin(); pb.mg_set_levels()
    O = mg.Hierarchy(mt.build_hierarchy(NEU_TET, nl), order, mesh=mt)
    for l in range(nl):
        got, ref = pb.KK[l].to_scipy(), O.A[l]
        assert np.array_equal(got.indptr, ref.indptr) and np.array_equal(got.indices, ref.indices)
        assert np.abs(got.data - ref.data).max() <= RTOL * np.abs(ref.data).max()
    trace_ref, eps_ref = O.mg_solve_trace(6)
    trace = []
    for _ in range(6):
        pb.mg_solve()
        trace.append(pb.residual_norm())
    for a, b in zip(trace, trace_ref):
        assert abs(a - b) <= RTOL * trace_ref[0], (trace, trace_ref)
    assert np.abs(pb.EPS.get() - eps_ref).max() <= 1e-11 * np.abs(eps_ref).max()
    del pb


def test_tet_and_hex_solutions_agree(ctx):
    """The same boundary-value problem (-lap u = 1, u = 0 on the unit cube) on the tetrahedral and on the
    hexahedral coarse mesh of the reference, two refinements each, quadratic elements, solved to convergence:
    the two discrete solutions agree at the cube centre and in the mean to discretisation accuracy."""
    from femus_b200 import hostapi
    from femus_b200.poisson import PoissonMG
    vals = []
    for path, order in ((NEU_TET, "quadratic"), (os.path.join(os.path.dirname(NEU_TET), "cube_hex27_2x2x2.neu"), "biquadratic")):
        H = hostapi.HostHierarchy.from_neu(path, 3)
        pb = PoissonMG(ctx, 0, 0, 0, 3, order, hier=H, npre=2, npost=2, coarse_rtol=1e-14)
        pb.assemble(); pb.galerkin(); pb.mg_set_levels()
        r0 = pb.residual_norm()
        for _ in range(200):
            pb.mg_solve()
            if pb.residual_norm() <= 1e-10 * r0:
                break
        assert pb.residual_norm() <= 1e-10 * r0
        u = pb.EPS.get()
        top = H.levels[-1]
        # one rank: the dofs of a family are the first dof_offset[family] nodes ([vertices][edges][faces, centres])
        xyz = top.xyz[:, :top.dof_offset[hostapi.FAMILY[order], -1]]
        c = np.argmin(np.abs(xyz - 0.5).sum(axis=0))
        assert np.abs(xyz[:, c] - 0.5).max() < 1e-12
        vals.append(u[c])
        del pb
    # u(1/2,1/2,1/2) of the continuous problem is 0.0562128...
    assert abs(vals[0] - 0.0562128) < 2e-4 and abs(vals[1] - 0.0562128) < 2e-4
    assert abs(vals[0] - vals[1]) < 2e-4


# ------------------------------------------------------------------------------ 20-node hexahedra
@pytest.mark.parametrize("shape,nl", [((2, 3, 2), 2), ((2, 2, 2), 3)])
def test_hex20_assembly_and_vcycle_trace(ctx, shape, nl):
    """The serendipity family on hexahedra (fe_order "serendipity" of the reference's input3D_Hex_serendipity.json,
    20 dofs per element on the 27-node geometry): table-driven assembly, general triple product, V-cycle trace
    against the oracle, whose 20-node tables are bit-exact with the compiled reference.  Richardson scale 0.3
    (SetRichardsonScaleFactor): the largest eigenvalue of D^-1 A is 4.18 for this family, so the default 0.5
    diverges -- in the oracle and here alike."""
    from femus_b200 import hostapi
    from femus_b200.poisson import PoissonMG
    from oracle import mg
    order = "quadratic"
    H = hostapi.HostHierarchy(*shape, nl)
    pb = PoissonMG(ctx, 0, 0, 0, nl, order, hier=H, coarse_rtol=1e-15, omega=0.3)
    assert pb.nve == 20 and not pb.fused
    pb.assemble()
    lv = mb.build_hierarchy(*shape, nl)
    Aref, rhs = mb.assemble(lv[-1], order)
    A = pb.KK[-1].to_scipy()
    assert np.array_equal(A.indptr, Aref.indptr) and np.array_equal(A.indices, Aref.indices)
    assert np.abs(A.data - Aref.data).max() <= RTOL * np.abs(Aref.data).max()
    assert np.abs(pb.RES.get() - rhs).max() <= RTOL * np.abs(rhs).max()
    pb.galerkin(); pb.mg_set_levels()
    O = mg.Hierarchy(lv, order)
    for l in range(nl):
        got, ref = pb.KK[l].to_scipy(), O.A[l]
        assert np.array_equal(got.indices, ref.indices)
        assert np.abs(got.data - ref.data).max() <= RTOL * np.abs(ref.data).max()
    trace_ref, eps_ref = O.mg_solve_trace(6, omega=0.3)
    trace = []
    for _ in range(6):
        pb.mg_solve()
        trace.append(pb.residual_norm())
    for a, b in zip(trace, trace_ref):
        assert abs(a - b) <= RTOL * trace_ref[0], (trace, trace_ref)
    assert np.abs(pb.EPS.get() - eps_ref).max() <= 1e-11 * np.abs(eps_ref).max()
    del pb


# ------------------------------------------------------------------------------ wedges and mixed meshes
GOLDEN = os.path.dirname(NEU_TET)


def test_wedge_golden_elements(ctx):
    """Single wedges of the committed fixture (element matrices / residuals of the compiled reference,
    6 / 15 / 21 dofs, 52 Gauss points) through the table-driven kernel."""
    from oracle import fe_wedge
    G = np.load(os.path.join(GOLDEN, "fe_wedge_ref.npz"))
    for order in ("linear", "quadratic", "biquadratic"):
        nve = fe_wedge.NDOFS[order]
        tabs = fe_wedge.tables(order)
        for k in range(G[f"{order}_X"].shape[0]):
            xyz = np.zeros((3, 27))
            xyz[:, :nve] = G[f"{order}_X"][k]
            conn = np.arange(27, dtype=np.int32)[None, :]
            d = np.arange(nve, dtype=np.int32)[None, :]
            A = capi.Csr.from_elements(ctx, nve, d)
            asm = capi.Assembler(capi.Mesh(ctx, xyz, conn), A, d, tabs)
            Uk = G[f"{order}_U"][k]
            U, R = ctx.vector(Uk), ctx.vector(nve)
            asm.poisson(U, R, 1.0, 1.0)
            B = A.to_scipy().toarray()
            Bref, Fref = G[f"{order}_B"][k], G[f"{order}_F"][k]
            assert np.abs(B - Bref).max() <= RTOL * np.abs(Bref).max()
            assert np.abs(R.get() - Fref).max() <= RTOL * (np.abs(Bref) @ np.abs(Uk)).max()


@pytest.mark.parametrize("name", ["cube_wedge18", "cube_mixed", "cube_mixed_3groups"])
@pytest.mark.parametrize("order", ["linear", "quadratic", "biquadratic"])
def test_wedge_and_mixed_mesh_single_level(ctx, name, order):
    """The reference's cube_Wedge.neu (16 wedges) and cube_all_shapes_Six_boundary_groups.neu (4 hexahedra, 10
    tetrahedra, 6 wedges), re-serialised: one assembly plan per element type accumulating into one matrix;
    pattern bit-exact, matrix and residual to 1e-12 against the oracle, single-level solve against a sparse
    direct solve.  The current solution is non-zero, so the residual carries B u of every element type.
    cube_mixed_3groups: the same mesh cut into three element groups, which the reference reorders by
    (material, group, index) (Mesh.cpp:621-702) -- another element order and numbering, same kernels."""
    import scipy.sparse.linalg as spla
    from femus_b200 import hostapi
    from femus_b200.poisson import PoissonMG
    from oracle import mesh_mixed as mm, mg
    path = os.path.join(GOLDEN, name + ".neu")
    H = hostapi.HostHierarchy.from_neu(path, 1)
    pb = PoissonMG(ctx, 0, 0, 0, 1, order, hier=H, coarse_rtol=1e-15)
    assert len(pb.plans) == (3 if name.startswith("cube_mixed") else 1)
    L = mm.read_neu(path)
    u = np.random.default_rng(4).standard_normal(pb.n)
    pb.SOL.put(u)
    pb.assemble()
    Aref, rhs = mm.assemble(L, order, u)
    A = pb.KK[-1].to_scipy()
    assert np.array_equal(A.indptr, Aref.indptr) and np.array_equal(A.indices, Aref.indices)
    assert np.abs(A.data - Aref.data).max() <= RTOL * np.abs(Aref.data).max()
    assert np.abs(pb.RES.get() - rhs).max() <= RTOL * (np.abs(Aref) @ np.abs(u) + np.abs(rhs)).max()
    pb.galerkin(); pb.mg_set_levels(); pb.mg_solve()
    idx = np.nonzero(mm.bdc_flags(L, order) < 1.5)[0]
    Ap = mg.penalty_fast(Aref, idx)
    b = rhs.copy(); b[idx] = 0.0
    x = spla.spsolve(Ap.tocsc(), b)
    assert np.abs(pb.EPS.get() - x).max() <= 1e-10 * np.abs(x).max()
    del pb


@pytest.mark.parametrize("name,order,nl", [("cube_wedge18", "linear", 3), ("cube_wedge18", "biquadratic", 3), ("cube_mixed", "quadratic", 3),
                                           ("cube_mixed", "biquadratic", 2)])
def test_wedge_and_mixed_mesh_vcycle_trace(ctx, name, order, nl):
    """Multigrid on refined wedges / mixed meshes: per-type assembly on the finest level, Galerkin chain by the
    general triple product, penalised level operators and six V-cycles against the oracle (Richardson 0.3)."""
    from femus_b200 import hostapi
    from femus_b200.poisson import PoissonMG
    from oracle import mesh_mixed as mm, mg
    path = os.path.join(GOLDEN, name + ".neu")
    H = hostapi.HostHierarchy.from_neu(path, nl)
    pb = PoissonMG(ctx, 0, 0, 0, nl, order, hier=H, coarse_rtol=1e-15, omega=0.3)
    pb.assemble(); pb.galerkin(); pb.mg_set_levels()
    O = mg.Hierarchy(mm.build_hierarchy(path, nl), order, mesh=mm)
    for l in range(nl):
        got, ref = pb.KK[l].to_scipy(), O.A[l]
        assert np.array_equal(got.indptr, ref.indptr) and np.array_equal(got.indices, ref.indices)
        assert np.abs(got.data - ref.data).max() <= RTOL * np.abs(ref.data).max()
    trace_ref, eps_ref = O.mg_solve_trace(6, omega=0.3)
    trace = []
    for _ in range(6):
        pb.mg_solve()
        trace.append(pb.residual_norm())
    for a, b in zip(trace, trace_ref):
        assert abs(a - b) <= RTOL * trace_ref[0], (trace, trace_ref)
    assert np.abs(pb.EPS.get() - eps_ref).max() <= 1e-11 * np.abs(eps_ref).max()
    del pb
