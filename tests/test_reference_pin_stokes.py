"""CPU: the multi-variable oracles (oracle/system.py, oracle/stokes.py) and the product's host layer (SystemLayout.hpp
through hostapi.SystemOnLevel) against REFERENCE OUTPUT: tests/golden/ref_stokes_*.npz hold what the reference's own
classes produced for a Taylor-Hood system U, V, W (SECOND) + P (FIRST) on a HEX27 box with the assembly callback of
applications/003_NavierStokes/SteadyStokes/main.cpp (AssembleMatrixResSteadyStokes, :290-598) compiled in place
(tests/cpp/ref_stokes.cpp on the host backend of oracle/ref_build; tests/golden/make_ref_stokes_golden.py).  Integers
(system dofs of every variable, KKoffset, sparsity counts and pattern, prolongator structure, Dirichlet flags per
variable) bit-exact; assembled matrix, residual, prolongator and Galerkin operator to the tolerances stated below."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
ORDERS = ["biquadratic"] * 3 + ["linear"]
# SteadyStokes/main.cpp:201-283 (SetBoundaryCondition) on the six boundary sets of a box: U and V natural on set 2
# (outflow), W Dirichlet everywhere, P natural on sets 1-4 -- and Dirichlet 0 on 5, 6, which the function does not name
DIRICHLET = [(1, 3, 4, 5, 6), (1, 3, 4, 5, 6), (1, 2, 3, 4, 5, 6), (5, 6)]
# the initial fields of tests/cpp/ref_stokes.cpp
INIT = [lambda x: 0.3 * x[1] * (1.0 - x[2]) + 0.1 * x[0] * x[0], lambda x: -0.2 * x[0] * x[2] + 0.05 * x[1],
        lambda x: 0.15 * x[0] * x[1] - 0.1 * x[2] * x[2], lambda x: 1.0 + 0.5 * x[0] - 0.25 * x[1] + 0.125 * x[2]]


def csr(g, key):
    return sp.csr_matrix((g[f"{key}_val"], g[f"{key}_col"], g[f"{key}_rowptr"]), shape=tuple(g[f"{key}_shape"]))


@pytest.mark.parametrize("name", ["box211_q2q1_2lev"])
def test_taylor_hood_system_matches_the_reference(name):
    from femus_b200 import hostapi
    from oracle import asm, fe_hex, mesh_box as mb, stokes, system as osys
    g = np.load(os.path.join(GOLDEN, f"ref_stokes_{name}.npz"))
    box, nl, IRe = tuple(int(v) for v in g["box"]), int(g["nlevels"]), float(g["IReynolds"])
    lv = mb.build_hierarchy(*box, nl)
    H = hostapi.HostHierarchy(*box, nl)
    nve = [27, 27, 27, 8]
    fi = [mb.FAMILY[o] for o in ORDERS]
    A_top = None
    for l in range(nl):
        nel, nnode, dim, nvar, soltype, nprocs = (int(v) for v in g[f"L{l}_info"])
        assert (nel, nnode, dim, nvar, nprocs) == (lv[l].nel, lv[l].nnode, 3, 4, 1)
        assert np.array_equal(g[f"L{l}_conn"].reshape(nel, 27), lv[l].conn)
        sysdof = g[f"L{l}_sysdof"].reshape(4, nel, 27)
        S = hostapi.SystemOnLevel(H.levels[l], ORDERS)
        # ---- numbering of the system: GetSystemDof per variable and element, KKoffset
        d_or = osys.elem_system_dofs(lv[l], mb, ORDERS)
        d_host = S.elem_dofs()
        for k in range(4):
            for e in range(nel):
                assert np.array_equal(sysdof[k, e, :nve[k]], d_or[k][e]), f"oracle: level {l} variable {k} element {e}"
                assert np.array_equal(sysdof[k, e, :nve[k]], d_host[e, k, :nve[k]]), f"host layer: level {l} variable {k} element {e}"
            assert np.all(sysdof[k, :, nve[k]:] == -1) and np.all(d_host[:, k, nve[k]:] == -1)
        kk = g[f"L{l}_KKoffset"].reshape(5, 1)
        assert np.array_equal(kk, asm.kk_offsets(lv[l], fi)) and np.array_equal(kk, hostapi.system_offsets(H.levels[l], ORDERS))
        n = int(kk[-1, -1])
        assert S.n == n
        # ---- sparsity: the counts GetSparsityPatternSize hands to init() couple every variable with every variable of an
        # element (LinearEquation.cpp:407-548); the callback then touches only B[k][k], B[k][p], B[p][k] and the zero
        # block B[p][p] (the pattern of the assembled matrix is compared with the oracle's assembly below)
        assert not g[f"L{l}_n_oz"].any()
        for who, (rp, ci) in (("oracle", osys.sparsity(lv[l], mb, ORDERS)), ("host layer", S.sparsity())):
            assert np.array_equal(g[f"L{l}_n_nz"], np.diff(rp)), f"{who}: level {l} sparsity counts"
            ref_pat = sp.csr_matrix((np.ones(len(g[f"L{l}_KK_col"]), dtype=np.int8), g[f"L{l}_KK_col"], g[f"L{l}_KK_rowptr"]), shape=(n, n))
            full = sp.csr_matrix((np.ones(len(ci), dtype=np.int8), ci, rp), shape=(n, n))
            assert (ref_pat - ref_pat.multiply(full)).nnz == 0, f"{who}: level {l}: an assembled entry outside the pattern"
        # ---- Dirichlet flags per variable (GenerateBdc with the application's own SetBoundaryCondition)
        bdc_ref = g[f"L{l}_Bdc"]
        assert np.array_equal(bdc_ref, osys.bdc(lv[l], mb, ORDERS, DIRICHLET)), f"oracle: level {l} Bdc"
        assert np.array_equal(bdc_ref, S.bdc(DIRICHLET)), f"host layer: level {l} Bdc"
        assert np.array_equal(np.nonzero(bdc_ref < 1.5)[0], g[f"L{l}_bdcIndex"])
        # ---- prolongator of the system, variable by variable, Dirichlet rows / columns zeroed in the pattern
        if l > 0:
            P = csr(g, f"L{l}_PP")
            Po = mb.zero_dirichlet(osys.prolongator(lv[l - 1], lv[l], mb, ORDERS), bdc_ref, g[f"L{l - 1}_Bdc"]).tocsr()
            assert abs(P - Po).max() <= 1e-15
            rph, cih, vh, shp = S.prolongator()
            assert tuple(shp) == P.shape and np.array_equal(rph, P.indptr) and np.array_equal(cih, P.indices), "host layer: prolongator structure"
            # the host layer hands out the prolongator BEFORE ZeroInterpolatorDirichletNodes (the device zeroes rows / columns):
            # its values where the reference kept them, and the reference's zeros exactly on Dirichlet rows / columns
            rows = np.repeat(np.arange(P.shape[0]), np.diff(rph))
            zeroed = (bdc_ref[rows] < 1.5) | (g[f"L{l - 1}_Bdc"][cih] < 1.5)
            assert np.abs(vh[~zeroed] - P.data[~zeroed]).max() <= 1e-15 and not P.data[zeroed].any()
    # ---- the assembled system of the finest level: AssembleMatrixResSteadyStokes at the initial fields
    top = nl - 1
    L = lv[top]
    d = osys.elem_system_dofs(L, mb, ORDERS)
    n = int(g[f"L{top}_KKoffset"][-1])
    # the fields the callback read (dumped by the driver in solution-dof numbering per variable): the initial functions of
    # ref_stokes.cpp, except on Dirichlet nodes, where GenerateBdc wrote the boundary values of SetBoundaryCondition
    sol = np.concatenate([g[f"L{top}_SOL_{v}"] for v in "UVWP"])          # one rank: system rows = [variable][dof]
    assert sol.shape[0] == n
    init = np.zeros(n)
    for k in range(4):
        for e in range(L.nel):
            init[d[k][e]] = INIT[k](L.xyz[:, L.conn[e, :nve[k]]])
    free = g[f"L{top}_Bdc"] > 1.5
    assert np.abs(sol[free] - init[free]).max() <= 1e-15
    Ao, rhs = stokes.assemble(L, mb, "biquadratic", "linear", sol, IRe, lambda t, o: fe_hex.tables(o))
    Ar = csr(g, f"L{top}_KK")
    assert np.array_equal(Ar.indptr, Ao.indptr) and np.array_equal(Ar.indices, Ao.indices)
    assert np.abs(Ar.data - Ao.data).max() <= 1e-13 * np.abs(Ao.data).max()
    res_ref = g[f"L{top}_RES"]
    assert np.abs(rhs - res_ref).max() <= 1e-13 * np.abs(res_ref).max()
    # ---- Galerkin operator of the level below: KK_{l-1} = PP^T KK_l PP on the un-penalised matrices (LinearImplicitSystem.cpp:347-370)
    if nl > 1:
        P = csr(g, f"L{top}_PP")
        Ac = csr(g, f"L{top - 1}_KK")
        G = (P.T @ Ao @ P).tocsr()
        assert abs(Ac - G).max() <= 1e-12 * abs(G).max()


def test_navier_stokes_routine_matches_the_reference():
    """The library routine femus::AssembleNavierStokes_AD (03_navier_stokes.hpp:21-413, nu = 1) run by the reference
    itself on the same Taylor-Hood system: the residual of the Galerkin form, the Jacobian that adept recorded -- against
    the oracle's ANALYTIC Newton Jacobian (what ns_kernel implements) -- and the boundary pressure block (prescribed
    pressure 0.75 on boundary set 2, the only set whose normal velocity component is not Dirichlet)."""
    from oracle import fe_hex, mesh_box as mb, navier_stokes as ons, system as osys
    g = np.load(os.path.join(GOLDEN, "ref_stokes_ns_box211_q2q1_2lev.npz"))
    box, nl, nu = tuple(int(v) for v in g["box"]), int(g["nlevels"]), float(g["IReynolds"])
    assert nu == 1.0
    lv = mb.build_hierarchy(*box, nl)
    top = nl - 1
    L = lv[top]
    n = int(g[f"L{top}_KKoffset"][-1])
    sol = np.concatenate([g[f"L{top}_SOL_{v}"] for v in "UVWP"])
    assert sol.shape[0] == n and np.array_equal(g[f"L{top}_Bdc"], osys.bdc(L, mb, ORDERS, DIRICHLET))
    Ao, rhs = ons.assemble(L, mb, "biquadratic", "linear", sol, nu, lambda t, o: fe_hex.tables(o))
    rhs = rhs + ons.pressure_boundary_rhs(L, mb, "biquadratic", "linear", {2: 0.75})
    Ar = csr(g, f"L{top}_KK")
    assert np.array_equal(Ar.indptr, Ao.indptr) and np.array_equal(Ar.indices, Ao.indices)      # every variable couples with every variable
    assert np.array_equal(g[f"L{top}_n_nz"], np.diff(Ar.indptr))
    assert np.abs(Ar.data - Ao.data).max() <= 1e-12 * np.abs(Ao.data).max()
    res_ref = g[f"L{top}_RES"]
    assert np.abs(rhs - res_ref).max() <= 1e-12 * np.abs(res_ref).max()
    P = csr(g, f"L{top}_PP")
    G = (P.T @ Ao @ P).tocsr()
    assert abs(csr(g, f"L{top - 1}_KK") - G).max() <= 1e-12 * abs(G).max()


def test_pressure_pinned_at_one_point_matches_the_reference():
    """MultiLevelSolution::FixSolutionAtOnePoint("P") on an enclosed flow (every velocity Dirichlet, pressure natural), run by
    the reference itself on three levels: the FIRST pressure dof of the COARSEST level becomes a Dirichlet row (flag 0),
    nothing changes on the levels above (MultiLevelSolution.cpp:826-830) -- the rule StokesMG(fix_pressure_at_one_point)
    applies (femus_b200/stokes.py), with the constant removed as the operators' null space above."""
    from femus_b200 import hostapi
    from oracle import mesh_box as mb, system as osys
    g = np.load(os.path.join(GOLDEN, "ref_stokes_ns_fix_box211_q2q1_3lev_bdc.npz"))
    box, nl = tuple(int(v) for v in g["box"]), int(g["nlevels"])
    enclosed = [(1, 2, 3, 4, 5, 6)] * 3 + [()]
    lv, H = mb.build_hierarchy(*box, nl), hostapi.HostHierarchy(*box, nl)
    for l in range(nl):
        S = hostapi.SystemOnLevel(H.levels[l], ORDERS)
        want = S.bdc(enclosed)
        assert np.array_equal(want, osys.bdc(lv[l], mb, ORDERS, enclosed))
        if l == 0:
            p0 = int(g["L0_KKoffset"][3])
            assert p0 == int(S.offsets[3, 0]) and want[p0] == 2.0
            want[p0] = 0.0                      # what stokes.py does: self.bdc[0][offsets[3, 0]] = 0.0
        assert np.array_equal(g[f"L{l}_Bdc"], want), f"level {l}"
        assert np.array_equal(g[f"L{l}_bdcIndex"], np.nonzero(want < 1.5)[0])


def test_newton_iteration_matches_the_reference():
    """The reference's OWN Newton loop (NonLinearImplicitSystem::MGsolve, NonLinearImplicitSystem.cpp:157-361) around its
    library routine, on one level, where every linear solve is the host backend's exact LU: the norms of the update and of
    the solution it prints per variable after every iteration (:137; 7 digits) against the oracle's loop -- assemble
    residual and ANALYTIC Jacobian at Sol, Dirichlet rows to the identity with zero residual (SetPenalty,
    ZerosBoundaryResiduals), exact solve, Sol += Eps.  Quadratic convergence: 0.68 -> 3.8e-3 -> 1.5e-7 -> round-off."""
    import scipy.sparse.linalg as spla
    from oracle import fe_hex, mesh_box as mb, navier_stokes as ons, system as osys
    g = np.load(os.path.join(GOLDEN, "ref_stokes_ns_newton_box221_q2q1_1lev.npz"))
    box, nl = tuple(int(v) for v in g["box"]), int(g["nlevels"])
    assert nl == 1
    L = mb.build_hierarchy(*box, 1)[0]
    kk = g["L0_KKoffset"]
    n = int(kk[-1])
    bdc = g["L0_Bdc"]
    assert np.array_equal(bdc, osys.bdc(L, mb, ORDERS, DIRICHLET))
    sol = np.concatenate([g[f"L0_SOL_{v}"] for v in "UVWP"])
    fixed = bdc < 1.5
    eps_ref, sol_ref = g["newton_eps_l2"], g["newton_sol_l2"]
    assert eps_ref.shape[0] >= 4
    for it in range(eps_ref.shape[0]):
        A, rhs = ons.assemble(L, mb, "biquadratic", "linear", sol, 1.0, lambda t, o: fe_hex.tables(o))
        rhs = rhs + ons.pressure_boundary_rhs(L, mb, "biquadratic", "linear", {2: 0.75})
        A = A.tolil()
        for i in np.nonzero(fixed)[0]:
            A.rows[i], A.data[i] = [int(i)], [1.0]
        rhs[fixed] = 0.0
        eps = spla.spsolve(A.tocsc(), rhs)
        sol = sol + eps
        for k in range(4):
            e, s = np.linalg.norm(eps[kk[k]:kk[k + 1]]), np.linalg.norm(sol[kk[k]:kk[k + 1]])
            if eps_ref[it, k] > 1e-10:          # the last iteration is round-off on both sides
                assert abs(e - eps_ref[it, k]) <= 2e-6 * eps_ref[it, k], (it, k, e, eps_ref[it, k])
            else:
                assert e < 1e-10
            assert abs(s - sol_ref[it, k]) <= 2e-6 * sol_ref[it, k], (it, k, s, sol_ref[it, k])
