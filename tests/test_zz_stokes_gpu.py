"""GPU parity of the steady Stokes path (first vertical slice): b2_stokes_assemble against the oracle's
restatement of applications/003_NavierStokes/SteadyStokes/main.cpp:290-598, and the multigrid solve of the system with
velocity-pressure Vanka blocks and a direct coarse solve against the oracle V-cycle.  Written without a GPU at hand: the
kernel's source is verified on the CPU emulator (tests/test_kernel_emulation.py); this file is the gate for the device."""
import os
import numpy as np
import pytest

# never run on a GPU yet: a kernel that hangs must not hang the box (the thread method ends the process)
pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900, method="thread")]
RTOL = 1e-12
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _case(name, nl):
    from femus_b200 import hostapi
    from oracle import fe_hex, mesh_box as mb, mesh_mixed as mm
    if name == "box":
        return hostapi.HostHierarchy(2, 2, 2, nl), mb.build_hierarchy(2, 2, 2, nl), mb, (lambda t, o: fe_hex.tables(o)), "biquadratic"
    path = os.path.join(GOLDEN, name + ".neu")
    return hostapi.HostHierarchy.from_neu(path, nl), mm.build_hierarchy(path, nl), mm, (lambda t, o: mm.FE[t].tables(o)), "quadratic"


@pytest.mark.parametrize("name", ["box", "cube_tet10"])
def test_stokes_assembly_matches_oracle(ctx, name):
    """System matrix (on the multi-variable pattern, bit-exact structure) and residual at a random solution:
    Q2-Q1 hexahedra and P2-P1 tetrahedra."""
    from femus_b200.stokes import StokesMG
    from oracle import stokes, mg
    H, lv, mesh, tables_of, ov = _case(name, 1)
    pb = StokesMG(ctx, H, order_v=ov, IRe=0.37)
    sol = np.random.default_rng(5).standard_normal(pb.n)
    pb.SOL.put(sol)
    pb.assemble()
    Aref, rref = stokes.assemble(lv[-1], mesh, ov, "linear", sol, 0.37, tables_of)
    rp, ci = pb.pattern[-1]
    Aref = mg.on_pattern(Aref, rp, ci)
    A = pb.KK[-1].to_scipy()
    assert np.array_equal(A.indptr, rp) and np.array_equal(A.indices, ci)
    assert np.abs(A.data - Aref.data).max() <= RTOL * np.abs(Aref.data).max()
    assert np.abs(pb.RES.get() - rref).max() <= RTOL * (np.abs(Aref) @ np.abs(sol)).max()
    del pb


@pytest.mark.parametrize("name,schedule,sub", [("box", "colours", "lu"), ("box", "levels", "lu"), ("box", "colours", "ilu")])
def test_stokes_vcycle_trace_with_vanka_blocks(ctx, name, schedule, sub):
    """Channel-like problem (velocity Dirichlet on five boundary sets, unit U on the top one, natural outflow on set 2):
    assembly, Galerkin chain, Vanka smoother (pressure = Schur variable, one element per block), direct coarse solve;
    six V-cycles against the oracle, which converge by four orders of magnitude (sub = "ilu": ILU(0) of the saddle-point
    blocks in system-dof order, velocities first, instead of their exact inverses).  (Hexahedra only: on the shipped
    tetrahedral cube the corner elements have every velocity node on the boundary, so the P2-P1 system is singular.)"""
    from femus_b200.stokes import StokesMG
    from oracle import stokes, mg, system as osys
    H, lv, mesh, tables_of, ov = _case(name, 2)
    fams = [ov] * 3 + ["linear"]
    walls = (1, 3, 4, 5, 6)
    pb = StokesMG(ctx, H, order_v=ov, IRe=1.0, velocity_dirichlet=walls, schedule=schedule, block_sub=sub)
    sol = np.zeros(pb.n)
    sol[osys.bdc(lv[-1], mesh, fams, [(6,), (), (), ()]) < 1.5] = 1.0
    pb.SOL.put(sol)
    pb.assemble(); pb.galerkin(); pb.mg_set_levels()
    A, rhs = stokes.assemble(lv[-1], mesh, ov, "linear", sol, 1.0, tables_of)
    A = mg.on_pattern(A, *pb.pattern[-1])
    blocks = [None] + [pb.asm_index[l].blocks() for l in range(1, pb.nlevels)]
    orders = [None] + [np.argsort(pb.asm_groups[l], kind="stable") for l in range(1, pb.nlevels)]
    O = mg.Hierarchy(lv, None, mesh=osys.SystemMesh(mesh, fams, [walls] * 3 + [()]), A_top=A, rhs=rhs, smoother="asm", asm_blocks=blocks,
                     asm_orders=orders, asm_sub=sub)
    got0 = pb.KK[0].to_scipy()
    assert np.abs(got0.data - O.A[0].data).max() <= 1e-11 * np.abs(O.A[0].data).max()
    trace_ref, eps_ref = O.mg_solve_trace(6, omega=1.0)
    trace = []
    for _ in range(6):
        pb.mg_solve()
        trace.append(pb.residual_norm())
    for a, b in zip(trace, trace_ref):
        assert abs(a - b) <= 1e-10 * trace_ref[0], (trace, trace_ref)
    assert trace[-1] < 1e-4 * trace[0]
    assert np.abs(pb.EPS.get() - eps_ref).max() <= 1e-9 * np.abs(eps_ref).max()
    del pb


@pytest.mark.parametrize("name", ["box", "cube_tet10"])
def test_navier_stokes_assembly_matches_oracle(ctx, name):
    """b2_ns_assemble: residual RES = -aRes and the analytic Newton Jacobian at a random solution against the oracle's
    restatement of 03_navier_stokes.hpp:305-413 (Jacobian checked against finite differences on the CPU)."""
    from femus_b200.stokes import StokesMG
    from oracle import navier_stokes as ons, mg
    H, lv, mesh, tables_of, ov = _case(name, 1)
    pb = StokesMG(ctx, H, order_v=ov, IRe=0.21, equation="navier_stokes")
    sol = 0.5 * np.random.default_rng(6).standard_normal(pb.n)
    pb.SOL.put(sol)
    pb.assemble()
    Aref, rref = ons.assemble(lv[-1], mesh, ov, "linear", sol, 0.21, tables_of)
    Aref = mg.on_pattern(Aref, *pb.pattern[-1])
    A = pb.KK[-1].to_scipy()
    assert np.abs(A.data - Aref.data).max() <= RTOL * np.abs(Aref.data).max()
    assert np.abs(pb.RES.get() - rref).max() <= RTOL * np.abs(rref).max()
    del pb


def test_navier_stokes_newton_iterations(ctx):
    """Channel-like problem at nu = 0.1 on a 2-level box: four Newton iterations, each = assembly of residual and
    Jacobian, Galerkin chain, three V-cycles with Vanka blocks and the direct coarse solve, update; the residual norms
    before every update against the oracle pipeline run with the same blocks, and quadratic-looking decay."""
    from femus_b200.stokes import StokesMG
    from oracle import navier_stokes as ons, mg, system as osys
    H, lv, mesh, tables_of, ov = _case("box", 2)
    fams = [ov] * 3 + ["linear"]
    walls = (1, 3, 4, 5, 6)
    pb = StokesMG(ctx, H, order_v=ov, IRe=0.1, velocity_dirichlet=walls, equation="navier_stokes")
    sol = np.zeros(pb.n)
    sol[osys.bdc(lv[-1], mesh, fams, [(6,), (), (), ()]) < 1.5] = 1.0
    pb.SOL.put(sol)
    blocks = [None] + [pb.asm_index[l].blocks() for l in range(1, pb.nlevels)]
    orders = [None] + [np.argsort(pb.asm_groups[l], kind="stable") for l in range(1, pb.nlevels)]
    smesh = osys.SystemMesh(mesh, fams, [walls] * 3 + [()])
    free = smesh.bdc_flags(lv[-1], None) > 1.1
    got, want = [], []
    for it in range(4):
        got.append(pb.newton_step(ncycles=3))
        A, rhs = ons.assemble(lv[-1], mesh, ov, "linear", sol, 0.1, tables_of)
        want.append(float(np.linalg.norm(np.where(free, rhs, 0.0))))
        O = mg.Hierarchy(lv, None, mesh=smesh, A_top=mg.on_pattern(A, *pb.pattern[-1]), rhs=rhs, smoother="asm", asm_blocks=blocks,
                         asm_orders=orders)
        _, eps = O.mg_solve_trace(3, omega=1.0)
        sol = sol + eps
    for a, b in zip(got, want):
        assert abs(a - b) <= 1e-9 * want[0], (got, want)
    assert got[-1] < 1e-3 * got[0]


def test_cpp_stokes_driver_through_the_adapters():
    """tests/cpp/stokes_driver.cpp: the multi-variable plugin surface (InitPdeSystem, LinearEquationSolverB200Asm with the
    pressure as Schur variable, SetCoarseDirect, MGInit / MGSetLevel / MGSolve) reproduces the oracle V-cycle trace."""
    import re
    import subprocess
    from femus_b200 import build, hostapi
    from oracle import stokes, mg, system as osys, mesh_box as mb, fe_hex
    ncyc = 4
    r = subprocess.run([build.STOKES_DRIVER, "2", "2", "2", "2", str(ncyc)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    res = [float(x) for x in re.findall(r"cycle \d+ residual (\S+)", r.stdout)]
    assert len(res) == ncyc + 1
    H, lv = hostapi.HostHierarchy(2, 2, 2, 2), mb.build_hierarchy(2, 2, 2, 2)
    fams = ["biquadratic"] * 3 + ["linear"]
    walls = (1, 3, 4, 5, 6)
    S = hostapi.SystemOnLevel(H.levels[-1], fams)
    rp, ci = S.sparsity()
    sol = np.zeros(S.n)
    sol[osys.bdc(lv[-1], mb, fams, [(6,), (), (), ()]) < 1.5] = 1.0
    A, rhs = stokes.assemble(lv[-1], mb, "biquadratic", "linear", sol, 1.0, lambda t, o: fe_hex.tables(o))
    ix = hostapi.AsmIndex(H.levels[1], fams, 1, nschur=1)
    grp, gptr, gblocks = hostapi.asm_schedule(rp, ci, ix.overlap_ptr, ix.overlap, "colours")
    smesh = osys.SystemMesh(mb, fams, [walls] * 3 + [()])
    O = mg.Hierarchy(lv, None, mesh=smesh, A_top=mg.on_pattern(A, rp, ci), rhs=rhs, smoother="asm", asm_blocks=[None, ix.blocks()],
                     asm_orders=[None, gblocks])
    trace, eps = O.mg_solve_trace(ncyc, omega=1.0)
    free = smesh.bdc_flags(lv[-1], None) > 1.1
    r0 = float(np.linalg.norm(np.where(free, rhs, 0.0)))
    assert abs(res[0] - r0) <= 1e-12 * r0
    for k in range(ncyc):
        assert abs(res[k + 1] - trace[k]) <= 1e-9 * r0, (k, res[k + 1], trace[k])
    m = re.search(r"vanka blocks (\d+) groups (\d+)", r.stdout)
    assert int(m.group(1)) == ix.nblocks and int(m.group(2)) == len(gptr) - 1


def test_stokes_assembly_on_the_mixed_mesh(ctx):
    """One plan per element type (hexahedra, tetrahedra, wedges) accumulating into one system matrix."""
    from femus_b200 import hostapi
    from femus_b200.stokes import StokesMG
    from oracle import stokes, mg, mesh_mixed as mm
    path = os.path.join(GOLDEN, "cube_mixed.neu")
    pb = StokesMG(ctx, hostapi.HostHierarchy.from_neu(path, 1), IRe=0.5)
    assert len(pb.plans) == 3
    sol = np.random.default_rng(7).standard_normal(pb.n)
    pb.SOL.put(sol)
    pb.assemble()
    Aref, rref = stokes.assemble(mm.read_neu(path), mm, "biquadratic", "linear", sol, 0.5, lambda t, o: mm.FE[t].tables(o))
    Aref = mg.on_pattern(Aref, *pb.pattern[-1])
    A = pb.KK[-1].to_scipy()
    assert np.abs(A.data - Aref.data).max() <= RTOL * np.abs(Aref.data).max()
    assert np.abs(pb.RES.get() - rref).max() <= RTOL * (np.abs(Aref) @ np.abs(sol)).max()
    del pb


def test_navier_stokes_boundary_pressure_term(ctx):
    """Pressure-driven channel: prescribed pressures on the open boundary sets 2 (outflow) and 1 (inflow), no-slip
    elsewhere: the residual at a random solution = volume part + boundary pressure block (03_navier_stokes.hpp:196-300)
    against the oracle."""
    from femus_b200 import hostapi
    from femus_b200.stokes import StokesMG
    from oracle import navier_stokes as ons, mesh_box as mb, fe_hex
    tau = {1: 2.0, 2: -0.5}
    H, lv = hostapi.HostHierarchy(2, 2, 2, 1), mb.build_hierarchy(2, 2, 2, 1)
    pb = StokesMG(ctx, H, IRe=0.3, velocity_dirichlet=(3, 4, 5, 6), equation="navier_stokes", boundary_pressure=tau)
    sol = 0.3 * np.random.default_rng(8).standard_normal(pb.n)
    pb.SOL.put(sol)
    pb.assemble()
    _, rref = ons.assemble(lv[-1], mb, "biquadratic", "linear", sol, 0.3, lambda t, o: fe_hex.tables(o))
    rb = ons.pressure_boundary_rhs(lv[-1], mb, "biquadratic", "linear", tau)
    assert np.abs(rb).max() > 0
    assert np.abs(pb.RES.get() - (rref + rb)).max() <= RTOL * np.abs(rref + rb).max()
    del pb


def test_cpp_driver_newton_iterations_of_navier_stokes():
    """stokes_driver ... ns: b2_ns_assemble through the adapters, one Newton iteration per cycle (three V-cycles each);
    the residual before every update against the oracle Newton-multigrid pipeline."""
    import re
    import subprocess
    from femus_b200 import build, hostapi
    from oracle import navier_stokes as ons, mg, system as osys, mesh_box as mb, fe_hex
    ncyc = 3
    r = subprocess.run([build.STOKES_DRIVER, "2", "2", "2", "2", str(ncyc), "ns"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    res = [float(x) for x in re.findall(r"cycle \d+ residual (\S+)", r.stdout)]
    assert len(res) == ncyc + 1
    H, lv = hostapi.HostHierarchy(2, 2, 2, 2), mb.build_hierarchy(2, 2, 2, 2)
    fams = ["biquadratic"] * 3 + ["linear"]
    walls = (1, 3, 4, 5, 6)
    S = hostapi.SystemOnLevel(H.levels[-1], fams)
    rp, ci = S.sparsity()
    ix = hostapi.AsmIndex(H.levels[1], fams, 1, nschur=1)
    grp, gptr, gblocks = hostapi.asm_schedule(rp, ci, ix.overlap_ptr, ix.overlap, "colours")
    smesh = osys.SystemMesh(mb, fams, [walls] * 3 + [()])
    free = smesh.bdc_flags(lv[-1], None) > 1.1
    sol = np.zeros(S.n)
    sol[osys.bdc(lv[-1], mb, fams, [(6,), (), (), ()]) < 1.5] = 1.0
    want = []
    for it in range(ncyc + 1):
        A, rhs = ons.assemble(lv[-1], mb, "biquadratic", "linear", sol, 0.1, lambda t, o: fe_hex.tables(o))
        want.append(float(np.linalg.norm(np.where(free, rhs, 0.0))))
        if it == ncyc:
            break
        O = mg.Hierarchy(lv, None, mesh=smesh, A_top=mg.on_pattern(A, rp, ci), rhs=rhs, smoother="asm", asm_blocks=[None, ix.blocks()],
                         asm_orders=[None, gblocks])
        sol = sol + O.mg_solve_trace(3, omega=1.0)[1]
    for a, b in zip(res, want):
        assert abs(a - b) <= 1e-9 * want[0], (res, want)
    assert res[-1] < 1e-3 * res[0]


@pytest.mark.parametrize("name,ov", [("box", "biquadratic")])
def test_lid_driven_cavity_with_vanka_blocks_and_pressure_null_space(ctx, name, ov):
    """BASELINE configs[3] as a parity case: steady Navier-Stokes in an ENCLOSED cavity (every velocity Dirichlet, the lid
    moving), two levels, velocity-pressure Vanka blocks as level smoother, Q2-Q1 hexahedra.  (Not on the shipped
    tetrahedral cube: with every wall Dirichlet its 105-element coarse level is discretely singular even with the pressure
    pinned -- a Taylor-Hood mesh with tetrahedra that have no interior vertex: smallest singular value 3e-17 of 2.1 -- so
    the reference's LU on level 0 has nothing to solve there; Tet10 assembly and V-cycles are covered by the tests above
    and tests/test_tet_gpu.py.)  The pressure is defined up to a constant:
    MultiLevelSolution::FixSolutionAtOnePoint("P") pins its first dof on the coarsest level
    (MultiLevelSolution.cpp:826-830) and removes the constant pressure as null space of the operators above
    (RemoveNullSpace, LinearEquationSolverPetsc.cpp:357-414: right-hand side and every preconditioned residual of the
    level smoother projected).  Newton iterations on the device against the oracle pipeline with the same blocks."""
    from femus_b200.stokes import StokesMG
    from oracle import navier_stokes as ons, mg, system as osys
    H, lv, mesh, tables_of, _ = _case(name, 2)
    fams = [ov] * 3 + ["linear"]
    walls = (1, 2, 3, 4, 5, 6)
    nu = 0.5
    pb = StokesMG(ctx, H, order_v=ov, IRe=nu, velocity_dirichlet=walls, equation="navier_stokes", fix_pressure_at_one_point=True)
    # the lid: unit x-velocity on the dofs of boundary set 6 that are on no other wall
    lid = osys.bdc(lv[-1], mesh, fams, [(6,), (), (), ()]) < 1.5
    other = osys.bdc(lv[-1], mesh, fams, [(1, 2, 3, 4, 5), (), (), ()]) < 1.5
    sol = np.zeros(pb.n)
    sol[lid & ~other] = 1.0
    pb.SOL.put(sol)
    blocks = [None] + [pb.asm_index[l].blocks() for l in range(1, pb.nlevels)]
    orders = [None] + [np.argsort(pb.asm_groups[l], kind="stable") for l in range(1, pb.nlevels)]

    class Cavity(osys.SystemMesh):           # FixSolutionAtOnePoint: the first pressure dof of level 0 is a Dirichlet row
        def bdc_flags(self, L, order, dirichlet_faces=None):
            b = osys.SystemMesh.bdc_flags(self, L, order, dirichlet_faces)
            if L is lv[0]:
                b[int(pb.sys[0].offsets[3, 0])] = 0.0
            return b
    smesh = Cavity(mesh, fams, [walls] * 3 + [()])
    assert np.array_equal(np.nonzero(smesh.bdc_flags(lv[0], None) < 1.5)[0], pb.bdc_idx[0])
    nulls = [None] + [pb.nullspace_base(l) for l in range(1, pb.nlevels)]
    free = smesh.bdc_flags(lv[-1], None) > 1.1
    got, want = [], []
    for it in range(4):
        got.append(pb.newton_step(ncycles=4))
        A, rhs = ons.assemble(lv[-1], mesh, ov, "linear", sol, nu, tables_of)
        want.append(float(np.linalg.norm(np.where(free, rhs, 0.0))))
        O = mg.Hierarchy(lv, None, mesh=smesh, A_top=mg.on_pattern(A, *pb.pattern[-1]), rhs=rhs, smoother="asm", asm_blocks=blocks,
                         asm_orders=orders, nullspace=nulls)
        # the constant pressure IS the null space of the penalised finest operator
        assert np.abs(O.A[-1] @ nulls[-1]).max() <= 1e-12 * np.abs(O.A[-1].data).max()
        _, eps = O.mg_solve_trace(4, omega=1.0)
        sol = sol + eps
    for a, b in zip(got, want):
        assert abs(a - b) <= 1e-8 * want[0], (got, want)
    assert got[-1] < 1e-2 * got[0], got
    del pb


@pytest.mark.parametrize("eq", ["stokes", "navier_stokes"])
def test_assembly_matches_the_reference_output(ctx, eq):
    """The device kernels against REFERENCE OUTPUT, without the oracle in between: tests/golden/ref_stokes_*.npz hold the
    matrix and the residual the reference's own classes assembled for a Taylor-Hood system on a 2 x 1 x 1 box refined once
    -- AssembleMatrixResSteadyStokes of applications/003_NavierStokes/SteadyStokes/main.cpp, resp. the library routine
    AssembleNavierStokes_AD (Jacobian recorded by adept, boundary pressure 0.75 on boundary set 2) -- at the fields the
    fixture carries (tests/cpp/ref_stokes.cpp, tests/golden/make_ref_stokes_golden.py)."""
    import scipy.sparse as sp
    from femus_b200 import hostapi
    from femus_b200.stokes import StokesMG
    ns = eq == "navier_stokes"
    g = np.load(os.path.join(GOLDEN, "ref_stokes_ns_box211_q2q1_2lev.npz" if ns else "ref_stokes_box211_q2q1_2lev.npz"))
    box, nl = tuple(int(v) for v in g["box"]), int(g["nlevels"])
    top = nl - 1
    pb = StokesMG(ctx, hostapi.HostHierarchy(*box, nl), IRe=float(g["IReynolds"]), equation=eq, boundary_pressure={2: 0.75} if ns else None)
    sol = np.concatenate([g[f"L{top}_SOL_{v}"] for v in "UVWP"])          # one rank: system rows = [variable][dof]
    assert sol.shape[0] == pb.n
    pb.SOL.put(sol)
    pb.assemble()
    Ar = sp.csr_matrix((g[f"L{top}_KK_val"], g[f"L{top}_KK_col"], g[f"L{top}_KK_rowptr"]), shape=(pb.n, pb.n))
    A = pb.KK[-1].to_scipy()             # on the full coupling pattern: the entries the reference's callback never touches stay zero
    assert abs(A - Ar).max() <= RTOL * np.abs(Ar.data).max()
    res = g[f"L{top}_RES"]
    assert np.abs(pb.RES.get() - res).max() <= RTOL * np.abs(res).max()
    del pb
