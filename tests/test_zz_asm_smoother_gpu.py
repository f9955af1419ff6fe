"""GPU parity of the element-block (ASM / Vanka) smoother (b2_schwarz.cu, smoother kind 2 of b2_mg.cu) against the
oracle's restatement of PCASM (basic, multiplicative, exact block solves; oracle/asm.py), through the C ABI.
Reference: LinearEquationSolverPetscAsm.cpp:91-340, MeshASMPartitioning.cpp:89-148, 001_Poisson main.cpp:234-250.
Bar: preconditioner application and V-cycle residual traces to 1e-12 relative (block inverses by Gauss-Jordan on the
GPU, LU in the oracle)."""
import os
import numpy as np
import pytest

# a kernel that hangs must not hang the box (the thread method ends the process)
pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900, method="thread")]
RTOL = 1e-12
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _oracle(pb, levels, order, mesh=None, **kw):
    from oracle import mg
    blocks = [None] + [pb.asm_index[l].blocks() for l in range(1, pb.nlevels)]
    orders = [None] + [np.argsort(pb.asm_groups[l], kind="stable") for l in range(1, pb.nlevels)]
    return mg.Hierarchy(levels, order, smoother="asm", asm_blocks=blocks, asm_orders=orders, mesh=mesh, **kw)


@pytest.mark.parametrize("schedule", ["levels", "colours"])
@pytest.mark.parametrize("order,shape,nb", [("linear", (2, 3, 2), 8), ("biquadratic", (2, 2, 2), 8), ("biquadratic", (2, 3, 2), 5),
                                            ("quadratic", (2, 2, 2), 8)])
def test_block_preconditioner_matches_oracle(ctx, schedule, order, shape, nb):
    """y = M^-1 r on the penalised finest operator of a 3-level box hierarchy: blocks of nb elements in the
    reference's order ("levels": its sequential sweep exactly) and in coloured order, ragged last block (nb = 5)."""
    from femus_b200.poisson import PoissonMG
    from oracle import mesh_box as mb
    pb = PoissonMG(ctx, *shape, 3, order, smoother="asm", asm_block_elems=nb, asm_schedule=schedule, fused=False)
    pb.assemble(); pb.galerkin(); pb.mg_set_levels()
    O = _oracle(pb, mb.build_hierarchy(*shape, 3), order)
    rng = np.random.default_rng(3)
    for l in (1, 2):
        n = pb.ndofs[l]
        r = rng.standard_normal(n)
        R, Y = ctx.vector(r), ctx.vector(n)
        pb.schwarz[l].apply(R, Y)
        want = O.asm[l].apply(r)
        assert np.abs(Y.get() - want).max() <= RTOL * np.abs(want).max()
        if schedule == "colours":
            assert pb.schwarz[l].ngroups <= 27
        assert pb.schwarz[l].nbytes == 8 * sum(len(b) ** 2 for b in pb.asm_index[l].blocks())
    del pb


@pytest.mark.parametrize("schedule,order,nl", [("levels", "linear", 3), ("colours", "linear", 3), ("colours", "biquadratic", 3),
                                               ("levels", "biquadratic", 2)])
def test_vcycle_trace_with_block_smoother(ctx, schedule, order, nl):
    """001_Poisson with "smoother": "asm" (Richardson around the block preconditioner, here scale 1): four V-cycles
    against the oracle, and the smoother's pay-off against Richardson + Jacobi on the same problem."""
    from femus_b200.poisson import PoissonMG
    from oracle import mesh_box as mb, mg
    pb = PoissonMG(ctx, 2, 2, 2, nl, order, smoother="asm", asm_block_elems=8, asm_schedule=schedule, omega=1.0, coarse_rtol=1e-15)
    pb.assemble(); pb.galerkin(); pb.mg_set_levels()
    lv = mb.build_hierarchy(2, 2, 2, nl)
    O = _oracle(pb, lv, order)
    trace_ref, eps_ref = O.mg_solve_trace(4, omega=1.0)
    trace = []
    for _ in range(4):
        pb.mg_solve()
        trace.append(pb.residual_norm())
    for a, b in zip(trace, trace_ref):
        assert abs(a - b) <= RTOL * trace_ref[0], (trace, trace_ref)
    assert np.abs(pb.EPS.get() - eps_ref).max() <= 1e-11 * np.abs(eps_ref).max()
    jac, _ = mg.Hierarchy(lv, order).mg_solve_trace(4)
    assert trace[-1] < 1e-2 * jac[-1]
    del pb


@pytest.mark.parametrize("name,order", [("cube_tet10", "quadratic"), ("cube_mixed_3groups", "biquadratic")])
def test_block_smoother_on_unstructured_meshes(ctx, name, order):
    """Tetrahedra and the mixed mesh with two material classes (solid blocks first): ragged blocks of 3 elements,
    coloured sweep, V-cycle trace against the oracle."""
    from femus_b200 import hostapi
    from femus_b200.poisson import PoissonMG
    from oracle import mesh_mixed as mm
    path = os.path.join(GOLDEN, name + ".neu")
    H = hostapi.HostHierarchy.from_neu(path, 2)
    pb = PoissonMG(ctx, 0, 0, 0, 2, order, hier=H, smoother="asm", asm_block_elems=3, omega=1.0, coarse_rtol=1e-15)
    pb.assemble(); pb.galerkin(); pb.mg_set_levels()
    O = _oracle(pb, mm.build_hierarchy(path, 2), order, mesh=mm)
    trace_ref, eps_ref = O.mg_solve_trace(4, omega=1.0)
    trace = []
    for _ in range(4):
        pb.mg_solve()
        trace.append(pb.residual_norm())
    for a, b in zip(trace, trace_ref):
        assert abs(a - b) <= RTOL * trace_ref[0], (trace, trace_ref)
    assert np.abs(pb.EPS.get() - eps_ref).max() <= 1e-11 * np.abs(eps_ref).max()
    del pb


@pytest.mark.parametrize("sub", ["ssor", "ilu"])
@pytest.mark.parametrize("order,nb,schedule", [("linear", 8, "colours"), ("biquadratic", 8, "levels"), ("biquadratic", 4096, "colours")])
def test_vcycle_trace_with_ssor_block_solves(ctx, order, nb, schedule, sub):
    """The application's own sub-preconditioner: one SSOR iteration per block (SetPreconditionerFineGrids(SOR_PRECOND),
    main.cpp:242).  nb = 4096 = SetElementBlockNumber(4) of main.cpp:249: on these small levels ONE block holds every
    element and the smoother is Richardson + SOR, what FEMuS_DEFAULT configures too.  sub = "ilu": ILU(0) block solves
    (ILU_PRECOND, the setting of most of the reference's applications; one block = Richardson + ILU(0))."""
    from femus_b200.poisson import PoissonMG
    from oracle import mesh_box as mb
    pb = PoissonMG(ctx, 2, 2, 2, 3, order, smoother="asm", asm_block_elems=nb, asm_schedule=schedule, asm_sub=sub, omega=1.0,
                   coarse_rtol=1e-15)
    pb.assemble(); pb.galerkin(); pb.mg_set_levels()
    O = _oracle(pb, mb.build_hierarchy(2, 2, 2, 3), order, asm_sub=sub)
    if nb == 4096:
        assert pb.asm_index[2].nblocks == 1 and (pb.schwarz[2].nbytes == 0 if sub == "ssor" else pb.schwarz[2].nbytes == 8 * pb.KK[2].nnz)
    trace_ref, eps_ref = O.mg_solve_trace(4, omega=1.0)
    trace = []
    for _ in range(4):
        pb.mg_solve()
        trace.append(pb.residual_norm())
    for a, b in zip(trace, trace_ref):
        assert abs(a - b) <= RTOL * trace_ref[0], (trace, trace_ref)
    assert np.abs(pb.EPS.get() - eps_ref).max() <= 1e-11 * np.abs(eps_ref).max()
    del pb


def test_vanka_blocks_of_a_saddle_point_system(ctx):
    """b2_schwarz on velocity-pressure Vanka blocks (host index sets with one Schur variable: blocks of up to 3 x 729 + 27 dofs)
    of a synthetic Stokes matrix in the reference's system numbering: application against the oracle."""
    from femus_b200 import capi, hostapi
    from oracle import asm, mesh_box as mb
    from tests import saddle_point as spt
    L = mb.build_hierarchy(2, 2, 2, 2)[1]
    A = spt.stokes_matrix(L)
    H = hostapi.HostHierarchy(2, 2, 2, 2)
    ix = hostapi.AsmIndex(H.levels[1], spt.FAMILIES, 8, nschur=1)
    grp, gptr, gblocks = hostapi.asm_schedule(A.indptr, A.indices, ix.overlap_ptr, ix.overlap, "colours")
    dA = ctx.csr_from_scipy(A)
    S = capi.Schwarz(ctx, dA, ix.overlap_ptr, ix.overlap, gptr, gblocks)
    S.setup()
    r = np.random.default_rng(9).standard_normal(A.shape[0])
    R, Y = ctx.vector(r), ctx.vector(A.shape[0])
    S.apply(R, Y)
    want = asm.BlockSmoother(A, ix.blocks(), order=gblocks).apply(r)
    assert np.abs(Y.get() - want).max() <= 1e-10 * np.abs(want).max()


@pytest.mark.parametrize("pc,order,npre", [("jacobi", "biquadratic", 1), ("jacobi", "linear", 3), ("ilu", "linear", 2), ("ilu", "biquadratic", 1),
                                           ("asm8", "linear", 2)])
def test_vcycle_trace_with_gmres_level_solver(ctx, pc, order, npre):
    """KSPGMRES as level solver (left-preconditioned, npre = npost iterations per smoothing call) around Jacobi, around
    ILU(0) of the whole level -- the reference's DEFAULT level solver, GMRES + ILU_PRECOND -- and around 8-element
    blocks with exact solves; four V-cycles against the oracle."""
    from femus_b200.poisson import PoissonMG
    from oracle import mesh_box as mb, mg
    kw = {} if pc == "jacobi" else dict(smoother="asm", asm_block_elems=10 ** 6 if pc == "ilu" else 8, asm_sub="ilu" if pc == "ilu" else "lu")
    pb = PoissonMG(ctx, 2, 2, 2, 3, order, ksp="gmres", npre=npre, npost=npre, coarse_rtol=1e-15, **kw)
    pb.assemble(); pb.galerkin(); pb.mg_set_levels()
    lv = mb.build_hierarchy(2, 2, 2, 3)
    if pc == "jacobi":
        O = mg.Hierarchy(lv, order, ksp="gmres")
    else:
        O = _oracle(pb, lv, order, asm_sub=kw["asm_sub"], ksp="gmres")
    trace_ref, eps_ref = O.mg_solve_trace(4, npre=npre, npost=npre)
    trace = []
    for _ in range(4):
        pb.mg_solve()
        trace.append(pb.residual_norm())
    for a, b in zip(trace, trace_ref):
        assert abs(a - b) <= 1e-10 * trace_ref[0], (trace, trace_ref)
    assert np.abs(pb.EPS.get() - eps_ref).max() <= 1e-9 * np.abs(eps_ref).max()
    del pb


@pytest.mark.parametrize("sub,nb", [("ilu", 10 ** 6), ("ssor", 10 ** 6), ("ilu", 8), ("ssor", 8)])
def test_level_scheduled_rows_equal_the_one_warp_walk(ctx, sub, nb):
    """b2_schwarz_set_row_levels on the device: bit-for-bit the result of the one-warp walk ON THE SAME OPERATOR, on
    8-element blocks and on ONE block holding the whole level (Richardson + ILU(0) / SOR of FEMuS_DEFAULT), and both
    against the oracle's block solve.  (The operator is assembled once: two assemblies differ in the last bits -- fp64
    atomics commit in scheduling order -- which is what round 1's version of this test actually measured.)"""
    from femus_b200.poisson import PoissonMG
    from oracle import mesh_box as mb
    pb = PoissonMG(ctx, 2, 2, 2, 3, "biquadratic", smoother="asm", asm_block_elems=nb, asm_sub=sub)
    pb.assemble(); pb.galerkin(); pb.mg_set_levels()
    O = _oracle(pb, mb.build_hierarchy(2, 2, 2, 3), "biquadratic", asm_sub=sub)
    for l in (1, 2):
        n = pb.ndofs[l]
        r = np.sin(np.arange(n) * 0.37)
        R, Y = ctx.vector(r), ctx.vector(n)
        S = pb.schwarz[l]
        ys = []
        for lev in (False, True, False):
            S.set_row_levels(lev)
            S.setup()                       # numeric phase again (ILU: the factor through the other row schedule)
            S.apply(R, Y)
            ys.append(Y.get())
            if lev:
                assert 1 < S.row_levels <= max(len(b) for b in pb.asm_index[l].blocks())
        assert np.array_equal(ys[0], ys[2])             # the walk itself is deterministic
        assert np.array_equal(ys[0], ys[1])             # and the level schedule reproduces it bit for bit
        want = O.asm[l].apply(r)
        assert np.abs(ys[1] - want).max() <= RTOL * np.abs(want).max()
    del pb


@pytest.mark.parametrize("sub", ["ssor", "ilu"])
def test_vcycle_trace_with_level_scheduled_rows(ctx, sub):
    """asm_row_levels=True through the V-cycle (one block per level = Richardson + SOR / ILU(0), FEMuS_DEFAULT): four
    cycles against the oracle."""
    from femus_b200.poisson import PoissonMG
    from oracle import mesh_box as mb
    pb = PoissonMG(ctx, 2, 2, 2, 3, "biquadratic", smoother="asm", asm_block_elems="all", asm_sub=sub, asm_row_levels=True, omega=1.0,
                   coarse_rtol=1e-15)
    pb.assemble(); pb.galerkin(); pb.mg_set_levels()
    O = _oracle(pb, mb.build_hierarchy(2, 2, 2, 3), "biquadratic", asm_sub=sub)
    trace_ref, eps_ref = O.mg_solve_trace(4, omega=1.0)
    trace = []
    for _ in range(4):
        pb.mg_solve()
        trace.append(pb.residual_norm())
    for a, b in zip(trace, trace_ref):
        assert abs(a - b) <= RTOL * trace_ref[0], (trace, trace_ref)
    assert np.abs(pb.EPS.get() - eps_ref).max() <= 1e-11 * np.abs(eps_ref).max()
    del pb


def test_block_smoother_fails_loudly(ctx):
    from femus_b200 import capi, hostapi
    H = hostapi.HostHierarchy(2, 2, 2, 2)
    L = H.levels[1]
    A = capi.Csr.from_elements(ctx, L.ndofs("linear"), L.system_dofs("linear"))
    ix = hostapi.AsmIndex(L, "linear", 8)
    rp, ci = L.sparsity("linear")
    grp, gptr, gblocks = hostapi.asm_schedule(rp, ci, ix.overlap_ptr, ix.overlap, "colours")
    bad = gblocks.copy()
    bad[1] = bad[0]
    with pytest.raises(RuntimeError):
        capi.Schwarz(ctx, A, ix.overlap_ptr, ix.overlap, gptr, bad)                 # a block listed twice
    unsorted = ix.overlap.copy()
    unsorted[[0, 1]] = unsorted[[1, 0]]
    with pytest.raises(RuntimeError):
        capi.Schwarz(ctx, A, ix.overlap_ptr, unsorted, gptr, gblocks)
    S = capi.Schwarz(ctx, A, ix.overlap_ptr, ix.overlap, gptr, gblocks)
    r, y = ctx.vector(A.shape[0]), ctx.vector(A.shape[0])
    with pytest.raises(RuntimeError):
        S.apply(r, y)                                                               # numeric phase not run
    with pytest.raises(RuntimeError):
        S.setup()                                                                   # all-zero operator: singular blocks


@pytest.mark.parametrize("mode,shape,nl,order", [("asm8", (2, 2, 2), 3, "biquadratic"), ("asmref8", (2, 2, 2), 3, "linear"),
                                                 ("asm5", (3, 2, 2), 2, "linear"), ("asmsor4096", (2, 2, 2), 3, "biquadratic"),
                                                 ("asmilu8", (2, 2, 2), 3, "linear"), ("gmresilu", (2, 2, 2), 3, "linear")])
def test_cpp_driver_with_the_asm_level_solver(mode, shape, nl, order):
    """tests/cpp/poisson_driver.cpp through LinearEquationSolverB200Asm (the LinearEquationSolverPetscAsm surface:
    SetNumberOfSchurVariables(0), SetElementBlockNumber(n), MGSetLevel building index sets + preconditioner):
    residual after every MGsolve cycle and the final solution against the oracle."""
    import re
    from femus_b200 import hostapi
    from oracle import mesh_box as mb, mg
    from tests.test_adapters import _run_driver
    fam = {"linear": 0, "biquadratic": 2}[order]
    nb = 10 ** 6 if mode == "gmresilu" else int(mode.lstrip("asmrefsoilu"))
    ncyc = 3
    out = _run_driver(list(shape) + [nl, fam, ncyc, mode])
    res = [float(x) for x in re.findall(r"cycle \d+ residual (\S+)", out)]
    assert len(res) == ncyc + 1
    H = hostapi.HostHierarchy(*shape, nl)
    blocks, orders = [None], [None]
    for l in range(1, nl):
        ix = hostapi.AsmIndex(H.levels[l], order, nb)
        rp, ci = H.levels[l].sparsity(order)
        grp, _, gblocks = hostapi.asm_schedule(rp, ci, ix.overlap_ptr, ix.overlap, "levels" if mode.startswith("asmref") else "colours")
        blocks.append(ix.blocks())
        orders.append(gblocks)
    O = mg.Hierarchy(mb.build_hierarchy(*shape, nl), order, smoother="asm", asm_blocks=blocks, asm_orders=orders,
                     asm_sub="ssor" if "sor" in mode else ("ilu" if "ilu" in mode else "lu"), ksp="gmres" if mode == "gmresilu" else "richardson")
    trace, eps = O.mg_solve_trace(ncyc, omega=1.0)
    free = O.bdc[-1] > 1.1
    r0 = float(np.linalg.norm(np.where(free, O.rhs, 0.0)))
    assert abs(res[0] - r0) <= 1e-12 * r0
    for k in range(ncyc):
        assert abs(res[k + 1] - trace[k]) <= 1e-11 * r0, (k, res[k + 1], trace[k])
    l2 = float(re.search(r"solution l2 (\S+)", out).group(1))
    assert abs(l2 - np.linalg.norm(eps)) <= 1e-10 * np.linalg.norm(eps)
