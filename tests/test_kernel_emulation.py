"""The device sources of the kernels written without a GPU at hand (femus_b200/csrc/b2_schwarz_kernels.cuh,
b2_neumann_kernel.cuh) executed on the CPU by a thread-per-CUDA-thread emulator (tests/cpp/cuda_emu.hpp: barriers
for __syncthreads / __syncwarp, warp shuffles, atomics) and compared with the oracle.  This checks the kernels'
logic -- indexing, the synchronisation protocol, the arithmetic -- on the SAME source nvcc compiles; the GPU parity
tests (tests/test_tri_faces_gpu.py, tests/test_zz_asm_smoother_gpu.py) remain the gate for the device itself."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

from femus_b200 import hostapi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
vp = ctypes.c_void_p


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("emu") / "libemu.so")
    r = subprocess.run(["g++", "-std=c++20", "-O1", "-pthread", "-shared", "-fPIC", "-o", so, os.path.join(ROOT, "tests", "cpp", "emu_kernels.cpp")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    L = ctypes.CDLL(so)
    L.emu_schwarz.restype = ctypes.c_int
    L.emu_schwarz.argtypes = [ctypes.c_int64, vp, vp, vp, ctypes.c_int64, vp, vp, ctypes.c_int64, vp, vp, vp, vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    L.emu_neumann.restype = None
    L.emu_neumann.argtypes = [ctypes.c_int64, vp, vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp, ctypes.c_int64, vp, vp, vp, vp, ctypes.c_int]
    L.emu_stokes.restype = None
    L.emu_stokes.argtypes = [ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp,
                             ctypes.c_double, ctypes.c_int]
    L.emu_pressure_faces.restype = None
    L.emu_pressure_faces.argtypes = [ctypes.c_int64, vp, vp, vp, ctypes.c_int, ctypes.c_int, vp, vp, ctypes.c_int64, vp, vp, vp, vp, ctypes.c_int]
    L.emu_gmres.restype = ctypes.c_int
    L.emu_gmres.argtypes = [ctypes.c_int64, vp, vp, vp, vp, vp, vp, ctypes.c_int, ctypes.c_int]
    L.emu_ns.restype = None
    L.emu_ns.argtypes = L.emu_stokes.argtypes
    return L


def _p(a):
    return a.ctypes.data_as(vp)


def _run_schwarz(emu, A, ix, gptr, gblocks, r, threads=64, grid=3, sub=0):
    rp = np.ascontiguousarray(A.indptr, dtype=np.int64)
    ci = np.ascontiguousarray(A.indices, dtype=np.int32)
    va = np.ascontiguousarray(A.data, dtype=np.float64)
    bp, bd = np.ascontiguousarray(ix.overlap_ptr, dtype=np.int64), np.ascontiguousarray(ix.overlap, dtype=np.int32)
    gp, gb = np.ascontiguousarray(gptr, dtype=np.int64), np.ascontiguousarray(gblocks, dtype=np.int32)
    y = np.full(A.shape[0], np.nan)
    inv = np.zeros(int(sum(len(b) ** 2 for b in ix.blocks())))
    err = emu.emu_schwarz(A.shape[0], _p(rp), _p(ci), _p(va), ix.nblocks, _p(bp), _p(bd), len(gp) - 1, _p(gp), _p(gb), _p(r), _p(y), _p(inv),
                          threads, grid, sub)
    return err, y, inv


@pytest.mark.parametrize("order,nb,schedule", [("linear", 8, "levels"), ("linear", 8, "colours"), ("linear", 5, "colours"),
                                               ("biquadratic", 1, "colours")])
def test_schwarz_kernels_on_the_emulator(emu, order, nb, schedule):
    """extract -> Gauss-Jordan inverse -> one launch per group, on the penalised level-1 operator of a 2x2x2 box:
    block inverses against numpy, the sweep against the oracle's PCASM restatement in the same block order;
    3 CTAs of 64 threads, so every CTA strides over several blocks of a group."""
    from oracle import mesh_box as mb, mg
    shape = (2, 2, 2) if order == "linear" else (1, 2, 1)
    lv = mb.build_hierarchy(*shape, 2)
    H = hostapi.HostHierarchy(*shape, 2)
    ix = hostapi.AsmIndex(H.levels[1], order, nb)
    rp, ci = H.levels[1].sparsity(order)
    grp, gptr, gblocks = hostapi.asm_schedule(rp, ci, ix.overlap_ptr, ix.overlap, schedule)
    O = mg.Hierarchy(lv, order, dirichlet_faces=(1, 3, 6), smoother="asm", asm_blocks=[None, ix.blocks()], asm_orders=[None, gblocks])
    A = O.A[1]
    r = np.random.default_rng(7).standard_normal(A.shape[0])
    err, y, inv = _run_schwarz(emu, A, ix, gptr, gblocks, r)
    assert err == 0
    pos = 0
    for b, M in zip(ix.blocks(), O.asm[1].dense):
        m = len(b)
        got = inv[pos:pos + m * m].reshape(m, m)
        want = np.linalg.inv(M)
        assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max()
        pos += m * m
    want = O.asm[1].apply(r)
    assert np.abs(y - want).max() <= 1e-12 * np.abs(want).max()
    if schedule == "levels":        # the reference's own sweep
        assert np.abs(y - O.asm[1].apply(r, range(ix.nblocks))).max() <= 1e-12 * np.abs(want).max()


@pytest.mark.parametrize("order,nb,schedule", [("linear", 8, "levels"), ("linear", 5, "colours"), ("biquadratic", 2, "colours"),
                                               ("linear", 10 ** 6, "colours")])
def test_schwarz_ssor_kernel_on_the_emulator(emu, order, nb, schedule):
    """The sweep with one SSOR iteration per block (001_Poisson's SOR_PRECOND sub-preconditioner) against the oracle;
    applied twice on the same scratch (stale membership marks of overlapping blocks must not matter).  nb = 10^6: ONE
    block with every element = Richardson + SOR, the application's FEMuS_DEFAULT smoother."""
    from oracle import mesh_box as mb, mg
    shape = (2, 2, 2) if order == "linear" else (1, 2, 1)
    lv = mb.build_hierarchy(*shape, 2)
    H = hostapi.HostHierarchy(*shape, 2)
    ix = hostapi.AsmIndex(H.levels[1], order, nb)
    rp, ci = H.levels[1].sparsity(order)
    grp, gptr, gblocks = hostapi.asm_schedule(rp, ci, ix.overlap_ptr, ix.overlap, schedule)
    O = mg.Hierarchy(lv, order, dirichlet_faces=(1, 3, 6), smoother="asm", asm_blocks=[None, ix.blocks()], asm_orders=[None, gblocks],
                     asm_sub="ssor")
    A = O.A[1]
    r = np.random.default_rng(8).standard_normal(A.shape[0])
    err, y, _ = _run_schwarz(emu, A, ix, gptr, gblocks, r, sub=1)
    assert err == 0
    want = O.asm[1].apply(r)
    assert np.abs(y - want).max() <= 1e-12 * np.abs(want).max()
    if nb >= 10 ** 6:
        assert ix.nblocks == 1 and len(gptr) == 2


def test_vanka_blocks_of_a_saddle_point_system_on_the_emulator(emu):
    """Velocity-pressure Vanka blocks (index sets with one Schur variable: velocities of the near elements,
    pressures of the block's element) of a synthetic Stokes matrix in system numbering: the system dofs put the
    velocities first, the pressure Schur complement last (the Gauss-Jordan inverse pivots on rows); block inverses
    and the multiplicative sweep against the oracle."""
    from oracle import asm, mesh_box as mb
    from tests import saddle_point as spt
    L = mb.build_hierarchy(1, 1, 1, 2)[1]
    A = spt.stokes_matrix(L)
    H = hostapi.HostHierarchy(1, 1, 1, 2)
    ix = hostapi.AsmIndex(H.levels[1], spt.FAMILIES, 1, nschur=1)
    assert ix.nblocks == 8 and all(len(b) == 3 * 125 + 8 for b in ix.blocks())
    grp, gptr, gblocks = hostapi.asm_schedule(A.indptr, A.indices, ix.overlap_ptr, ix.overlap, "colours")
    assert len(gptr) - 1 == 8                   # every block holds every velocity: nothing commutes
    S = asm.BlockSmoother(A, ix.blocks(), order=gblocks)
    r = np.random.default_rng(9).standard_normal(A.shape[0])
    err, y, inv = _run_schwarz(emu, A, ix, gptr, gblocks, r, threads=64, grid=2)
    assert err == 0
    m = len(ix.blocks()[0])
    want_inv = np.linalg.inv(S.dense[0])
    assert np.abs(inv[:m * m].reshape(m, m) - want_inv).max() <= 1e-10 * np.abs(want_inv).max()
    want = S.apply(r)
    assert np.abs(y - want).max() <= 1e-10 * np.abs(want).max()


@pytest.mark.parametrize("order,nb,schedule", [("linear", 8, "levels"), ("linear", 5, "colours"), ("biquadratic", 2, "colours"),
                                               ("linear", 10 ** 6, "colours")])
def test_schwarz_ilu_kernels_on_the_emulator(emu, order, nb, schedule):
    """ILU(0) block solves (ILU_PRECOND on the blocks, the reference applications' usual setting): factorisation group
    by group and the sweep against the oracle's IKJ ILU(0) on the pattern of A[B, B]; one block with every element =
    Richardson + ILU(0), FEMuS_DEFAULT with ILU_PRECOND.  With a single dof per row coupling (exact pattern) ILU(0) of a
    tridiagonal-like block would be exact; here it is a genuine incomplete factorisation."""
    from oracle import mesh_box as mb, mg
    shape = (2, 2, 2) if order == "linear" else (1, 2, 1)
    lv = mb.build_hierarchy(*shape, 2)
    H = hostapi.HostHierarchy(*shape, 2)
    ix = hostapi.AsmIndex(H.levels[1], order, nb)
    rp, ci = H.levels[1].sparsity(order)
    grp, gptr, gblocks = hostapi.asm_schedule(rp, ci, ix.overlap_ptr, ix.overlap, schedule)
    O = mg.Hierarchy(lv, order, dirichlet_faces=(1, 3, 6), smoother="asm", asm_blocks=[None, ix.blocks()], asm_orders=[None, gblocks],
                     asm_sub="ilu")
    A = O.A[1]
    r = np.random.default_rng(10).standard_normal(A.shape[0])
    err, y, _ = _run_schwarz(emu, A, ix, gptr, gblocks, r, sub=2)
    assert err == 0
    want = O.asm[1].apply(r)
    assert np.abs(y - want).max() <= 1e-12 * np.abs(want).max()
    exact = mg.Hierarchy(lv, order, dirichlet_faces=(1, 3, 6), smoother="asm", asm_blocks=[None, ix.blocks()], asm_orders=[None, gblocks]).asm[1].apply(r)
    assert np.abs(want - exact).max() > 1e-6 * np.abs(exact).max()        # incomplete, not exact


@pytest.mark.parametrize("sub", ["ssor", "ilu"])
@pytest.mark.parametrize("order,nb,shape", [("linear", 8, (2, 2, 2)), ("linear", 10 ** 6, (2, 2, 2)), ("biquadratic", 8, (2, 2, 2))])
def test_level_scheduled_rows_equal_the_one_warp_walk(emu, sub, order, nb, shape):
    """b2_schwarz_set_row_levels: the rows of every block sorted into dependency levels of its triangular patterns, all
    warps of the CTA working inside a level -- the same arithmetic per row, so the result equals the one-warp walk BIT
    FOR BIT (and the oracle to 1e-12); nb = 10^6: one block with the whole level, the case it is meant for.  The staged
    walk of b2_schwarz_walk.cuh (local indices, shared-memory block vectors, next row prefetched, map-based ILU(0)
    factorisation) is the third variant of the same arithmetic."""
    from oracle import mesh_box as mb, mg      # the 4 x 4 x 4 Q2 case has rows of 343 entries: longer than one batch of the walks
    lv = mb.build_hierarchy(*shape, 2)
    H = hostapi.HostHierarchy(*shape, 2)
    ix = hostapi.AsmIndex(H.levels[1], order, nb)
    rp, ci = H.levels[1].sparsity(order)
    grp, gptr, gblocks = hostapi.asm_schedule(rp, ci, ix.overlap_ptr, ix.overlap, "colours")
    O = mg.Hierarchy(lv, order, dirichlet_faces=(1, 3, 6), smoother="asm", asm_blocks=[None, ix.blocks()], asm_orders=[None, gblocks], asm_sub=sub)
    A = O.A[1]
    r = np.random.default_rng(14).standard_normal(A.shape[0])
    code = {"ssor": 1, "ilu": 2}[sub]
    err0, y0, _ = _run_schwarz(emu, A, ix, gptr, gblocks, r, sub=code)
    err1, y1, _ = _run_schwarz(emu, A, ix, gptr, gblocks, r, sub=10 + code, threads=128)
    assert err0 == 0 and err1 == 0
    assert np.array_equal(y0, y1)
    want = O.asm[1].apply(r)
    assert np.abs(y1 - want).max() <= 1e-12 * np.abs(want).max()
    # the staged walk (what the library runs without row levels): 64-thread CTAs as the library launches them
    err2, y2, _ = _run_schwarz(emu, A, ix, gptr, gblocks, r, sub=20 + code, threads=64)
    assert err2 == 0
    assert np.array_equal(y0, y2)


def test_schwarz_invert_kernel_reports_singular_blocks(emu):
    import scipy.sparse as sp
    H = hostapi.HostHierarchy(1, 1, 1, 2)
    ix = hostapi.AsmIndex(H.levels[1], "linear", 4)
    rp, ci = H.levels[1].sparsity("linear")
    grp, gptr, gblocks = hostapi.asm_schedule(rp, ci, ix.overlap_ptr, ix.overlap, "colours")
    A = sp.csr_matrix((np.zeros(len(ci)), ci, rp))
    err, _, _ = _run_schwarz(emu, A, ix, gptr, gblocks, np.ones(A.shape[0]), threads=32, grid=1)
    assert err == 1                 # 1 + index of the first singular block


@pytest.mark.parametrize("name,order", [("cube_tet10", "quadratic"), ("cube_wedge18", "biquadratic"), ("cube_mixed", "linear"),
                                        ("cube_mixed", "quadratic"), ("cube_hex27_2x2x2", "biquadratic")])
def test_neumann_kernel_on_the_emulator(emu, name, order):
    """neumann_kernel with the (plan, face kind) groups the driver builds: triangular faces with 13 Gauss points and
    3 / 6 / 7 dofs, quadrilateral ones with 16 points and 4 / 8 / 9 dofs, against the oracle's boundary vector."""
    from femus_b200.poisson import neumann_face_groups
    from oracle import mesh_mixed as mm
    path = os.path.join(GOLDEN, name + ".neu")
    level = hostapi.HostHierarchy.from_neu(path, 1).levels[0]
    neumann = {1: 0.2, 4: -1.5, 6: 0.7}
    mixed = level.elem_type < 0
    dofs = level.system_dofs27(order)
    rhs = np.zeros(level.ndofs(order))
    xyz = np.ascontiguousarray(level.xyz, dtype=np.float64)
    ngroups = 0
    for t in (sorted(set(level.elem_types.tolist())) if mixed else [level.elem_type]):
        sel = np.nonzero(level.elem_types == t)[0] if mixed else slice(None)
        nve = hostapi.elem_nve(t, order)
        conn = np.ascontiguousarray(level.conn[sel], dtype=np.int32)
        dof = np.ascontiguousarray(dofs[sel][:, :nve], dtype=np.int32)
        for (fe, fl, fv), (phi, dxi, deta, w), fnodes in neumann_face_groups(level, order, neumann, t, sel):
            tab = np.concatenate([phi.ravel(), dxi.ravel(), deta.ravel(), w.ravel()])
            fn = np.ascontiguousarray(fnodes, dtype=np.int32)
            emu.emu_neumann(len(fe), _p(fe), _p(fl), _p(fv), phi.shape[1], phi.shape[0], nve, _p(tab), _p(fn), level.nnode, _p(xyz), _p(conn),
                            _p(dof), _p(rhs), 1)
            ngroups += 1
    want = mm.neumann_rhs(mm.read_neu(path), order, neumann)
    assert ngroups >= 1 and np.abs(rhs - want).max() <= 1e-13 * np.abs(want).max()


@pytest.mark.parametrize("guarded", [False, True], ids=["pair_instantiation", "guarded_instantiation"])
@pytest.mark.parametrize("name,order_v,order_p", [("box", "biquadratic", "linear"), ("cube_tet10", "quadratic", "linear"),
                                                  ("cube_wedge18", "biquadratic", "linear")])
def test_stokes_kernel_on_the_emulator(emu, name, order_v, order_p, guarded):
    """stokes_kernel (one warp per element, table-driven) on hexahedra (Q2-Q1, 64 points), tetrahedra (P2-P1, 31
    points) and wedges (21 / 6 dofs, 52 points): system matrix on the host-built pattern and residual at a random
    solution against the oracle's restatement of SteadyStokes/main.cpp:290-598.  Both the instantiation of the element
    pair and the guarded one that any other (nv, np) would run."""
    from oracle import stokes, mesh_box as mb, mesh_mixed as mm, fe_hex, mg
    fams = [order_v] * 3 + [order_p]
    if name == "box":
        H, L, mesh = hostapi.HostHierarchy(2, 1, 2, 1), mb.build_hierarchy(2, 1, 2, 1)[0], mb
        tables_of = lambda t, o: fe_hex.tables(o)
    else:
        path = os.path.join(GOLDEN, name + ".neu")
        H, L, mesh = hostapi.HostHierarchy.from_neu(path, 1), mm.read_neu(path), mm
        tables_of = lambda t, o: mm.FE[t].tables(o)
    level = H.levels[0]
    S = hostapi.SystemOnLevel(level, fams)
    rp, ci = S.sparsity()
    edof = np.ascontiguousarray(S.elem_dofs(), dtype=np.int32)
    t = level.elem_type
    tv, tp = hostapi.elem_tables(t, order_v), hostapi.elem_tables(t, order_p)
    nv, npr, ng = tv[0].shape[1], tp[0].shape[1], tv[4].shape[0]
    tabv = np.concatenate([tv[1].ravel(), tv[2].ravel(), tv[3].ravel(), tv[4].ravel()])
    tabp = np.ascontiguousarray(tp[0])
    sol = np.random.default_rng(11).standard_normal(S.n)
    val, rhs = np.zeros(len(ci)), np.zeros(S.n)
    xyz, conn = np.ascontiguousarray(level.xyz), np.ascontiguousarray(level.conn, dtype=np.int32)
    IRe = 0.37
    emu.emu_stokes(level.nel, level.nnode, nv, npr, ng, _p(xyz), _p(conn), _p(edof), _p(tabv), _p(tabp), _p(rp), _p(ci), _p(val), _p(sol), _p(rhs),
                   IRe, -2 if guarded else 2)
    Aref, rref = stokes.assemble(L, mesh, order_v, order_p, sol, IRe, tables_of)
    Aref = mg.on_pattern(Aref, rp, ci)
    assert np.abs(val - Aref.data).max() <= 1e-12 * np.abs(Aref.data).max()
    assert np.abs(rhs - rref).max() <= 1e-12 * (np.abs(Aref) @ np.abs(sol)).max()


@pytest.mark.parametrize("guarded", [False, True], ids=["pair_instantiation", "guarded_instantiation"])
@pytest.mark.parametrize("name,order_v,order_p", [("box", "biquadratic", "linear"), ("cube_tet10", "quadratic", "linear")])
def test_navier_stokes_kernel_on_the_emulator(emu, name, order_v, order_p, guarded):
    """ns_kernel (one CTA per element): residual RES = -aRes and the analytic Newton Jacobian at a random solution
    against the oracle's restatement of 03_navier_stokes.hpp:305-413 (whose Jacobian is checked against finite
    differences in tests/test_oracle_ns.py and against the reference's adept Jacobian in tests/test_reference_pin_stokes.py).
    Both the instantiation with compile-time pair counts and the guarded one."""
    from oracle import navier_stokes as ons, mesh_box as mb, mesh_mixed as mm, fe_hex, mg
    fams = [order_v] * 3 + [order_p]
    if name == "box":
        H, L, mesh = hostapi.HostHierarchy(2, 1, 1, 1), mb.build_hierarchy(2, 1, 1, 1)[0], mb
        tables_of = lambda t, o: fe_hex.tables(o)
    else:
        path = os.path.join(GOLDEN, name + ".neu")
        H, L, mesh = hostapi.HostHierarchy.from_neu(path, 1), mm.read_neu(path), mm
        tables_of = lambda t, o: mm.FE[t].tables(o)
    level = H.levels[0]
    S = hostapi.SystemOnLevel(level, fams)
    rp, ci = S.sparsity()
    edof = np.ascontiguousarray(S.elem_dofs(), dtype=np.int32)
    t = level.elem_type
    tv, tp = hostapi.elem_tables(t, order_v), hostapi.elem_tables(t, order_p)
    nv, npr, ng = tv[0].shape[1], tp[0].shape[1], tv[4].shape[0]
    tabv = np.concatenate([tv[0].ravel(), tv[1].ravel(), tv[2].ravel(), tv[3].ravel(), tv[4].ravel()])
    tabp = np.ascontiguousarray(tp[0])
    sol = 0.5 * np.random.default_rng(12).standard_normal(S.n)
    val, rhs = np.zeros(len(ci)), np.zeros(S.n)
    xyz, conn = np.ascontiguousarray(level.xyz), np.ascontiguousarray(level.conn, dtype=np.int32)
    nu = 0.21
    emu.emu_ns(level.nel, level.nnode, nv, npr, ng, _p(xyz), _p(conn), _p(edof), _p(tabv), _p(tabp), _p(rp), _p(ci), _p(val), _p(sol), _p(rhs), nu,
               -3 if guarded else 3)
    Aref, rref = ons.assemble(L, mesh, order_v, order_p, sol, nu, tables_of)
    Aref = mg.on_pattern(Aref, rp, ci)
    assert np.abs(val - Aref.data).max() <= 1e-12 * np.abs(Aref.data).max()
    assert np.abs(rhs - rref).max() <= 1e-12 * np.abs(rref).max()


def test_stokes_plans_of_a_mixed_mesh_on_the_emulator(emu):
    """One launch per element type of the reference's mixed mesh (hexahedra, tetrahedra, wedges; Q2-Q1 / P2-P1 /
    21-6 dofs), all accumulating into ONE system matrix and residual, as the driver's plans do."""
    from oracle import stokes, mesh_mixed as mm, mg
    path = os.path.join(GOLDEN, "cube_mixed.neu")
    level, L = hostapi.HostHierarchy.from_neu(path, 1).levels[0], mm.read_neu(path)
    fams = ["biquadratic"] * 3 + ["linear"]
    S = hostapi.SystemOnLevel(level, fams)
    rp, ci = S.sparsity()
    edofs = S.elem_dofs()
    sol = np.random.default_rng(13).standard_normal(S.n)
    val, rhs = np.zeros(len(ci)), np.zeros(S.n)
    xyz = np.ascontiguousarray(level.xyz)
    for t in sorted(set(level.elem_types.tolist())):
        sel = np.nonzero(level.elem_types == t)[0]
        tv, tp = hostapi.elem_tables(t, "biquadratic"), hostapi.elem_tables(t, "linear")
        nv, npr, ng = tv[0].shape[1], tp[0].shape[1], tv[4].shape[0]
        tabv = np.concatenate([tv[1].ravel(), tv[2].ravel(), tv[3].ravel(), tv[4].ravel()])
        tabp = np.ascontiguousarray(tp[0])
        conn = np.ascontiguousarray(level.conn[sel], dtype=np.int32)
        edof = np.ascontiguousarray(edofs[sel], dtype=np.int32)
        emu.emu_stokes(len(sel), level.nnode, nv, npr, ng, _p(xyz), _p(conn), _p(edof), _p(tabv), _p(tabp), _p(rp), _p(ci), _p(val), _p(sol),
                       _p(rhs), 0.5, 2)
    Aref, rref = stokes.assemble(L, mm, "biquadratic", "linear", sol, 0.5, lambda t, o: mm.FE[t].tables(o))
    Aref = mg.on_pattern(Aref, rp, ci)
    assert np.abs(val - Aref.data).max() <= 1e-12 * np.abs(Aref.data).max()
    assert np.abs(rhs - rref).max() <= 1e-12 * (np.abs(Aref) @ np.abs(sol)).max()


@pytest.mark.parametrize("name,order_v", [("cube_tet10", "quadratic"), ("cube_wedge18", "biquadratic"), ("cube_hex27_2x2x2", "biquadratic")])
def test_boundary_pressure_kernel_on_the_emulator(emu, name, order_v):
    """pressure_face_kernel: RES[U_k] -= int phi_i tau n_k over triangular and quadrilateral boundary faces (the boundary
    block of 03_navier_stokes.hpp:196-300) against the oracle; on a closed surface with constant tau the contributions of
    every component sum to zero (divergence theorem)."""
    from femus_b200.poisson import neumann_face_groups
    from oracle import navier_stokes as ons, mesh_mixed as mm
    path = os.path.join(GOLDEN, name + ".neu")
    level, L = hostapi.HostHierarchy.from_neu(path, 1).levels[0], mm.read_neu(path)
    fams = [order_v] * 3 + ["linear"]
    S = hostapi.SystemOnLevel(level, fams)
    edof = np.ascontiguousarray(S.elem_dofs(), dtype=np.int32)
    xyz, conn = np.ascontiguousarray(level.xyz), np.ascontiguousarray(level.conn, dtype=np.int32)
    for tau in ({2: 1.5, 5: -0.4}, {b: 1.0 for b in range(1, 7)}):
        rhs = np.zeros(S.n)
        for (fe, fl, fv), (phi, dxi, deta, w), fnodes in neumann_face_groups(level, order_v, tau, level.elem_type, slice(None)):
            tab = np.concatenate([phi.ravel(), dxi.ravel(), deta.ravel(), w.ravel()])
            fn = np.ascontiguousarray(fnodes, dtype=np.int32)
            emu.emu_pressure_faces(len(fe), _p(fe), _p(fl), _p(fv), phi.shape[1], phi.shape[0], _p(tab), _p(fn), level.nnode, _p(xyz), _p(conn),
                                   _p(edof), _p(rhs), 2)
        want = ons.pressure_boundary_rhs(L, mm, order_v, "linear", tau)
        assert np.abs(want).max() > 0 and np.abs(rhs - want).max() <= 1e-13 * np.abs(want).max()
        if len(tau) == 6:
            nq = level.ndofs(order_v)
            for k in range(3):
                assert abs(rhs[k * nq:(k + 1) * nq].sum()) <= 1e-13


def _libtsan():
    r = subprocess.run(["gcc", "-print-file-name=libtsan.so"], capture_output=True, text=True)
    p = r.stdout.strip()
    return p if r.returncode == 0 and os.path.isabs(p) and os.path.exists(p) else None


@pytest.mark.skipif(_libtsan() is None, reason="libtsan not available")
def test_kernels_are_race_free_under_thread_sanitizer(tmp_path):
    """The emulator plays every CUDA thread with a host thread and every __syncthreads / __syncwarp with a barrier, so a
    missing synchronisation in a kernel IS a data race ThreadSanitizer reports.  First the detector is proven on a probe
    kernel (racy without its barrier, clean with it); then the block-smoother (exact, SSOR, ILU(0)), Stokes,
    Navier-Stokes, Neumann and boundary-pressure kernels run under it against the oracle without a single report."""
    cpp = os.path.join(ROOT, "tests", "cpp")
    probe = str(tmp_path / "probe")
    r = subprocess.run(["g++", "-std=c++20", "-O1", "-g", "-pthread", "-fsanitize=thread", "-o", probe, os.path.join(cpp, "emu_race_probe.cpp")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    env = dict(os.environ, TSAN_OPTIONS="exitcode=0 report_signal_unsafe=0")
    racy = subprocess.run([probe], capture_output=True, text=True, env=env)
    clean = subprocess.run([probe, "sync"], capture_output=True, text=True, env=env)
    if "FATAL: ThreadSanitizer" in racy.stderr + clean.stderr:
        pytest.skip("ThreadSanitizer cannot run in this environment: " + (racy.stderr + clean.stderr)[:200])
    assert "WARNING: ThreadSanitizer: data race" in racy.stderr and "ThreadSanitizer" not in clean.stderr
    so = str(tmp_path / "libemu_tsan.so")
    r = subprocess.run(["g++", "-std=c++20", "-O1", "-pthread", "-shared", "-fPIC", "-fsanitize=thread", "-o", so, os.path.join(cpp, "emu_kernels.cpp")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    env["LD_PRELOAD"] = _libtsan()
    run = subprocess.run([sys.executable, os.path.join(cpp, "tsan_runner.py"), so], capture_output=True, text=True, env=env, timeout=1500)
    assert "tsan-run-finished" in run.stdout, run.stdout[-2000:] + run.stderr[-4000:]
    assert "WARNING: ThreadSanitizer" not in run.stderr, run.stderr[:6000]


@pytest.mark.parametrize("k", [1, 2, 4])
def test_library_gmres_cycle_on_host_vectors(emu, k):
    """The GMRES cycle of the level solver (femus_b200/csrc/b2_gmres.hpp: the code b2_mg.cu runs on device vectors)
    instantiated on host vectors with a CSR operator and the Jacobi preconditioner: equal to the oracle's KSPGMRES
    restatement from a non-zero and from a zero initial guess, and, when the Krylov space is exhausted, exact."""
    from oracle import mesh_box as mb, mg
    O = mg.Hierarchy(mb.build_hierarchy(2, 2, 2, 2), "linear", ksp="gmres")
    A = O.A[1]
    rp, ci, va = np.ascontiguousarray(A.indptr, dtype=np.int64), np.ascontiguousarray(A.indices, dtype=np.int32), np.ascontiguousarray(A.data)
    dinv = np.ascontiguousarray(O.dinv[1])
    rng = np.random.default_rng(15)
    b, x0 = rng.standard_normal(A.shape[0]), rng.standard_normal(A.shape[0])
    for zero in (0, 1):
        x = x0.copy()
        assert emu.emu_gmres(A.shape[0], _p(rp), _p(ci), _p(va), _p(dinv), _p(b), _p(x), k, zero) == 0
        want = O.gmres(1, np.zeros_like(x0) if zero else x0, b, k)
        assert np.abs(x - want).max() <= 1e-12 * np.abs(want).max()
    # exhausted Krylov space (happy breakdown): a 3 x 3 diagonal operator, 5 iterations asked for -> the exact solution
    d = np.array([1.0, 2.0, 4.0])
    rp3, ci3 = np.arange(4, dtype=np.int64), np.arange(3, dtype=np.int32)
    b3, x3, one = np.array([3.0, -1.0, 0.5]), np.zeros(3), np.ones(3)
    assert emu.emu_gmres(3, _p(rp3), _p(ci3), _p(d), _p(one), _p(b3), _p(x3), 5, 1) == 0
    assert np.abs(x3 - b3 / d).max() <= 1e-14
    # zero right-hand side: nothing to do, no division by zero
    z3, x3 = np.zeros(3), np.zeros(3)
    assert emu.emu_gmres(3, _p(rp3), _p(ci3), _p(d), _p(one), _p(z3), _p(x3), 2, 1) == 0 and not x3.any()


# ---- the multigrid ORCHESTRATION (b2_mg.cu + b2_vec.cu + b2_schwarz.cu) compiled for the emulator ------------------
@pytest.fixture(scope="module")
def emu_mg(tmp_path_factory):
    """b2_vec.cu, b2_schwarz.cu, b2_mg.cu compiled with g++ (-include emu_prefix.hpp: CUDA keywords, B2_LAUNCH ->
    emu::launch; emu_rt/cuda_runtime.h: stub runtime) + tests/cpp/emu_mg.cpp (plain-CSR stand-ins for the SpMV family)."""
    so = _build_emulated_library(tmp_path_factory.mktemp("emu_mg"), [])
    L = ctypes.CDLL(so)
    L.emu_mg_run.restype = ctypes.c_int
    L.emu_stokes_plan.restype = ctypes.c_int
    L.b2_last_error.restype = ctypes.c_char_p
    return L


def _build_emulated_library(outdir, extra_flags):
    """b2_vec.cu, b2_schwarz.cu, b2_mg.cu, b2_stokes.cu + emu_mg.cpp compiled (in parallel) for the emulator and linked."""
    from concurrent.futures import ThreadPoolExecutor
    cpp = os.path.join(ROOT, "tests", "cpp")
    srcs = [os.path.join(ROOT, "femus_b200", "csrc", f) for f in ("b2_vec.cu", "b2_schwarz.cu", "b2_mg.cu", "b2_stokes.cu")] + [os.path.join(cpp, "emu_mg.cpp")]

    def compile_one(src):
        obj = os.path.join(str(outdir), os.path.basename(src) + ".o")
        r = subprocess.run(["g++", "-std=c++20", "-O1", "-pthread", "-fPIC"] + extra_flags + ["-I", os.path.join(cpp, "emu_rt"), "-I", cpp, "-include",
                            os.path.join(cpp, "emu_prefix.hpp"), "-x", "c++", "-c", src, "-o", obj], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-3000:]
        return obj

    with ThreadPoolExecutor(max_workers=5) as ex:
        objs = list(ex.map(compile_one, srcs))
    so = os.path.join(str(outdir), "libemu_mg.so")
    r = subprocess.run(["g++", "-shared", "-pthread"] + extra_flags + ["-o", so] + objs, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return so


def _ptr_array(arrays):
    keep = [None if a is None else np.ascontiguousarray(a) for a in arrays]
    return (vp * len(keep))(*[None if a is None else a.ctypes.data_as(vp) for a in keep]), keep


def _run_emu_mg(L, O, blocks, orders, smoother, sub, ksp, coarse_direct, row_levels, npre, omega, ncycles):
    nl = len(O.levels)
    n = np.array([A.shape[0] for A in O.A_raw], dtype=np.int64)
    As = [A.copy() for A in O.A_raw]
    for A in As:
        A.sort_indices()
    rp, k1 = _ptr_array([A.indptr.astype(np.int64) for A in As])
    col, k2 = _ptr_array([A.indices.astype(np.int32) for A in As])
    val, k3 = _ptr_array([A.data.astype(np.float64) for A in As])
    Ps = [None] + [O.P[l].tocsr() for l in range(1, nl)]
    prp, k4 = _ptr_array([None if P is None else P.indptr.astype(np.int64) for P in Ps])
    pcol, k5 = _ptr_array([None if P is None else P.indices.astype(np.int32) for P in Ps])
    pval, k6 = _ptr_array([None if P is None else P.data.astype(np.float64) for P in Ps])
    nbdc = np.array([len(b) for b in O.bdc_idx], dtype=np.int64)
    bdc, k7 = _ptr_array([np.asarray(b, dtype=np.int32) for b in O.bdc_idx])
    nblk, ngrp = np.zeros(nl, dtype=np.int64), np.zeros(nl, dtype=np.int64)
    bp, bd, gp, gb = [None] * nl, [None] * nl, [None] * nl, [None] * nl
    for l in range(1, nl):
        if blocks is None:
            continue
        nblk[l] = len(blocks[l])
        bp[l] = np.concatenate([[0], np.cumsum([len(b) for b in blocks[l]])]).astype(np.int64)
        bd[l] = np.concatenate(blocks[l]).astype(np.int32)
        grp = orders[l]["grp"]
        ngrp[l] = grp.max() + 1
        gp[l] = np.concatenate([[0], np.cumsum(np.bincount(grp, minlength=ngrp[l]))]).astype(np.int64)
        gb[l] = np.argsort(grp, kind="stable").astype(np.int32)
    bpp, k8 = _ptr_array(bp)
    bdp, k9 = _ptr_array(bd)
    gpp, k10 = _ptr_array(gp)
    gbp, k11 = _ptr_array(gb)
    res = O.rhs.astype(np.float64).copy()
    eps = np.zeros_like(res)
    free = (O.bdc[-1] > 1.1).astype(np.uint8)
    trace = np.zeros(ncycles)
    rc = L.emu_mg_run(ctypes.c_int(nl), _p(n), rp, col, val, prp, pcol, pval, _p(nbdc), bdc, ctypes.c_int(smoother), ctypes.c_int(sub), ctypes.c_int(ksp),
                      ctypes.c_int(coarse_direct), ctypes.c_int(row_levels), _p(nblk), bpp, bdp, _p(ngrp), gpp, gbp, ctypes.c_int(npre), ctypes.c_int(npre),
                      ctypes.c_double(omega), ctypes.c_int(ncycles), _p(res), _p(eps), _p(free), _p(trace))
    assert rc == 0, L.b2_last_error().decode()
    return trace, eps


@pytest.mark.parametrize("case", ["jacobi", "gmres_jacobi", "asm_lu", "asm_ilu_levels_gmres", "asm_ssor_direct"])
def test_multigrid_orchestration_on_the_emulator(emu_mg, case):
    """MGSetLevel on every level + MGSolve cycles of the REAL b2_mg.cu (penalty, Galerkin-free level setup, V-cycle,
    residual update; two cycles), b2_vec.cu and b2_schwarz.cu on the CPU emulator: Richardson + Jacobi, GMRES + Jacobi, Richardson +
    element blocks (exact), GMRES + ILU(0) blocks with level-scheduled rows, SSOR blocks with the direct coarse solve --
    residual traces and corrections against the oracle V-cycle."""
    from oracle import mesh_box as mb, mg
    lv = mb.build_hierarchy(2, 2, 2, 2)
    H = hostapi.HostHierarchy(2, 2, 2, 2)
    order = "linear"
    cfg = {"jacobi": dict(smoother=0, sub=0, ksp=0, direct=0, rowlev=0, omega=0.5),
           "gmres_jacobi": dict(smoother=0, sub=0, ksp=1, direct=0, rowlev=0, omega=0.5),
           "asm_lu": dict(smoother=2, sub=0, ksp=0, direct=0, rowlev=0, omega=1.0),
           "asm_ilu_levels_gmres": dict(smoother=2, sub=2, ksp=1, direct=0, rowlev=1, omega=1.0),
           "asm_ssor_direct": dict(smoother=2, sub=1, ksp=0, direct=1, rowlev=0, omega=1.0)}[case]
    blocks = orders = None
    kw = {}
    if cfg["smoother"] == 2:
        ix = hostapi.AsmIndex(H.levels[1], order, 8)
        rp, ci = H.levels[1].sparsity(order)
        grp, gptr, gblocks = hostapi.asm_schedule(rp, ci, ix.overlap_ptr, ix.overlap, "colours")
        blocks, orders = [None, ix.blocks()], [None, {"grp": grp}]
        kw = dict(smoother="asm", asm_blocks=blocks, asm_orders=[None, gblocks], asm_sub={0: "lu", 1: "ssor", 2: "ilu"}[cfg["sub"]])
    O = mg.Hierarchy(lv, order, ksp="gmres" if cfg["ksp"] else "richardson", **kw)
    npre = 2 if cfg["ksp"] else 1
    trace, eps = _run_emu_mg(emu_mg, O, blocks, orders, cfg["smoother"], cfg["sub"], cfg["ksp"], cfg["direct"], cfg["rowlev"], npre, cfg["omega"], 2)
    trace_ref, eps_ref = O.mg_solve_trace(2, npre=npre, npost=npre, omega=cfg["omega"])
    r0 = float(np.linalg.norm(np.where(O.bdc[-1] > 1.1, O.rhs, 0.0)))           # the scale: the residual before the first cycle
    for a, b in zip(trace, trace_ref):
        assert abs(a - b) <= 1e-11 * r0, (trace, trace_ref, r0)
    assert trace[-1] < 0.5 * r0
    assert np.abs(eps - eps_ref).max() <= 1e-10 * np.abs(eps_ref).max()


def test_stokes_vcycle_orchestration_on_the_emulator(emu_mg):
    """The velocity-pressure path of b2_mg.cu on the emulator: indefinite level operators (steady Stokes, oracle-assembled,
    Galerkin coarse operator), Vanka blocks with the pressure as Schur variable and exact block solves by the Gauss-Jordan
    kernel, DIRECT coarse solve through a one-block Schwarz object; two V-cycles against the oracle (a through-flow
    channel on a 1 -> 8 element hierarchy: small enough for the emulator, 8 blocks of 383 dofs)."""
    from oracle import stokes, mesh_box as mb, fe_hex, system as osys, mg
    lv, H = mb.build_hierarchy(1, 1, 1, 2), hostapi.HostHierarchy(1, 1, 1, 2)
    fams = ["biquadratic"] * 3 + ["linear"]
    walls = (3, 4, 5, 6)          # two open boundary sets: the one-element coarse problem stays well posed
    S = hostapi.SystemOnLevel(H.levels[-1], fams)
    rp, ci = S.sparsity()
    sol = np.zeros(S.n)
    sol[osys.bdc(lv[-1], mb, fams, [(6,), (), (), ()]) < 1.5] = 1.0
    A, rhs = stokes.assemble(lv[-1], mb, "biquadratic", "linear", sol, 1.0, lambda t, o: fe_hex.tables(o))
    ix = hostapi.AsmIndex(H.levels[1], fams, 1, nschur=1)
    grp, gptr, gblocks = hostapi.asm_schedule(rp, ci, ix.overlap_ptr, ix.overlap, "colours")
    O = mg.Hierarchy(lv, None, mesh=osys.SystemMesh(mb, fams, [walls] * 3 + [()]), A_top=mg.on_pattern(A, rp, ci), rhs=rhs, smoother="asm",
                     asm_blocks=[None, ix.blocks()], asm_orders=[None, gblocks])
    trace, eps = _run_emu_mg(emu_mg, O, [None, ix.blocks()], [None, {"grp": grp}], 2, 0, 0, 1, 0, 1, 1.0, 2)
    trace_ref, eps_ref = O.mg_solve_trace(2, omega=1.0)
    r0 = float(np.linalg.norm(np.where(O.bdc[-1] > 1.1, O.rhs, 0.0)))
    for a, b in zip(trace, trace_ref):
        assert abs(a - b) <= 1e-10 * r0, (trace, trace_ref, r0)
    assert trace[-1] < 1e-2 * r0
    assert np.abs(eps - eps_ref).max() <= 1e-9 * np.abs(eps_ref).max()


@pytest.mark.parametrize("ns", [0, 1])
def test_stokes_plan_entry_points_on_the_emulator(emu_mg, ns):
    """b2_stokes_create / b2_ns_create (table packing, shared-memory sizing), b2_stokes_assemble / b2_ns_assemble and
    b2_ns_pressure_faces of the real b2_stokes.cu, called as the bindings call them, against the oracle (P2-P1 tetrahedra:
    31 Gauss points, triangular faces with 13)."""
    from femus_b200.poisson import neumann_face_groups
    from oracle import stokes, navier_stokes as ons, mesh_mixed as mm, mg
    path = os.path.join(GOLDEN, "cube_tet10.neu")
    level, Lm = hostapi.HostHierarchy.from_neu(path, 1).levels[0], mm.read_neu(path)
    fams = ["quadratic"] * 3 + ["linear"]
    S = hostapi.SystemOnLevel(level, fams)
    rp, ci = S.sparsity()
    edof = np.ascontiguousarray(S.elem_dofs(), dtype=np.int32)
    t = level.elem_type
    tv, tp = [np.ascontiguousarray(a) for a in hostapi.elem_tables(t, "quadratic")], [np.ascontiguousarray(a) for a in hostapi.elem_tables(t, "linear")]
    xyz, conn = np.ascontiguousarray(level.xyz), np.ascontiguousarray(level.conn, dtype=np.int32)
    sol = 0.4 * np.random.default_rng(16).standard_normal(S.n)
    tau = {2: 1.5, 5: -0.4}
    groups = neumann_face_groups(level, "quadratic", tau, t, slice(None)) if ns else []
    val, rhs = np.zeros(len(ci)), np.zeros(S.n)
    args = [ctypes.c_int64(level.nnode), ctypes.c_int64(level.nel), _p(xyz), _p(conn), ctypes.c_int64(S.n), _p(rp), _p(ci), _p(edof),
            ctypes.c_int(tv[0].shape[1]), ctypes.c_int(tp[0].shape[1]), ctypes.c_int(tv[4].shape[0]), _p(tv[0]), _p(tv[1]), _p(tv[2]), _p(tv[3]), _p(tv[4]),
            _p(tp[0]), _p(sol), ctypes.c_double(0.3), ctypes.c_int(ns)]
    if groups:
        (fe, fl, fv), (phi, dxi, deta, w), fnodes = groups[0]
        fn = np.ascontiguousarray(fnodes, dtype=np.int32)
        keep = [np.ascontiguousarray(a) for a in (fe, fl, fv, phi, dxi, deta, w)]
        args += [ctypes.c_int64(len(fe)), _p(keep[0]), _p(keep[1]), _p(keep[2]), ctypes.c_int(phi.shape[1]), ctypes.c_int(phi.shape[0]), _p(keep[3]),
                 _p(keep[4]), _p(keep[5]), _p(keep[6]), _p(fn)]
    else:
        args += [ctypes.c_int64(0), None, None, None, ctypes.c_int(0), ctypes.c_int(0), None, None, None, None, None]
    rc = emu_mg.emu_stokes_plan(*args, _p(val), _p(rhs))
    assert rc == 0, emu_mg.b2_last_error().decode()
    tof = lambda tt, o: mm.FE[tt].tables(o)      # noqa: E731
    if ns:
        Aref, rref = ons.assemble(Lm, mm, "quadratic", "linear", sol, 0.3, tof)
        rref = rref + ons.pressure_boundary_rhs(Lm, mm, "quadratic", "linear", tau)
    else:
        Aref, rref = stokes.assemble(Lm, mm, "quadratic", "linear", sol, 0.3, tof)
    Aref = mg.on_pattern(Aref, rp, ci)
    assert np.abs(val - Aref.data).max() <= 1e-12 * np.abs(Aref.data).max()
    assert np.abs(rhs - rref).max() <= 1e-12 * np.abs(rref).max()


@pytest.mark.parametrize("ns", [0, 1])
def test_stokes_plan_rejects_a_pattern_without_an_element_coupling(emu_mg, ns):
    """The element -> CSR slot map is built at plan creation; an element coupling that is not an entry of the system
    matrix's pattern (here: one velocity-pressure entry removed from the pattern) makes b2_stokes_create / b2_ns_create
    fail loudly instead of scattering into a neighbouring slot."""
    path = os.path.join(GOLDEN, "cube_tet10.neu")
    level = hostapi.HostHierarchy.from_neu(path, 1).levels[0]
    fams = ["quadratic"] * 3 + ["linear"]
    S = hostapi.SystemOnLevel(level, fams)
    rp, ci = S.sparsity()
    edof = np.ascontiguousarray(S.elem_dofs(), dtype=np.int32)
    row, c = int(edof[0, 0, 0]), int(edof[0, 3, 1])                 # (U dof of node 0, P dof of node 1) of element 0
    k = rp[row] + int(np.searchsorted(ci[rp[row]:rp[row + 1]], c))
    assert ci[k] == c
    ci2 = np.ascontiguousarray(np.delete(ci, k))
    rp2 = rp.copy()
    rp2[row + 1:] -= 1
    t = level.elem_type
    tv, tp = [np.ascontiguousarray(a) for a in hostapi.elem_tables(t, "quadratic")], [np.ascontiguousarray(a) for a in hostapi.elem_tables(t, "linear")]
    xyz, conn = np.ascontiguousarray(level.xyz), np.ascontiguousarray(level.conn, dtype=np.int32)
    sol, val, rhs = np.zeros(S.n), np.zeros(len(ci2)), np.zeros(S.n)
    args = [ctypes.c_int64(level.nnode), ctypes.c_int64(level.nel), _p(xyz), _p(conn), ctypes.c_int64(S.n), _p(rp2), _p(ci2), _p(edof),
            ctypes.c_int(tv[0].shape[1]), ctypes.c_int(tp[0].shape[1]), ctypes.c_int(tv[4].shape[0]), _p(tv[0]), _p(tv[1]), _p(tv[2]), _p(tv[3]), _p(tv[4]),
            _p(tp[0]), _p(sol), ctypes.c_double(0.3), ctypes.c_int(ns),
            ctypes.c_int64(0), None, None, None, ctypes.c_int(0), ctypes.c_int(0), None, None, None, None, None]
    rc = emu_mg.emu_stokes_plan(*args, _p(val), _p(rhs))
    assert rc != 0
    assert "not an entry of the system matrix's pattern" in emu_mg.b2_last_error().decode()
    assert not val.any()


@pytest.mark.skipif(_libtsan() is None, reason="libtsan not available")
def test_orchestration_is_race_free_under_thread_sanitizer(tmp_path):
    """b2_vec.cu (two-level reductions with a ticket counter), b2_schwarz.cu, b2_mg.cu, b2_stokes.cu built with
    ThreadSanitizer for the emulator: V-cycles with GMRES + Jacobi (dot products and norms of the vector layer) and with SSOR
    blocks + the direct coarse solve run against the oracle without a report (the ILU kernels are covered at kernel level)."""
    cpp = os.path.join(ROOT, "tests", "cpp")
    so = _build_emulated_library(tmp_path, ["-fsanitize=thread"])
    env = dict(os.environ, TSAN_OPTIONS="exitcode=0 report_signal_unsafe=0", LD_PRELOAD=_libtsan())
    run = subprocess.run([sys.executable, os.path.join(cpp, "tsan_runner_mg.py"), so, "gmres_jacobi", "asm_ssor_direct"], capture_output=True,
                         text=True, env=env, timeout=1500)
    if "FATAL: ThreadSanitizer" in run.stderr:
        pytest.skip("ThreadSanitizer cannot run in this environment: " + run.stderr[:200])
    assert "tsan-run-finished" in run.stdout, run.stdout[-2000:] + run.stderr[-4000:]
    assert "WARNING: ThreadSanitizer" not in run.stderr, run.stderr[:6000]


# ---- the sum-factorised Hex27 assembly kernel (b2_assemble_sumfac.cuh) -------------------------------------------------
@pytest.fixture(scope="module")
def emu_sf(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("emu_sf") / "libemu_sumfac.so")
    r = subprocess.run(["g++", "-std=c++20", "-O1", "-pthread", "-shared", "-fPIC", "-o", so, os.path.join(ROOT, "tests", "cpp", "emu_sumfac.cpp")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    L = ctypes.CDLL(so)
    L.emu_sumfac.restype = ctypes.c_int
    L.emu_sumfac.argtypes = ([ctypes.c_int64, ctypes.c_int64] + [vp] * 13 + [ctypes.c_double, ctypes.c_double, ctypes.c_int] + [vp] * 8 +
                             [ctypes.c_int])
    return L


def _distort(L, amp):
    x, y, z = L.xyz
    out = L.xyz.copy()
    out[0] += amp * np.sin(2 * np.pi * y) * np.sin(np.pi * z) * x * (1 - x)
    out[1] += amp * np.sin(3 * np.pi * z) * np.sin(np.pi * x) * y * (1 - y)
    out[2] += amp * np.sin(2 * np.pi * x) * np.sin(np.pi * y) * z * (1 - z)
    return out


@pytest.mark.parametrize("amp,gal", [(0.0, 0), (0.07, 0), (0.07, 1)])
def test_sum_factorised_assembly_on_the_emulator(emu_sf, amp, gal):
    """assemble_q2_sumfac_kernel (SOURCE of the GPU kernel, on host threads) against the oracle: fine matrix and
    residual on a distorted 2-level mesh with a non-zero solution; fused: the Galerkin coarse operator P^T A P with
    the Dirichlet rows / columns of P zeroed, and the recorded coarse element matrices summing to it."""
    from femus_b200 import hostapi
    from oracle import fe_hex, mesh_box as mb
    import scipy.sparse as sp
    order = "biquadratic"
    lv = mb.build_hierarchy(1, 2, 1, 2)
    C, F = lv
    F.xyz = _distort(F, amp)
    n, d = mb.ndofs(F, order), np.ascontiguousarray(mb.system_dof(F, order), dtype=np.int32)
    u = np.random.default_rng(5).standard_normal(n)
    Aref, rhs_ref = mb.assemble(F, order, u, fsrc=1.5)
    rp, ci = Aref.indptr.astype(np.int64), Aref.indices.astype(np.int32)
    val, rhs = np.zeros(Aref.nnz), np.zeros(n)
    phi, dxi, deta, dzeta, w = [np.ascontiguousarray(t) for t in fe_hex.tables(order)]
    xyz, conn = np.ascontiguousarray(F.xyz), np.ascontiguousarray(F.conn, dtype=np.int32)
    null = ctypes.c_void_p(0)
    args_gal = [null] * 8
    if gal:
        nc = mb.ndofs(C, order)
        cd = np.ascontiguousarray(mb.system_dof(C, order), dtype=np.int32)
        P = mb.prolongator(C, F, order).tocsr()
        bf, bc = mb.bdc_flags(F, order) < 1.5, mb.bdc_flags(C, order) < 1.5
        # child prolongators from the assembled P: Pc[j][n][J] = P[dof of node n of child j of element 0, coarse dof J of element 0]
        Pd = P.toarray()
        Pc = np.zeros((8, 27, 27))
        for j in range(8):
            Pc[j] = Pd[np.ix_(d[j], cd[0])]
        Cpat = mb.sparsity(C, order)
        Cp, Cc = Cpat[0].astype(np.int64), Cpat[1].astype(np.int32)
        Cv = np.zeros(len(Cc))
        fmask, cmask = bf.astype(np.uint8), bc.astype(np.uint8)
        emat = np.zeros((C.nel, 729))
        args_gal = [_p(cd), _p(Pc), _p(Cp), _p(Cc), _p(Cv), _p(fmask), _p(cmask), _p(emat)]
    rc = emu_sf.emu_sumfac(F.nel, F.nnode, _p(xyz), _p(conn), _p(d), _p(phi), _p(dxi), _p(deta), _p(dzeta), _p(w), _p(rp), _p(ci), _p(val), _p(u), _p(rhs),
                           1.0, 1.5, gal, *args_gal, 2)
    assert rc == 0
    assert np.abs(val - Aref.data).max() <= 1e-12 * np.abs(Aref.data).max()
    mag = np.abs(Aref) @ np.abs(u) + np.abs(rhs_ref)
    assert np.all(np.abs(rhs - rhs_ref) <= 1e-12 * mag.max())
    if gal:
        Pz = mb.zero_dirichlet(P, mb.bdc_flags(F, order), mb.bdc_flags(C, order))
        want = (Pz.T @ Aref @ Pz).toarray()
        got = sp.csr_matrix((Cv, Cc, Cp), shape=(nc, nc)).toarray()
        assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max()
        # the recorded element matrices: un-masked on the coarse side, they sum to P^T A P with only the FINE Dirichlet rows zeroed
        Pf = P.copy().tolil()
        Pf[np.nonzero(bf)[0], :] = 0
        Pf = Pf.tocsr()
        want_e = (Pf.T @ Aref @ Pf).toarray()
        acc = np.zeros((nc, nc))
        for E in range(C.nel):
            np.add.at(acc, (cd[E][:, None], cd[E][None, :]), emat[E].reshape(27, 27))
        assert np.abs(acc - want_e).max() <= 1e-12 * np.abs(want_e).max()
