"""Element-block (ASM / Vanka) smoother, CPU side: the host layer's index sets (MeshASMPartitioning::DoPartition,
LinearEquationSolverPetscAsm::BuildASMIndex) bit-exact against the oracle's literal restatement, the sweep
schedules (dependency levels = the reference's sequential sweep, colours), and the oracle smoother itself."""
import os
import numpy as np
import pytest
import scipy.sparse as sp

from femus_b200 import hostapi
from oracle import asm, mesh_box as mb, mesh_mixed as mm, mg

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _same(ix, be, rng, loc, ovl):
    assert ix.nblocks == len(be) and ix.block_type_range.tolist() == list(rng)
    for a, b in zip(ix.blocks("elem"), be):
        assert a.tolist() == list(b)
    for a, b in zip(ix.blocks("local"), loc):
        assert np.array_equal(a, b)
    for a, b in zip(ix.blocks("overlap"), ovl):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("nprocs", [1, 2, 3])
@pytest.mark.parametrize("order", ["linear", "quadratic", "biquadratic"])
def test_index_sets_bit_exact_on_boxes(nprocs, order):
    """Blocks of 1, 5, 8, 64 and 'all' elements on two refined levels of a 2x3x3 box, every rank of 1, 2 and 3-rank
    slab partitions (ghost dofs of other ranks end up in the overlapping sets only)."""
    H = hostapi.HostHierarchy(2, 3, 3, 3, nprocs=nprocs)
    lv = mb.build_hierarchy(2, 3, 3, 3, nprocs=nprocs)
    for l in (1, 2):
        ed = mb.system_dof(lv[l], order)
        for nb in (1, 5, 8, 64, 10 ** 6):
            for iproc in range(nprocs):
                ix = hostapi.AsmIndex(H.levels[l], order, nb, iproc)
                _same(ix, *asm.level_blocks(lv[l], ed, mb.FAMILY[order], nb, iproc))
                if iproc > 0 and nb == 8:        # shared nodes belong to the lowest rank: ranks > 0 see ghosts
                    d0, d1 = lv[l].dof_offset[mb.FAMILY[order]][iproc:iproc + 2]
                    assert any(((b < d0) | (b >= d1)).any() for b in ix.blocks("overlap"))
                    assert all(((b >= d0) & (b < d1)).all() for b in ix.blocks("local"))
    # every owned dof is in exactly one local set
    ix = hostapi.AsmIndex(H.levels[2], order, 8, 0)
    allloc = np.concatenate(ix.blocks("local"))
    assert len(np.unique(allloc)) == len(allloc) == lv[2].dof_offset[mb.FAMILY[order]][1]


@pytest.mark.parametrize("name", ["cube_mixed_3groups", "cube_mixed", "cube_tet10", "cube_wedge18"])
def test_index_sets_bit_exact_on_unstructured_meshes(name):
    """Mixed element types and two material classes (cube_mixed_3groups: 7 solid + 13 fluid elements on the coarse
    level => solid blocks first, _blockTypeRange = [ns, ns, ns + nf])."""
    path = os.path.join(GOLDEN, name + ".neu")
    H = hostapi.HostHierarchy.from_neu(path, 2)
    lv = mm.build_hierarchy(path, 2)
    for l in (0, 1):
        for order in ("linear", "quadratic", "biquadratic"):
            ed = mm.element_dofs(lv[l], order)
            for nb in (1, 3, 8):
                ix = hostapi.AsmIndex(H.levels[l], order, nb)
                be, rng, loc, ovl = asm.level_blocks(lv[l], ed, mm.FAMILY[order], nb)
                _same(ix, be, rng, loc, ovl)
                if name == "cube_mixed_3groups":
                    assert 0 < rng[0] == rng[1] < rng[2]


@pytest.mark.parametrize("nprocs", [1, 2])
def test_vanka_index_sets_of_a_velocity_pressure_system(nprocs):
    """Three triquadratic velocities + a trilinear pressure numbered [rank][variable][dof] (KKoffset), the pressure a
    Schur variable: blocks = velocities of one layer of near elements + pressures of the block's own elements
    (BuildASMIndex with NSchurVar = 1, FastVankaBlock == false); also 0 and 2 Schur variables.  Bit-exact against
    the oracle's literal loops on every rank."""
    lv = mb.build_hierarchy(2, 2, 3, 2, nprocs=nprocs)
    H = hostapi.HostHierarchy(2, 2, 3, 2, nprocs=nprocs)
    L = lv[1]
    fams, fi = ["biquadratic"] * 3 + ["linear"], [2, 2, 2, 0]
    assert np.array_equal(hostapi.system_offsets(H.levels[1], fams), asm.kk_offsets(L, fi))
    edq, edl = mb.system_dof(L, "biquadratic"), mb.system_dof(L, "linear")
    near = asm.near_elements(L.conn[:, :8])
    assert near[0][0] == 0 and near[0][1:] == sorted(near[0][1:])
    for ip in range(nprocs):
        for nb in (1, 8, 5):
            be, _, _, _ = asm.level_blocks(L, edq, 2, nb, ip)
            for ns in (1, 0, 2):
                loc, ovl = asm.build_asm_index_system(L, [edq, edq, edq, edl], fi, ns, be, near, ip)
                ix = hostapi.AsmIndex(H.levels[1], fams, nb, ip, nschur=ns)
                assert ix.nblocks == len(be)
                for a, b in zip(ix.blocks("local"), loc):
                    assert np.array_equal(a, b)
                for a, b in zip(ix.blocks("overlap"), ovl):
                    assert np.array_equal(a, b)
    with pytest.raises(ValueError):
        hostapi.AsmIndex(H.levels[1], fams, 8, 0, nschur=5)


def test_bad_arguments_fail_loudly():
    H = hostapi.HostHierarchy(2, 2, 2, 2)
    with pytest.raises(ValueError):
        hostapi.AsmIndex(H.levels[1], "linear", 0)
    with pytest.raises(ValueError):
        hostapi.AsmIndex(H.levels[1], "linear", 8, iproc=3)
    ix = hostapi.AsmIndex(H.levels[1], "linear", 8)
    rp, ci = H.levels[1].sparsity("linear")
    with pytest.raises(ValueError):
        hostapi.asm_schedule(rp[:5], ci, ix.overlap_ptr, ix.overlap, "levels")      # dofs outside the operator


def _independent(A, blocks, members):
    """no block of `members` writes what another one reads or writes"""
    pat = sp.csr_matrix(A)
    written = {}
    for b in members:
        for d in blocks[b]:
            assert written.setdefault(int(d), b) == b
    for b in members:
        cols = np.unique(np.concatenate([pat.indices[pat.indptr[r]:pat.indptr[r + 1]] for r in blocks[b]]))
        for c in cols:
            assert written.get(int(c), b) == b


@pytest.mark.parametrize("order", ["linear", "biquadratic"])
def test_schedules_reproduce_the_sequential_sweep(order):
    """levels: groups equal the oracle's dependency levels, blocks of a group are independent, and the sweep in
    (group, block) order equals the sequential sweep BIT FOR BIT; colours: independent inside a colour, 8 colours on
    2x2x2-element blocks of a box, and the coloured sweep equals the oracle's sweep over the stably sorted list."""
    lv = mb.build_hierarchy(2, 2, 2, 3)
    H = hostapi.HostHierarchy(2, 2, 2, 3)
    ix = hostapi.AsmIndex(H.levels[2], order, 8)
    blocks = ix.blocks()
    O = mg.Hierarchy(lv, order, smoother="asm", asm_blocks=[None, hostapi.AsmIndex(H.levels[1], order, 8).blocks(), blocks])
    A = O.A[2]
    rp, ci = H.levels[2].sparsity(order)
    assert np.array_equal(rp, A.indptr) and np.array_equal(ci, A.indices)
    grp, gptr, gblocks = hostapi.asm_schedule(rp, ci, ix.overlap_ptr, ix.overlap, "levels")
    assert np.array_equal(grp, asm.schedule(A, blocks))
    r = np.random.default_rng(1).standard_normal(A.shape[0])
    S = O.asm[2]
    y_seq = S.apply(r)
    for g in range(len(gptr) - 1):
        _independent(A, blocks, gblocks[gptr[g]:gptr[g + 1]].tolist())
    assert np.array_equal(S.apply(r, gblocks), y_seq)
    assert np.array_equal(S.apply(r, gblocks[::-1][np.argsort(grp[gblocks[::-1]], kind="stable")]), y_seq)   # any order inside a group
    grp2, gptr2, gblocks2 = hostapi.asm_schedule(rp, ci, ix.overlap_ptr, ix.overlap, "colours")
    assert len(gptr2) - 1 == 8 and np.array_equal(np.bincount(grp2), np.full(8, 8))
    for g in range(8):
        _independent(A, blocks, gblocks2[gptr2[g]:gptr2[g + 1]].tolist())
    y_col = S.apply(r, gblocks2)
    assert np.abs(y_col - y_seq).max() > 1e-6 * np.abs(y_seq).max()         # another sweep order: another preconditioner
    assert np.abs(A @ y_col - r).max() < np.abs(r).max()                      # ... and still a contraction of the residual


def test_oracle_block_smoother_properties():
    """One block holding every dof: M^-1 is the exact inverse.  Blocks = single dofs in order: the sweep is forward
    Gauss-Seidel.  In the V-cycle the block smoother beats Richardson + Jacobi by orders of magnitude."""
    lv = mb.build_hierarchy(2, 2, 2, 2)
    H0 = mg.Hierarchy(lv, "linear")
    A = H0.A[1]
    n = A.shape[0]
    r = np.random.default_rng(2).standard_normal(n)
    assert np.abs(A @ asm.BlockSmoother(A, [np.arange(n)]).apply(r) - r).max() < 1e-12
    y = asm.BlockSmoother(A, [np.array([i]) for i in range(n)]).apply(r)
    Lw = sp.tril(A).tocsr()
    import scipy.sparse.linalg as spla
    assert np.abs(y - spla.spsolve_triangular(Lw, r, lower=True)).max() < 1e-12
    ed = mb.system_dof(lv[1], "linear")
    _, _, _, ovl = asm.level_blocks(lv[1], ed, 0, 8)
    Ha = mg.Hierarchy(lv, "linear", smoother="asm", asm_blocks=[None, ovl])
    ta, _ = Ha.mg_solve_trace(4, omega=1.0)
    tj, _ = H0.mg_solve_trace(4)
    assert ta[-1] < 1e-3 * tj[-1]
    Hs = mg.Hierarchy(lv, "linear", smoother="asm", asm_blocks=[None, ovl], asm_sub="ssor")
    ts, _ = Hs.mg_solve_trace(4, omega=1.0)
    assert ts[-1] < tj[-1]


def test_oracle_gmres_level_solver_minimises_the_preconditioned_residual():
    """oracle.mg.Hierarchy.gmres (KSPGMRES restated: left preconditioning, modified Gram-Schmidt, Givens rotations)
    returns the least-squares minimiser of ||M^-1 (b - A x)|| over x0 + K_k(M^-1 A, M^-1 r0), for the Jacobi and for the
    one-block ILU(0) preconditioner (= the reference's default level solver); as a smoother it beats Richardson."""
    lv = mb.build_hierarchy(2, 2, 2, 2)
    H = hostapi.HostHierarchy(2, 2, 2, 2)
    blocks = [None, hostapi.AsmIndex(H.levels[1], "linear", 10 ** 6).blocks()]
    rng = np.random.default_rng(3)
    for O in (mg.Hierarchy(lv, "linear", ksp="gmres"), mg.Hierarchy(lv, "linear", smoother="asm", asm_blocks=blocks, asm_sub="ilu", ksp="gmres")):
        A = O.A[1]
        b, x0 = rng.standard_normal(A.shape[0]), rng.standard_normal(A.shape[0])
        for k in (1, 3):
            xk = O.gmres(1, x0, b, k)
            z0 = O.pc_apply(1, b - A @ x0)
            K = [z0]
            for _ in range(k - 1):
                K.append(O.pc_apply(1, A @ K[-1]))
            K = np.array(K).T
            MAK = np.array([O.pc_apply(1, A @ K[:, j]) for j in range(k)]).T
            c = np.linalg.lstsq(MAK, z0, rcond=None)[0]
            assert np.abs(xk - (x0 + K @ c)).max() <= 1e-11 * np.abs(xk).max()
    tg, _ = mg.Hierarchy(lv, "linear", ksp="gmres").mg_solve_trace(4)
    tr, _ = mg.Hierarchy(lv, "linear").mg_solve_trace(4)
    assert tg[-1] < 0.1 * tr[-1]


@pytest.mark.parametrize("case", ["box322_3lev", "cube_mixed_3groups_2lev"])
def test_element_blocks_match_the_reference(case):
    """REFERENCE OUTPUT: tests/golden/ref_partition.json holds what the reference's own MeshASMPartitioning::DoPartition
    returned on every level of a refined HEX27 box and of the mixed mesh with three material groups (7 solid + 13 fluid
    coarse elements), for several block sizes, with every element's material flag in the reference's element order
    (tests/cpp/ref_partition.cpp on the host backend of oracle/ref_build; tests/golden/make_ref_stokes_golden.py).  The
    oracle's restatement and the product's host layer (AsmPartition.hpp) reproduce element blocks and block type ranges
    bit-exactly -- with the element numbering of the refined levels, which is what the blocks are made of -- and the
    oracle the reference's near-element lists (elem::BuildElementNearElement), from which Vanka blocks take their
    velocity dofs."""
    import json
    ref = json.load(open(os.path.join(GOLDEN, "ref_partition.json")))[case]
    if case.startswith("box"):
        n, nl = [int(v) for v in ref["args"][1:4]], int(ref["args"][4])
        H, lv = hostapi.HostHierarchy(*n, nl), mb.build_hierarchy(*n, nl)
    else:
        path, nl = os.path.join(GOLDEN, os.path.basename(ref["args"][1])), int(ref["args"][2])
        H, lv = hostapi.HostHierarchy.from_neu(path, nl), mm.build_hierarchy(path, nl)
    assert len(ref["levels"]) == nl
    for l, R in enumerate(ref["levels"]):
        assert R["nel"] == lv[l].nel
        material = getattr(lv[l], "material", None)
        if material is None or len(material) == 0:
            material = np.full(lv[l].nel, 2)
        assert np.array_equal(R["material"], material), f"level {l}: material flags in element order"
        # one layer of near elements (elem::BuildElementNearElement): the element itself, then its vertex neighbours ascending
        if case.startswith("box"):
            verts = lv[l].conn[:, :8]
        else:
            verts = [lv[l].conn[e, :mm.NVE[int(lv[l].etype[e])][0]] for e in range(lv[l].nel)]
        assert [list(map(int, x)) for x in asm.near_elements(verts)] == R["near"], f"level {l}: near elements"
        for part in R["partitions"]:
            bs = part["block_size"]
            be, rng = asm.do_partition(material, lv[l].elem_offset, 0, (bs, bs, bs))
            assert [list(map(int, b)) for b in be] == part["blocks"] and list(rng) == part["block_type_range"], f"oracle: level {l} block size {bs}"
            ix = hostapi.AsmIndex(H.levels[l], "linear", min(bs, lv[l].nel))        # LinearImplicitSystem.cpp:1198 clamps the block size
            want = part["blocks"]
            assert [b.tolist() for b in ix.blocks("elem")] == want, f"host layer: level {l} block size {bs}"
            assert ix.block_type_range.tolist() == part["block_type_range"]
