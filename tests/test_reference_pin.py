"""CPU: the oracle (numpy restatement) and the product's host layer (femus_b200/host/*.hpp through hostapi) against
REFERENCE OUTPUT: tests/golden/ref_poisson_*.npz hold what the reference's own applications/001_Poisson/main.cpp --
compiled UNMODIFIED with its own Mesh / MeshRefinement / MultiLevelSolution / LinearEquation / LinearImplicitSystem
sources on the single-process host backend of oracle/ref_build -- produced on small inputs
(tests/golden/make_ref_golden.py).  Integers (node numbering, element dofs, dof offsets, KKoffset, sparsity counts and
patterns, prolongator structure, Dirichlet rows) bit-exact; coordinates, matrices, right-hand sides, prolongators and
the residual norms the reference prints (LinearImplicitSystem.cpp:426) to the tolerances stated below."""
import json
import os

import numpy as np
import pytest
import scipy.sparse as sp

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
ORDER = {"first": "linear", "serendipity": "quadratic", "second": "biquadratic"}
# boundary index of the six box faces: MeshGeneration.cpp:1038-1071 sets face-element indices -2 ... -7 and
# Elem.cpp:361-364 turns them into 1 ... 6; SetBoundaryInfo(index - 1, name) names them
FACE = {"bottom": 1, "front": 2, "right": 3, "behind": 4, "left": 5, "top": 6}

BOX_CASES = ["box222_q2_3lev", "box222_q1_3lev", "box324_q2_2lev_neumann", "box232_q1_2lev_source_xyz"]
FILE_CASES = [("cube_hex_q2_2lev", "cube_hex27_2x2x2.neu"), ("cube_tet_q2_2lev", "cube_tet10.neu"), ("cube_tet_serendipity_2lev", "cube_tet10.neu"),
              ("cube_wedge_q2_2lev", "cube_wedge18.neu"), ("cube_mixed_q2_2lev", "cube_mixed.neu"), ("cube_mixed_q1_2lev", "cube_mixed.neu")]


def load(name):
    g = np.load(os.path.join(GOLDEN, f"ref_poisson_{name}.npz"))
    spec = json.loads(str(g["input_json"]))
    nl = spec["multilevel_problem"]["multilevel_mesh"]["first"]["system"]["poisson"]["linear_solver"]["type"]["multigrid"]["nlevels"]
    var = spec["multilevel_solution"]["multilevel_mesh"]["first"]["variable"]["first"]
    return g, spec, nl, ORDER[var["fe_order"]], var


def csr(g, key):
    shape = tuple(g[f"{key}_shape"])
    return sp.csr_matrix((g[f"{key}_val"], g[f"{key}_col"], g[f"{key}_rowptr"]), shape=shape)


def ref_level(g, l):
    nel, nnode, dim, nvar, soltype, nprocs = g[f"L{l}_info"]
    return dict(nel=int(nel), nnode=int(nnode), conn=g[f"L{l}_conn"].reshape(nel, 27), etype=g[f"L{l}_etype"], xyz=g[f"L{l}_xyz"].reshape(3, nnode),
                face=g[f"L{l}_face_index"].reshape(nel, 6), sysdof=g[f"L{l}_sysdof"].reshape(nel, 27), dofoff=g[f"L{l}_dofOffset"].reshape(3, nprocs + 1),
                kkoff=g[f"L{l}_KKoffset"], bdc=g[f"L{l}_Bdc"], bdc_idx=g[f"L{l}_bdcIndex"], n_nz=g[f"L{l}_n_nz"], n_oz=g[f"L{l}_n_oz"])


def box_faces(var):
    """boundary index (1..6) of the box faces by name, as MeshGeneration.cpp:1038-1071 assigns them"""
    # every face is homogeneous Dirichlet unless the input says otherwise (InitializeBdc_with_ParsedFunction,
    # MultiLevelSolution.cpp:575-596)
    neumann = {FACE[b["facename"]]: float(b["bdc_func"]) for b in var["boundary_conditions"] if b["bdc_type"] == "neumann"}
    dirichlet = tuple(i for i in range(1, 7) if i not in neumann)
    return dirichlet, neumann


def check_mesh_level(R, conn, xyz, sysdof, nve_of_elem, rp, ci, bdc, dofoff=None):
    """one level: integers bit-exact, coordinates to 1e-15"""
    assert R["nel"] == conn.shape[0] and R["nnode"] == xyz.shape[1]
    nn = (R["conn"] >= 0).sum(axis=1)
    for e in range(R["nel"]):
        assert np.array_equal(R["conn"][e, :nn[e]], conn[e, :nn[e]]), f"element {e}: node numbering"
    assert np.abs(R["xyz"] - xyz).max() <= 1e-15 * max(1.0, np.abs(xyz).max())
    for e in range(R["nel"]):
        k = nve_of_elem[e]
        assert np.array_equal(R["sysdof"][e, :k], sysdof[e, :k]), f"element {e}: system dofs"
        assert np.all(R["sysdof"][e, k:] == -1)
    if dofoff is not None:
        assert np.array_equal(R["dofoff"], dofoff)
    # sparsity: the counts GetSparsityPatternSize handed to init() and the pattern the assembly really touched
    assert np.array_equal(R["n_nz"], np.diff(rp)) and not R["n_oz"].any()
    assert np.array_equal(np.nonzero(R["bdc"] < 1.5)[0], R["bdc_idx"])
    assert np.array_equal(R["bdc"], bdc)


@pytest.mark.parametrize("name", BOX_CASES)
def test_box_oracle_and_host_layer_match_the_reference(name):
    from femus_b200 import hostapi
    from oracle import mesh_box as mb, mg
    g, spec, nl, order, var = load(name)
    box = spec["multilevel_mesh"]["first"]["type"]["box"]
    n = (box["nx"], box["ny"], box["nz"])
    dirichlet, neumann = box_faces(var)
    lv = mb.build_hierarchy(*n, nl)
    H = hostapi.HostHierarchy(*n, nl)
    for l in range(nl):
        R = ref_level(g, l)
        A = csr(g, f"L{l}_KK") if f"L{l}_KK_val" in g else None
        rp_ref, ci_ref = g[f"L{l}_KK_rowptr"], g[f"L{l}_KK_col"]
        for who, conn, xyz, sysdof, (rp, ci), bdc, dofoff in (
                ("oracle", lv[l].conn, lv[l].xyz, mb.system_dof(lv[l], order), mb.sparsity(lv[l], order), mb.bdc_flags(lv[l], order, dirichlet),
                 np.array([o for o in lv[l].dof_offset])),
                ("host layer", H.levels[l].conn, H.levels[l].xyz, H.levels[l].system_dofs(order), H.levels[l].sparsity(order),
                 H.levels[l].bdc(order, dirichlet), None)):
            nve = np.full(R["nel"], 27 if order == "biquadratic" else 8)
            check_mesh_level(R, conn, xyz, sysdof, nve, rp, ci, bdc, dofoff)
            assert np.array_equal(rp, rp_ref) and np.array_equal(ci, ci_ref), f"{who}: level {l} sparsity pattern"
        # boundary faces: -(index + 1) on exterior faces (Elem.cpp:361-364)
        fe, fl, fb = H.levels[l].boundary_faces()
        assert np.array_equal(-(R["face"][fe, fl] + 1), fb)
        assert int((R["face"] < -1).sum()) == len(fe)
        if l > 0:
            P = csr(g, f"L{l}_PP")
            Po = mb.zero_dirichlet(mb.prolongator(lv[l - 1], lv[l], order), mb.bdc_flags(lv[l], order, dirichlet), mb.bdc_flags(lv[l - 1], order, dirichlet))
            Po = Po.tocsr()
            # the reference zeroes Dirichlet rows / columns IN the pattern (ZeroInterpolatorDirichletNodes): compare as operators
            assert abs(P - Po).max() <= 1e-15
            rph, cih, vh, shp = H.prolongator(l, order)
            assert np.array_equal(rph, P.indptr) and np.array_equal(cih, P.indices), "host layer: prolongator structure"
    # assembled system of the finest level (the application's own element loop) and the Galerkin operators below it
    top = nl - 1
    src = var["func_source"]
    if src in ("0.", "1."):
        Ao, rhs = mb.assemble(lv[top], order, None, float(src))
        if neumann:
            rhs = rhs + mb.neumann_rhs(lv[top], order, neumann)
        res_ref = g[f"L{top}_RES"]
        assert np.abs(rhs - res_ref).max() <= 1e-13 * np.abs(res_ref).max()
        if f"L{top}_KK_val" in g:
            Ar = csr(g, f"L{top}_KK")
            assert np.array_equal(Ar.indptr, Ao.indptr) and np.array_equal(Ar.indices, Ao.indices)
            assert np.abs(Ar.data - Ao.data).max() <= 1e-13 * np.abs(Ao.data).max()
        else:
            assert np.abs(np.asarray(Ao.sum(axis=1)).ravel() - g[f"L{top}_KK_rowsum"]).max() <= 1e-12 * np.abs(Ao.data).max()
            assert np.abs(np.asarray(abs(Ao).sum(axis=1)).ravel() - g[f"L{top}_KK_absrowsum"]).max() <= 1e-12 * np.abs(Ao.data).max()
            assert np.abs(Ao.diagonal() - g[f"L{top}_KK_diag"]).max() <= 1e-13 * np.abs(Ao.data).max()
        # the V-cycle the application configures: Richardson(0.5) around PCSOR (main.cpp:239-242), six cycles; the
        # reference prints the norms with 7 digits
        blocks = [None] + [[np.arange(mb.ndofs(lv[l], order))] for l in range(1, nl)]
        Hm = mg.Hierarchy(lv, order, fsrc=float(src), dirichlet_faces=dirichlet, neumann=neumann or None, smoother="asm", asm_blocks=blocks, asm_sub="ssor")
        trace, _ = Hm.mg_solve_trace(6, omega=0.5)
        ref = g["residual_trace"]
        assert len(ref) == 6 and np.all(np.abs(np.array(trace) - ref) <= 2e-6 * ref), (trace, ref)
        for l in range(top):
            if f"L{l}_KK_val" in g:
                assert abs(csr(g, f"L{l}_KK") - Hm.A_raw[l]).max() <= 1e-12 * abs(Hm.A_raw[l]).max()


@pytest.mark.parametrize("name,neu", FILE_CASES)
def test_gambit_meshes_oracle_and_host_layer_match_the_reference(name, neu):
    """The reference's own Gambit reader, AddBiquadraticNodesNotInMeshFile, refinement and numbering on its four shipped
    3-D meshes (re-serialised fixtures): tetrahedra, wedges, the mixed mesh, hexahedra; boundary conditions of the
    application's SetBoundaryCondition (main.cpp:26-37: flux 0.2 on boundary 3, Dirichlet elsewhere)."""
    from femus_b200 import hostapi
    from oracle import mesh_mixed as mm, mg
    g, spec, nl, order, var = load(name)
    path = os.path.join(GOLDEN, neu)
    lv = mm.build_hierarchy(path, nl)
    H = hostapi.HostHierarchy.from_neu(path, nl)
    groups = sorted(set(int(b) for L in lv for b in np.unique(L.boundary_faces()[2]))) if hasattr(lv[0], "boundary_faces") else None
    fe, fl, fb = H.levels[0].boundary_faces()
    dirichlet = tuple(int(b) for b in np.unique(fb) if b != 3)
    neumann = {3: 0.2}
    for l in range(nl):
        R = ref_level(g, l)
        Lh = H.levels[l]
        nve_o = np.array([hostapi.elem_nve(int(t), order) for t in lv[l].etype]) if hasattr(lv[l], "etype") else None
        nve_h = np.array([hostapi.elem_nve(int(t), order) for t in (Lh.elem_types if Lh.elem_type < 0 else np.full(Lh.nel, Lh.elem_type))])
        assert np.array_equal(R["etype"], Lh.elem_types if Lh.elem_type < 0 else np.full(Lh.nel, Lh.elem_type))
        rp_ref, ci_ref = g[f"L{l}_KK_rowptr"], g[f"L{l}_KK_col"]
        sd_h = Lh.system_dofs27(order) if Lh.elem_type < 0 else np.pad(Lh.system_dofs(order), ((0, 0), (0, 27 - Lh.system_dofs(order).shape[1])), constant_values=-1)
        check_mesh_level(R, Lh.conn, Lh.xyz, sd_h, nve_h, *Lh.sparsity(order), Lh.bdc(order, dirichlet))
        rp, ci = Lh.sparsity(order)
        assert np.array_equal(rp, rp_ref) and np.array_equal(ci, ci_ref), f"host layer: level {l} sparsity pattern"
        sd_o = mm.system_dofs27(lv[l], order)
        check_mesh_level(R, lv[l].conn, lv[l].xyz, sd_o, nve_h, *mm.sparsity(lv[l], order), mm.bdc_flags(lv[l], order, dirichlet))
        fe, fl, fb = Lh.boundary_faces()
        assert np.array_equal(-(R["face"][fe, fl] + 1), fb)
        if l > 0:
            P = csr(g, f"L{l}_PP")
            from oracle import mesh_box as mb
            Po = mb.zero_dirichlet(mm.prolongator(lv[l - 1], lv[l], order), mm.bdc_flags(lv[l], order, dirichlet), mm.bdc_flags(lv[l - 1], order, dirichlet)).tocsr()
            assert abs(P - Po).max() <= 1e-14
            rph, cih, vh, shp = H.prolongator(l, order)
            assert np.array_equal(rph, P.indptr) and np.array_equal(cih, P.indices), "host layer: prolongator structure"
    top = nl - 1
    Ao, rhs = mm.assemble(lv[top], order, None, 0.0)
    rhs = rhs + mm.neumann_rhs(lv[top], order, neumann)
    res_ref = g[f"L{top}_RES"]
    assert np.abs(rhs - res_ref).max() <= 1e-12 * np.abs(res_ref).max()
    Ar = csr(g, f"L{top}_KK")
    assert np.array_equal(Ar.indptr, Ao.indptr) and np.array_equal(Ar.indices, Ao.indices)
    assert np.abs(Ar.data - Ao.data).max() <= 1e-12 * np.abs(Ao.data).max()
    blocks = [None] + [[np.arange(mm.ndofs(lv[l], order))] for l in range(1, nl)]
    Hm = mg.Hierarchy(lv, order, fsrc=0.0, dirichlet_faces=dirichlet, neumann=neumann, smoother="asm", asm_blocks=blocks, asm_sub="ssor", mesh=mm)
    trace, _ = Hm.mg_solve_trace(6, omega=0.5)
    ref = g["residual_trace"]
    assert len(ref) == 6 and np.all(np.abs(np.array(trace) - ref) <= 2e-6 * ref), (trace, ref)
