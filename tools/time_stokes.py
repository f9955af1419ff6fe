"""Timing of the Stokes / Navier-Stokes assembly kernels (stokes_kernel, ns_kernel: register accumulators + slot maps)
on a Q2-Q1 box and on the refined tetrahedral cube (P2-P1, BASELINE config 4's element), with the bytes the scatter
moves; one JSON line per measurement.

    python tools/time_stokes.py [n0=8] [tet_levels=5]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from femus_b200 import capi, hostapi
from femus_b200.stokes import StokesMG

n0 = int(sys.argv[1]) if len(sys.argv) > 1 else 8
tet_levels = int(sys.argv[2]) if len(sys.argv) > 2 else 5
ctx = capi.Context(0)


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    ctx.sync()
    ctx.timer_start()
    for _ in range(reps):
        fn()
    return ctx.timer_stop_ms() / reps


def run(tag, H, order_v, nv, np_):
    """the assembly plans alone (no smoother objects: their index sets are host work that grows with the mesh)"""
    top = H.levels[-1]
    fams = [order_v] * 3 + ["linear"]
    S = hostapi.SystemOnLevel(top, fams)
    pat = S.sparsity()
    A = ctx.csr(S.n, S.n, *pat)
    mesh = capi.Mesh(ctx, top.xyz, np.ascontiguousarray(top.conn))
    edofs = np.ascontiguousarray(S.elem_dofs())
    SOL, RES = ctx.vector(S.n), ctx.vector(S.n)
    SOL.put(0.1 * np.sin(np.arange(S.n) * 0.01))
    nel, nnz = top.nel, int(pat[0][-1])
    t_zero = timed(lambda: (RES.zero(), A.zero()))
    for eq in ("stokes", "navier_stokes"):
        plan = capi.StokesAssembler(mesh, A, edofs, hostapi.elem_tables(top.elem_type, order_v), hostapi.elem_tables(top.elem_type, "linear"),
                                    navier_stokes=(eq == "navier_stokes"))
        fn = (lambda: plan.assemble_ns(SOL, RES, 0.1)) if eq == "navier_stokes" else (lambda: plan.assemble(SOL, RES, 0.1))
        ms = timed(fn)
        blocks = (3 if eq == "stokes" else 9) * nv * nv + 6 * nv * np_
        ndof = 3 * nv + np_
        print(json.dumps({"kernel": eq + "_assembly", "workload": f"{tag}, {nel} elements, {S.n} rows, {nnz} non-zeros", "ms_kernel": ms,
                          "ms_zero_fill_of_matrix_and_residual": t_zero, "element_dof_updates_per_s": nel * ndof / ms * 1e3,
                          "atomics_per_element": blocks + ndof, "scatter_GBs_rmw": nel * (blocks + ndof) * 16 / ms / 1e6,
                          "slot_map_bytes": nel * blocks * 2}), flush=True)
        del plan


run(f"{n0 * 4}^3 Q2-Q1 box", hostapi.HostHierarchy(n0, n0, n0, 3), "biquadratic", 27, 8)
path = os.path.join(ROOT, "tests", "golden", "cube_tet10.neu")
run(f"tetrahedral cube refined {tet_levels - 1} times, P2-P1", hostapi.HostHierarchy.from_neu(path, tet_levels), "quadratic", 10, 4)

# One Newton step of the lid-driven cavity (BASELINE config 4's problem on Q2-Q1 hexahedra): Vanka blocks of `be` elements
# with ILU(0) block solves (the reference's ILU_PRECOND inside ASM), pressure pinned on the coarsest level + null space
# removed above; phases by device timer
if len(sys.argv) > 3:
    nc, be = int(sys.argv[3]), int(sys.argv[4]) if len(sys.argv) > 4 else 8
    import time
    t0 = time.time()
    pn = StokesMG(ctx, hostapi.HostHierarchy(nc, nc, nc, 3), IRe=0.5, equation="navier_stokes", block_elems=be, block_sub="ilu",
                  fix_pressure_at_one_point=True)
    setup_s = time.time() - t0
    sol0 = np.zeros(pn.n)
    lid = pn.sys[-1].bdc([(6,), (), (), ()]) < 1.5          # unit U on the lid (boundary set 6)
    sol0[lid] = 1.0
    pn.SOL.put(sol0)
    res, phases = [], {}
    for it in range(4):
        pn.EPS.zero()
        for name, fn in (("assemble", pn.assemble), ("galerkin_ptap", pn.galerkin), ("level_setup_ilu_factor", pn.mg_set_levels)):
            ctx.sync(); ctx.timer_start(); fn(); phases.setdefault(name, []).append(ctx.timer_stop_ms())
            if name == "assemble":
                res.append(pn.residual_norm())
        ctx.sync(); ctx.timer_start()
        for _ in range(2):
            pn.mg_solve()
        phases.setdefault("two_vcycles", []).append(ctx.timer_stop_ms())
        pn.SOL.axpy(1.0, pn.EPS)
    print(json.dumps({"kernel": "ns_newton_step_lid_driven_cavity", "workload": f"{nc * 4}^3 Q2-Q1, {pn.n} rows, Vanka blocks of {be} elements, ILU(0) block solves",
                      "setup_s": setup_s, "residuals": res, "phases_ms": {k: float(np.median(v)) for k, v in phases.items()},
                      "block_factor_bytes": [s.nbytes for s in pn.schwarz[1:]]}), flush=True)
