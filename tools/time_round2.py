"""Timing of the kernels written after round 1's GPU budget was spent (run on a B200 at the start of round 2):
the element-block smoother (exact and SSOR block solves, coloured sweep) on the finest level of an n^3 HEX27 box
hierarchy -- ms per application, algorithmic GB/s against the measured HBM peak, V-cycle contraction next to
Richardson + Jacobi -- and the table-driven assembly kernel on a refined tetrahedral mesh.

    python tools/time_round2.py [n0=8] [levels=4] [order=biquadratic]      # one JSON line per measurement
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from femus_b200 import capi, hostapi
from femus_b200.poisson import PoissonMG

n0 = int(sys.argv[1]) if len(sys.argv) > 1 else 8
nl = int(sys.argv[2]) if len(sys.argv) > 2 else 4
order = sys.argv[3] if len(sys.argv) > 3 else "biquadratic"
peak = None
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))).get("hbm_gbs")
except Exception:
    pass
ctx = capi.Context(0)


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    ctx.sync()
    ctx.timer_start()
    for _ in range(reps):
        fn()
    return ctx.timer_stop_ms() / reps


for sub in ("lu", "ssor", "ilu"):
    pb = PoissonMG(ctx, n0, n0, n0, nl, order, smoother="asm", asm_block_elems=8, asm_schedule="colours", asm_sub=sub, omega=1.0)
    pb.assemble(); pb.galerkin(); pb.mg_set_levels()
    top = nl - 1
    S, ix = pb.schwarz[top], pb.asm_index[top]
    n = pb.n
    r, y = ctx.vector(np.sin(np.arange(n) * 0.001)), ctx.vector(n)
    ms = timed(lambda: S.apply(r, y))
    A = pb.KK[top]
    m = np.diff(ix.overlap_ptr)
    rows_nnz = int(A.nnz * (m.sum() / n))                  # every dof sits in m.sum()/n blocks on average
    # exact: inverse + one pass over the rows; SSOR: three passes over the rows; ILU(0): one pass + the factor twice
    alg = {"lu": 8 * int((m.astype(np.int64) ** 2).sum()) + 12 * rows_nnz, "ssor": 3 * 12 * rows_nnz, "ilu": 12 * rows_nnz + 2 * 12 * rows_nnz}[sub] \
        + 24 * int(m.sum())
    trace = []
    for _ in range(4):
        pb.mg_solve()
        trace.append(pb.residual_norm())
    ms_cycle = timed(lambda: pb.mg_solve(), reps=3, warm=1)
    print(json.dumps({"kernel": "schwarz_apply_" + sub, "workload": f"{n0 * 2 ** (nl - 1)}^3 {order}", "blocks": int(ix.nblocks),
                      "groups": int(S.ngroups), "ms": ms, "algorithmic_bytes": alg, "GBs": alg / ms / 1e6, "hbm_peak_GBs": peak,
                      "frac": (alg / ms / 1e6 / peak) if peak else None, "inverse_bytes": S.nbytes, "vcycle_ms": ms_cycle,
                      "residual_trace": trace}))
    del pb, S

for ksp in ("richardson", "gmres"):
    pb = PoissonMG(ctx, n0, n0, n0, nl, order, ksp=ksp)
    pb.assemble(); pb.galerkin(); pb.mg_set_levels()
    trace = []
    for _ in range(4):
        pb.mg_solve()
        trace.append(pb.residual_norm())
    print(json.dumps({"kernel": f"vcycle_{ksp}_jacobi", "vcycle_ms": timed(lambda: pb.mg_solve(), reps=3, warm=1), "residual_trace": trace}))
    del pb

# table-driven assembly kernel on tetrahedra: the reference's cube_Tet coarse mesh (105 elements) refined
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "cube_tet10.neu")
lev = int(os.environ.get("TET_LEVELS", "5"))
H = hostapi.HostHierarchy.from_neu(path, lev)
for fam in ("quadratic", "biquadratic"):
    pt = PoissonMG(ctx, 0, 0, 0, lev, fam, hier=H)
    ms = timed(lambda: (pt.RES.zero(), pt.KK[-1].zero(), pt.plans[0][1].poisson(pt.SOL, pt.RES, 1.0, 1.0)))
    nve = pt.nve
    print(json.dumps({"kernel": "assemble_general_kernel", "workload": f"tet {fam} {pt.nel} elements", "ms": ms,
                      "element_dof_updates_per_s": pt.nel * nve / ms * 1e3, "dofs": pt.n}))
    del pt

# Stokes / Navier-Stokes assembly kernels on a Q2-Q1 box (first timing of stokes_kernel / ns_kernel)
from femus_b200.stokes import StokesMG
ns0 = int(os.environ.get("STOKES_N0", "8"))
Hs = hostapi.HostHierarchy(ns0, ns0, ns0, 3)
for eq in ("stokes", "navier_stokes"):
    ps = StokesMG(ctx, Hs, IRe=0.1, velocity_dirichlet=(1, 3, 4, 5, 6), equation=eq)
    ps.SOL.put(0.1 * np.sin(np.arange(ps.n) * 0.01))
    ms = timed(lambda: ps.assemble())
    nel = Hs.levels[-1].nel
    print(json.dumps({"kernel": eq + "_assembly", "workload": f"{ns0 * 4}^3 Q2-Q1, {ps.n} rows", "ms": ms,
                      "element_dof_updates_per_s": nel * 89 / ms * 1e3}))
    del ps
# Newton-multigrid on a small hierarchy (the dense Vanka inverses, 8.6 MB per one-element block, bound the size)
pn = StokesMG(ctx, hostapi.HostHierarchy(2, 2, 2, 3), IRe=0.1, velocity_dirichlet=(1, 3, 4, 5, 6), equation="navier_stokes")
sol0 = np.zeros(pn.n)
sol0[pn.sys[-1].bdc([(6,), (), (), ()]) < 1.5] = 1.0
pn.SOL.put(sol0)
print(json.dumps({"kernel": "ns_newton_multigrid", "rows": pn.n, "residuals": [pn.newton_step(ncycles=2) for _ in range(4)],
                  "vanka_inverse_bytes": [s.nbytes for s in pn.schwarz[1:]]}))
