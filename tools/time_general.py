"""Timing of the table-driven assembly kernel (assemble_general_kernel) on the refined tetrahedral cube of the reference
(cube_Tet: 105 elements, refined TET_LEVELS - 1 times): Tet10 and Tet15 unknowns; one JSON line each (zero fill of the
matrix and the residual included, as in the round-1 timing)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from femus_b200 import capi, hostapi
from femus_b200.poisson import PoissonMG

ctx = capi.Context(0)


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    ctx.sync()
    ctx.timer_start()
    for _ in range(reps):
        fn()
    return ctx.timer_stop_ms() / reps


lev = int(sys.argv[1]) if len(sys.argv) > 1 else 5
H = hostapi.HostHierarchy.from_neu(os.path.join(ROOT, "tests", "golden", "cube_tet10.neu"), lev)
for fam in ("linear", "quadratic", "biquadratic"):
    pt = PoissonMG(ctx, 0, 0, 0, lev, fam, hier=H)
    ms0 = timed(lambda: (pt.RES.zero(), pt.KK[-1].zero()))
    ms = timed(lambda: (pt.RES.zero(), pt.KK[-1].zero(), pt.plans[0][1].poisson(pt.SOL, pt.RES, 1.0, 1.0)))
    print(json.dumps({"kernel": "assemble_general_kernel", "workload": f"tet {fam} ({pt.nve} dofs) {pt.nel} elements", "ms": ms, "ms_zero_fill": ms0,
                      "element_dof_updates_per_s": pt.nel * pt.nve / ms * 1e3, "dofs": pt.n}), flush=True)
    del pt
