"""Timing of the element-block smoother's apply kernels (exact, SSOR, ILU(0) block solves; coloured sweep over blocks
of 8 HEX27 elements) on the finest level of an n^3 box hierarchy: ms per application, algorithmic GB/s against the
measured HBM peak, V-cycle time and contraction.  One JSON line per block solve.

    python tools/time_schwarz.py [n0=8] [levels=4] [subs=ssor,ilu]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from femus_b200 import capi
from femus_b200.poisson import PoissonMG

n0 = int(sys.argv[1]) if len(sys.argv) > 1 else 8
nl = int(sys.argv[2]) if len(sys.argv) > 2 else 4
subs = sys.argv[3].split(",") if len(sys.argv) > 3 else ["ssor", "ilu"]
peak = None
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs")
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


for sub in subs:
    pb = PoissonMG(ctx, n0, n0, n0, nl, "biquadratic", smoother="asm", asm_block_elems=8, asm_schedule="colours", asm_sub=sub, omega=1.0)
    pb.assemble(); pb.galerkin()
    ms_setup = timed(lambda: pb.mg_set_levels(), reps=2, warm=1)
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
    print(json.dumps({"kernel": "schwarz_apply_" + sub, "workload": f"{n0 * 2 ** (nl - 1)}^3 biquadratic", "blocks": int(ix.nblocks),
                      "groups": int(S.ngroups), "ms": ms, "algorithmic_bytes": alg, "GBs": alg / ms / 1e6, "hbm_peak_GBs": peak,
                      "frac": (alg / ms / 1e6 / peak) if peak else None, "factor_bytes": S.nbytes, "level_setup_ms": ms_setup,
                      "vcycle_ms": ms_cycle, "residual_trace": trace}), flush=True)
    del pb, S
