"""Times the Hex27 assembly kernels on the bench workload (128^3, 4 levels): fused / plain, per variant and CTA size.
    python tools/time_asm.py [n0] [levels]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from femus_b200 import capi
from femus_b200.poisson import PoissonMG

n0 = int(sys.argv[1]) if len(sys.argv) > 1 else 16
nl = int(sys.argv[2]) if len(sys.argv) > 2 else 4
ctx = capi.Context(0)
pb = PoissonMG(ctx, n0, n0, n0, nl, "biquadratic")
pb.step()
ctx.sync()


def timed(fn, reps=5):
    ts = []
    for _ in range(reps):
        pb.KK[-1].zero(); pb.RES.zero()
        ctx.sync()
        ctx.profile_only(pb.asm)
        fn()
        ctx.sync()
        nk, ms = ctx.profile_read(pb.asm)
        ctx.profile_only(None)
        ts.append(ms / max(nk, 1))
    return float(np.median(ts)), float(min(ts))


for variant, warps in ((3, 12), (3, 16), (1, 12)):
    ctx.set_option("asm_variant", variant)
    ctx.set_option("asm_warps", warps)
    plain = timed(lambda: pb.asm.poisson(pb.SOL, pb.RES, 1.0, 1.0))
    fused = timed(lambda: pb.asm.poisson_galerkin(pb.gal[-1], pb.SOL, pb.RES, 1.0, 1.0))
    print(json.dumps({"asm_variant": variant, "asm_warps": warps, "plain_ms_median_min": plain, "fused_ms_median_min": fused,
                      "elements": pb.nel}))
ctx.set_option("asm_variant", 3)
ctx.set_option("asm_warps", 12)
