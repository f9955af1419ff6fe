"""Target for `ncu --profile-from-start off`: builds the bench workload, warms it up, then brackets
ONE launch of each hot kernel (finest-level assembly, Galerkin product, y=Ax / r=b-Ax / Jacobi sweep,
P and P^T SpMV) with cudaProfilerStart/Stop so that only those are captured.

    ncu --set full --clock-control none --import-source on --profile-from-start off \
        -o gpurun_out/prof python tools/ncu_target.py 16 4 biquadratic
"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from femus_b200 import capi
from femus_b200.poisson import PoissonMG

n0 = int(sys.argv[1]) if len(sys.argv) > 1 else 16
nl = int(sys.argv[2]) if len(sys.argv) > 2 else 4
order = sys.argv[3] if len(sys.argv) > 3 else "biquadratic"
what = sys.argv[4].split(",") if len(sys.argv) > 4 else ["asm", "galerkin", "spmv", "pr"]

cudart = None
for name in ("libcudart.so", "libcudart.so.12", "/usr/local/cuda/lib64/libcudart.so"):
    try:
        cudart = ctypes.CDLL(name)
        break
    except OSError:
        pass
assert cudart is not None, "libcudart not found"

ctx = capi.Context(0)
pb = PoissonMG(ctx, n0, n0, n0, nl, order)
pb.step()
pb.step()
ctx.sync()
A = pb.KK[-1]
x = ctx.vector(np.sin(np.arange(pb.n) * 0.001))
y = ctx.vector(pb.n)
b = ctx.vector(np.cos(np.arange(pb.n) * 0.002))
dinv = ctx.vector(np.full(pb.n, 0.5))
A.spmv(x, y)
if nl > 1:
    P = pb.PP[-1]
    xc = ctx.vector(np.sin(np.arange(P.shape[1]) * 0.001))
    yf = ctx.vector(P.shape[0])
    R = P.transpose()
if "vcycle" in what:
    pb.assemble(); pb.galerkin(); pb.mg_set_levels()
ctx.sync()

cudart.cudaProfilerStart()
if "asm" in what:
    pb.KK[-1].zero()
    pb.RES.zero()
    pb.asm.poisson(pb.SOL, pb.RES, 1.0, 1.0)
if "fused" in what:
    pb.KK[-1].zero()
    pb.RES.zero()
    pb.asm.poisson_galerkin(pb.gal[-1], pb.SOL, pb.RES, 1.0, 1.0)
if "chain" in what and nl > 2:
    pb.gal[-2].apply_from_elements(pb.gal[-1])
if "galerkin" in what:
    pb.gal[-1].apply()
if "spmv" in what:
    A.spmv(x, y)
    A.resid(b, x, y)
    A.jacobi_sweep(dinv, b, x, y, 0.5)
if "pr" in what and nl > 1:
    P.spmv(xc, yf)
    R.spmv(yf, xc)
if "vcycle" in what:      # every kernel of one V-cycle, the persistent coarse PCG among them
    pb.mg_solve()
ctx.sync()
cudart.cudaProfilerStop()
print("ncu target done; launches", ctx.launches())
