"""ncu target: one y=Ax and one Jacobi sweep on the finest-level bench matrix for the SpMV variants in argv[4:]."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from femus_b200 import capi
from femus_b200.poisson import PoissonMG
n0, nl, order = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
variants = [int(v) for v in sys.argv[4:]] or [2]
cudart = ctypes.CDLL("libcudart.so.12") if os.path.exists("/usr/local/cuda/lib64/libcudart.so.12") else ctypes.CDLL("libcudart.so")
ctx = capi.Context(0)
pb = PoissonMG(ctx, n0, n0, n0, nl, order)
pb.assemble(); ctx.sync()
A = pb.KK[-1]; n = pb.n
rng = np.random.default_rng(0)
x = ctx.vector(rng.standard_normal(n)); b = ctx.vector(rng.standard_normal(n)); dinv = ctx.vector(rng.random(n) + 0.5)
y = ctx.vector(n)
for v in variants:
    ctx.set_option("spmv_variant", v)
    A.spmv(x, y); A.jacobi_sweep(dinv, b, x, y, 0.5); ctx.sync()
    cudart.cudaProfilerStart()
    A.spmv(x, y)
    A.jacobi_sweep(dinv, b, x, y, 0.5)
    ctx.sync()
    cudart.cudaProfilerStop()
print("done")
