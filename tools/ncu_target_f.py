"""Target for `ncu --profile-from-start off`: ONE launch of the kernels of the widened rows -- the Stokes and Navier-Stokes
assembly kernels (Q2-Q1 box, P2-P1 tetrahedra) and one application of the element-block smoother with SSOR and with
ILU(0) block solves (staged row walk; 8 launches each, one per colour) -- bracketed by cudaProfilerStart/Stop.

    ncu --set full --clock-control none --import-source on --profile-from-start off \
        -o gpurun_out/prof_f python tools/ncu_target_f.py [n0_stokes=8] [n0_schwarz=8]
"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from femus_b200 import capi, hostapi
from femus_b200.poisson import PoissonMG

ns0 = int(sys.argv[1]) if len(sys.argv) > 1 else 8
nb0 = int(sys.argv[2]) if len(sys.argv) > 2 else 8
cudart = None
for name in ("libcudart.so", "libcudart.so.12", "/usr/local/cuda/lib64/libcudart.so"):
    try:
        cudart = ctypes.CDLL(name)
        break
    except OSError:
        pass
assert cudart is not None, "libcudart not found"
ctx = capi.Context(0)


def stokes_plans(H, order_v):
    top = H.levels[-1]
    S = hostapi.SystemOnLevel(top, [order_v] * 3 + ["linear"])
    A = ctx.csr(S.n, S.n, *S.sparsity())
    mesh = capi.Mesh(ctx, top.xyz, np.ascontiguousarray(top.conn))
    edofs = np.ascontiguousarray(S.elem_dofs())
    SOL, RES = ctx.vector(0.1 * np.sin(np.arange(S.n) * 0.01)), ctx.vector(S.n)
    tv, tp = hostapi.elem_tables(top.elem_type, order_v), hostapi.elem_tables(top.elem_type, "linear")
    return (A, mesh, SOL, RES, capi.StokesAssembler(mesh, A, edofs, tv, tp), capi.StokesAssembler(mesh, A, edofs, tv, tp, navier_stokes=True))


cases = [stokes_plans(hostapi.HostHierarchy(ns0, ns0, ns0, 3), "biquadratic"),
         stokes_plans(hostapi.HostHierarchy.from_neu(os.path.join(ROOT, "tests", "golden", "cube_tet10.neu"), 4), "quadratic")]
for A, mesh, SOL, RES, ps, pn in cases:          # warm-up
    ps.assemble(SOL, RES, 0.1)
    pn.assemble_ns(SOL, RES, 0.1)
sm = []
for sub in ("ssor", "ilu"):
    pb = PoissonMG(ctx, nb0, nb0, nb0, 4, "biquadratic", smoother="asm", asm_block_elems=8, asm_schedule="colours", asm_sub=sub, omega=1.0)
    pb.assemble(); pb.galerkin(); pb.mg_set_levels()
    r, y = ctx.vector(np.sin(np.arange(pb.n) * 0.001)), ctx.vector(pb.n)
    pb.schwarz[3].apply(r, y)
    sm.append((pb, r, y))
ctx.sync()

cudart.cudaProfilerStart()
for A, mesh, SOL, RES, ps, pn in cases:
    ps.assemble(SOL, RES, 0.1)
    pn.assemble_ns(SOL, RES, 0.1)
for pb, r, y in sm:
    pb.schwarz[3].apply(r, y)
ctx.sync()
cudart.cudaProfilerStop()
print("ncu target done; launches", ctx.launches())
