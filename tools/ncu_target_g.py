"""Target for `ncu --profile-from-start off`: ONE launch of the table-driven assembly kernel on the refined tetrahedral
cube (Tet10 and Tet15 unknowns) bracketed by cudaProfilerStart/Stop."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from femus_b200 import capi, hostapi
from femus_b200.poisson import PoissonMG

cudart = None
for name in ("libcudart.so", "libcudart.so.12", "/usr/local/cuda/lib64/libcudart.so"):
    try:
        cudart = ctypes.CDLL(name)
        break
    except OSError:
        pass
assert cudart is not None, "libcudart not found"
ctx = capi.Context(0)
lev = int(sys.argv[1]) if len(sys.argv) > 1 else 5
H = hostapi.HostHierarchy.from_neu(os.path.join(ROOT, "tests", "golden", "cube_tet10.neu"), lev)
pts = [PoissonMG(ctx, 0, 0, 0, lev, fam, hier=H) for fam in ("quadratic", "biquadratic")]
for pt in pts:
    pt.KK[-1].zero(); pt.RES.zero(); pt.plans[0][1].poisson(pt.SOL, pt.RES, 1.0, 1.0)
ctx.sync()
cudart.cudaProfilerStart()
for pt in pts:
    pt.plans[0][1].poisson(pt.SOL, pt.RES, 1.0, 1.0)
ctx.sync()
cudart.cudaProfilerStop()
print("ncu target done; launches", ctx.launches())
