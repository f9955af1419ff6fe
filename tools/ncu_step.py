"""ncu target: ONE step of the hot path (assembly fused with the finest Galerkin product, Galerkin chain,
MGSetLevel on every level, one V-cycle) after warm-up, bracketed by cudaProfilerStart/Stop.

    ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
        --log-file gpurun_out/launches.csv python tools/ncu_step.py 16 4 biquadratic
"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from femus_b200 import capi
from femus_b200.poisson import PoissonMG
n0, nl, order = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
cudart = ctypes.CDLL("libcudart.so.12") if os.path.exists("/usr/local/cuda/lib64/libcudart.so.12") else ctypes.CDLL("libcudart.so")
ctx = capi.Context(0)
pb = PoissonMG(ctx, n0, n0, n0, nl, order)
pb.step(); pb.step(); ctx.sync()
cudart.cudaProfilerStart()
pb.step()
ctx.sync()
cudart.cudaProfilerStop()
print("done", ctx.launches())
