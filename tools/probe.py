"""Ad-hoc timing probe of the hot path phases on one GPU (development aid, not the bench)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from femus_b200 import capi
from femus_b200.poisson import PoissonMG

n0 = int(sys.argv[1]) if len(sys.argv) > 1 else 16
nl = int(sys.argv[2]) if len(sys.argv) > 2 else 4
order = sys.argv[3] if len(sys.argv) > 3 else "biquadratic"
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
ctx = capi.Context(0)
t = time.time()
pb = PoissonMG(ctx, n0, n0, n0, nl, order)
ctx.sync()
print(f"setup {time.time()-t:.1f}s  dofs {pb.ndofs}  nnz {[A.nnz for A in pb.KK]}  dev bytes {ctx.bytes_in_use()/1e9:.2f} GB", flush=True)


def timed(f, name, n=reps):
    ts = []
    for _ in range(n):
        ctx.timer_start()
        f()
        ts.append(ctx.timer_stop_ms())
    print(f"{name:14s} ms: " + " ".join(f"{x:9.3f}" for x in ts), flush=True)
    return min(ts)

ta = timed(pb.assemble, "assemble")
tg = timed(pb.galerkin, "galerkin")
ts = timed(pb.mg_set_levels, "mg_set_levels")
pb.assemble(); pb.galerkin(); pb.mg_set_levels()
tv = timed(pb.mg_solve, "mg_solve")
print("coarse its", pb.mg.coarse_iterations())
pb.assemble(); pb.galerkin(); pb.mg_set_levels(); pb.EPS.zero()
tr = []
for i in range(6):
    pb.mg_solve()
    tr.append(pb.residual_norm())
print("residual trace", ["%.6e" % v for v in tr])
x = ctx.vector(np.sin(np.arange(pb.n) * 0.001)); y = ctx.vector(pb.n)
A = pb.KK[-1]
tsp = timed(lambda: A.spmv(x, y), "spmv fine", 10)
print(f"spmv: {pb.spmv_bytes()/tsp/1e6:.1f} GB/s algorithmic ({pb.spmv_bytes()/1e9:.3f} GB);  assembly {pb.nel*pb.nve/ta/1e3:.3e} elem-DOF/s")
for l in range(nl - 1):
    Al = pb.KK[l]
    xx = ctx.vector(Al.shape[0]); yy = ctx.vector(Al.shape[0])
    timed(lambda: Al.spmv(xx, yy), f"spmv L{l}", 5)
if nl > 1:
    P = pb.PP[-1]
    xc = ctx.vector(P.shape[1]); yf = ctx.vector(P.shape[0])
    timed(lambda: P.spmv(xc, yf), "P spmv", 5)
# explicit teardown: library objects must go before the context (and before interpreter shutdown)
del x, y, A
try:
    del Al, xx, yy, P, xc, yf
except NameError:
    pass
del pb
ctx.close()

