"""SpMV variant sweep on the bench matrix (development aid): every kernel variant of the SpMV
family on the finest-level CSR, checked against the streaming kernel, GB/s of algorithmic bytes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from femus_b200 import capi
from femus_b200.poisson import PoissonMG

n0 = int(sys.argv[1]) if len(sys.argv) > 1 else 16
nl = int(sys.argv[2]) if len(sys.argv) > 2 else 4
order = sys.argv[3] if len(sys.argv) > 3 else "biquadratic"
ctx = capi.Context(0)
pb = PoissonMG(ctx, n0, n0, n0, nl, order)
pb.assemble()
ctx.sync()
A = pb.KK[-1]
n = pb.n
rng = np.random.default_rng(0)
x = ctx.vector(rng.standard_normal(n)); b = ctx.vector(rng.standard_normal(n)); dinv = ctx.vector(rng.random(n) + 0.5)
y = ctx.vector(n); yref = ctx.vector(n)
bytes_ax = pb.spmv_bytes()
ops = {"y=Ax": (lambda o: A.spmv(x, o), bytes_ax), "r=b-Ax": (lambda o: A.resid(b, x, o), bytes_ax + 8 * n),
       "jacobi": (lambda o: A.jacobi_sweep(dinv, b, x, o, 0.5), bytes_ax + 24 * n)}
ref = {}
ctx.set_option("spmv_variant", 0)
for name, (f, _) in ops.items():
    f(yref); ref[name] = yref.get()
for var in (0, 1, 2):
    ctx.set_option("spmv_variant", var)
    for name, (f, nbytes) in ops.items():
        f(y)
        err = np.abs(y.get() - ref[name]).max() / np.abs(ref[name]).max()
        ts = []
        for _ in range(8):
            ctx.timer_start(); f(y); ts.append(ctx.timer_stop_ms())
        t = float(np.median(ts))
        print(f"variant {var} {name:8s} {t:7.3f} ms  {nbytes / t / 1e6:7.1f} GB/s  relerr {err:.2e}", flush=True)
if os.environ.get('SPMV_TIMING'):
    for var in (1, 2):
        ctx.set_option('spmv_variant', var); ctx.set_option('spmv_timing', 1); A.spmv(x, y); ctx.set_option('spmv_timing', 0)
# P / R
if nl > 1:
    P = pb.PP[-1]; R = P.transpose()
    xc = ctx.vector(rng.standard_normal(P.shape[1])); yf = ctx.vector(P.shape[0]); yc = ctx.vector(P.shape[1])
    for var in (0, 1):
        ctx.set_option("spmv_variant", var)
        for name, f, M in (("P", lambda: P.spmv(xc, yf), P), ("R", lambda: R.spmv(yf, yc), R)):
            ts = []
            for _ in range(8):
                ctx.timer_start(); f(); ts.append(ctx.timer_stop_ms())
            t = float(np.median(ts))
            nb = M.nnz * 12 + M.shape[0] * 20 + M.shape[1] * 8
            print(f"variant {var} {name} spmv {t:7.3f} ms {nb / t / 1e6:7.1f} GB/s  sum {yf.sum() if name=='P' else yc.sum():.12e}", flush=True)
for l in range(nl - 1):
    Al = pb.KK[l]; xx = ctx.vector(rng.standard_normal(Al.shape[0])); yy = ctx.vector(Al.shape[0])
    for var in (0, 1):
        ctx.set_option("spmv_variant", var)
        ts = []
        for _ in range(8):
            ctx.timer_start(); Al.spmv(xx, yy); ts.append(ctx.timer_stop_ms())
        print(f"variant {var} level {l} spmv {np.median(ts):7.4f} ms  sum {yy.sum():.12e}", flush=True)
del pb
