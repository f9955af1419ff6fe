"""One single-GPU run with MORE THAN 2^31 non-zeros in the finest matrix (north_star's 256^3-at-one-GPU target implies
64-bit row starts / entry offsets everywhere): Hex27, n0^3 coarse elements, 4 levels -- n0 = 24: 192^3 elements,
385^3 = 57 M dofs, 1537^3 = 3.63e9 non-zeros -- through assembly (fused with the finest Galerkin product), the Galerkin
chain, level setup and V-cycles, checked by the size-independent properties of tests/test_gpu_parity.py::
test_full_size_properties: A 1 = 0, symmetry, fused == element-gather Galerkin product, monotone contraction, and the
analytic centre value of -Laplace(u) = 1.

    python tools/big_run.py [n0=24] > gpurun_out/big_run.json"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from femus_b200 import capi
from femus_b200.poisson import PoissonMG


def run(ctx, n0=24, cycles=45):
    t0 = time.time()
    pb = PoissonMG(ctx, n0, n0, n0, 4, "biquadratic")
    ctx.sync()
    out = {"elements": int(pb.nel), "dofs": int(pb.n), "nnz": int(pb.KK[-1].nnz), "setup_s": time.time() - t0,
           "device_bytes": int(ctx.bytes_in_use())}
    n = pb.n
    assert out["nnz"] == (8 * 8 * n0 + 1) ** 3 and n == (2 * 8 * n0 + 1) ** 3
    for _ in range(2):
        pb.step()
    ctx.sync()
    ctx.timer_start()
    for _ in range(3):
        pb.step()
    out["ms_per_step"] = ctx.timer_stop_ms() / 3
    out["dof_per_s"] = n / (out["ms_per_step"] * 1e-3)
    pb.EPS.zero()
    pb.assemble()
    A = pb.KK[-1]
    rng = np.random.default_rng(5)
    h = 1.0 / (8 * n0)
    dmax = 0.0355555555555555 * 2.0 * (h * 128)       # largest entry scales with h
    one, y, x, z = ctx.vector(np.ones(n)), ctx.vector(n), ctx.vector(rng.standard_normal(n)), ctx.vector(rng.standard_normal(n))
    A.spmv(one, y)
    out["A1_max"] = y.norm(0)
    assert out["A1_max"] <= 1e-12 * 125 * dmax
    A.spmv(z, y)
    xAz = x.dot(y)
    A.spmv(x, y)
    zAx = z.dot(y)
    out["symmetry_defect"] = abs(xAz - zAx)
    assert out["symmetry_defect"] <= 1e-12 * x.norm(2) * z.norm(2) * 125 * dmax
    C2 = pb.KK[-2]
    xc, yc1, yc2 = ctx.vector(rng.standard_normal(C2.shape[0])), ctx.vector(C2.shape[0]), ctx.vector(C2.shape[0])
    C2.spmv(xc, yc1)
    pb.gal[-1].apply()
    C2.spmv(xc, yc2)
    yc2.axpy(-1.0, yc1)
    out["fused_vs_gather_galerkin"] = yc2.norm(0) / yc1.norm(0)
    assert out["fused_vs_gather_galerkin"] <= 1e-12
    pb.galerkin(); pb.mg_set_levels()
    trace = [pb.residual_norm()]
    for _ in range(cycles):
        pb.mg_solve()
        trace.append(pb.residual_norm())
    out["residual_trace"] = trace
    assert all(b < a for a, b in zip(trace, trace[1:])), trace
    top = pb.hier.levels[-1]
    centre = int(np.nonzero((np.abs(top.xyz - 0.5) < 1e-12).all(axis=0))[0][0])
    k = np.arange(1, 200, 2)
    sgn = np.where(((k - 1) // 2) % 2 == 0, 1.0, -1.0)
    I, J, K = np.meshgrid(k, k, k, indexing="ij")
    S = sgn[:, None, None] * sgn[None, :, None] * sgn[None, None, :]
    exact = float((64.0 / np.pi ** 5 * S / (I * J * K * (I * I + J * J + K * K))).sum())
    got = float(pb.EPS.get_indexed(np.array([centre], dtype=np.int32))[0])
    out["centre_value"], out["centre_exact"] = got, exact
    assert trace[-1] < 1e-8 * trace[0] and abs(got - exact) <= 2e-6 * exact, (got, exact, trace[-1] / trace[0])
    out["coarse_pcg_iterations"] = pb.mg.coarse_iterations()
    return out


if __name__ == "__main__":
    ctx = capi.Context(0)
    print(json.dumps(run(ctx, int(sys.argv[1]) if len(sys.argv) > 1 else 24)))
