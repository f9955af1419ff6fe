"""Where the end-to-end leg of bench.py loses time against the device-resident step (one GPU): the same loop with the
upload, the download and the per-step scalar read switched off one at a time."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from femus_b200 import capi
from femus_b200.poisson import PoissonMG

ctx = capi.Context(0)
pb = PoissonMG(ctx, 16, 16, 16, 4, "biquadratic")
top = pb.hier.levels[-1]
n = pb.n
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
h_xyz, h_conn = pin(top.xyz), pin(top.conn)
h_sol = torch.zeros(n, dtype=torch.float64).pin_memory()
h_eps = torch.zeros(n, dtype=torch.float64).pin_memory()
state = {"shadow": ctx.vector(n), "eps_shadow": ctx.vector(n)}


def upload_next():
    ctx.open_copies()
    pb.mesh.prefetch(h_xyz.data_ptr(), h_conn.data_ptr())
    state["shadow"].prefetch(h_sol.data_ptr(), n)
    ctx.mark_copies()


def run(K, upload, download, scalar, order="scalar_first"):
    for _ in range(3):
        pb.step()
    ctx.sync()
    if upload:
        upload_next()
        for k in range(2):
            ctx.wait_marked(); pb.mesh.swap(); pb.SOL, state["shadow"] = state["shadow"], pb.SOL
            upload_next() if k < 1 else None
            pb.step()
        ctx.join_copies()
    ctx.sync()
    t_host = 0.0
    ctx.timer_start()
    if upload:
        upload_next()
    for k in range(K):
        if upload:
            ctx.wait_marked()
            pb.mesh.swap()
            pb.SOL, state["shadow"] = state["shadow"], pb.SOL
            if k + 1 < K:
                upload_next()
        pb.EPS, state["eps_shadow"] = state["eps_shadow"], pb.EPS
        t0 = time.perf_counter()
        pb.step()
        t_host += time.perf_counter() - t0
        if scalar and order == "scalar_first":
            pb.residual_norm()
        if download:
            pb.EPS.fetch(h_eps.data_ptr(), n)
        if scalar and order != "scalar_first":
            pb.residual_norm()
    ctx.join_copies()
    ms = ctx.timer_stop_ms()
    return ms / K, t_host / K * 1e3


# the host -> device link itself: the step's inputs (coordinates + connectivity + solution) from pinned memory, alone
ctx.sync()
t0 = time.perf_counter()
for _ in range(3):
    upload_next()
    ctx.join_copies()
    ctx.sync()
dt = (time.perf_counter() - t0) / 3
nbytes = h_xyz.numel() * 8 + h_conn.numel() * 4 + n * 8
print(json.dumps({"h2d_bytes": nbytes, "h2d_ms_alone": dt * 1e3, "h2d_GBs": nbytes / dt / 1e9}))
for up, down, sc, order in ((0, 0, 0, ""), (0, 0, 1, ""), (0, 1, 1, "download_first"), (0, 1, 1, "scalar_first"), (1, 0, 1, ""),
                            (1, 1, 1, "download_first"), (1, 1, 1, "scalar_first"), (1, 1, 0, "")):
    ms, host = run(6, up, down, sc, order)
    print(json.dumps({"upload": up, "download": down, "scalar_per_step": sc, "order": order, "ms_per_step": ms, "host_enqueue_ms_per_step": host}))
