#!/usr/bin/env python
"""Benchmark of the FEMuS hot path on B200: per-element Poisson assembly + geometric-multigrid
V-cycle (BASELINE.json: "DOF/sec assembled + V-cycle SpMV GB/s (fp64)").

    python bench.py --gpus N --steps K --warmup W [--impl reference]

A step = one pass of the hot path over the mesh: SetResZero, element assembly of the finest-level
matrix and residual, Galerkin chain P^T A P down the hierarchy, MGSetLevel on every level
(Dirichlet penalty, Jacobi diagonal), one MGSolve (= one multiplicative V-cycle, outer PREONLY).
`value` = finest-level DOFs / step time with all inputs resident in HBM; `e2e` = the same with the
mesh (coordinates, connectivity) and the current solution copied from pinned host memory and the
correction EPS + residual norm copied back inside the timed region.

Workload at N=1: BASELINE configs[1] (3-D Poisson, Hex27, 128^3 elements, 4-level MG).  N>1: the
box grows with N (weak scaling, z-slab partition) up to configs[2] (256^3 on 8 GPUs).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "dof_per_sec_assembly_plus_vcycle"
UNIT = "DOF/s"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get("hbm_gbs", None), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device=0):
        self.device = device
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.device)], stdout=open(self.path, "w"),
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            f = [t.strip() for t in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if sm:
            out["sm_mhz"] = float(np.median(sm))
            out["sm_max_mhz"] = float(max(mx))
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


# ---------------------------------------------------------------------------------------------
class HostMemoryNearGpu:
    """Places the pinned host buffers of the end-to-end leg on the NUMA node of this rank's GPU: on a two-socket box
    eight ranks pulling 6 GB per step through one socket's memory controllers (first touch puts every rank's pages
    where its process happens to run) is what limits the end-to-end leg at N = 8.  set_mempolicy(MPOL_PREFERRED) around
    the allocations, restored afterwards; reports what it found / did (containers may forbid the call)."""
    SYS_set_mempolicy = 238          # x86_64
    MPOL_DEFAULT, MPOL_PREFERRED = 0, 1

    def __init__(self, torch, local_rank):
        self.info = {"gpu_numa_node": None, "nodes": None, "policy": "default"}
        self.node = None
        try:
            nodes = [int(d[4:]) for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()]
            self.info["nodes"] = len(nodes)
            p = torch.cuda.get_device_properties(local_rank)
            bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
            node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
            self.info["gpu_numa_node"] = node
            if node >= 0 and len(nodes) > 1 and node in nodes:
                self.node = node
        except Exception as e:          # no sysfs / no pci ids: leave the default policy
            self.info["policy"] = f"default ({type(e).__name__})"

    def _set(self, mode, node):
        import ctypes
        libc = ctypes.CDLL(None, use_errno=True)
        if mode == self.MPOL_DEFAULT:
            rc = libc.syscall(self.SYS_set_mempolicy, mode, None, 0)
        else:
            mask = (ctypes.c_ulong * 16)()
            mask[node // 64] = 1 << (node % 64)
            rc = libc.syscall(self.SYS_set_mempolicy, mode, mask, 16 * 64 + 1)
        if rc != 0:
            raise OSError(ctypes.get_errno(), os.strerror(ctypes.get_errno()))

    def __enter__(self):
        if self.node is not None:
            try:
                self._set(self.MPOL_PREFERRED, self.node)
                self.info["policy"] = "preferred"
            except Exception as e:
                self.info["policy"] = f"default (set_mempolicy: {e})"
                self.node = None
        return self

    def __exit__(self, *a):
        if self.node is not None:
            try:
                self._set(self.MPOL_DEFAULT, 0)
            except Exception:
                pass
        return False


# ---------------------------------------------------------------------------------------------
def cpu_reference_problem(n0, nlevels, order, nthreads):
    """Setup (untimed) of the CPU sample, what system.init() does once: oracle meshes, patterns,
    prolongators with their Dirichlet rows/columns zeroed, Dirichlet index lists."""
    from oracle import mesh_box as mb, mg, ref, cpu_port, fe_hex
    lv = mb.build_hierarchy(n0, n0, n0, nlevels)
    top = lv[-1]
    rp, ci = mb.sparsity(top, order)
    dof = mb.system_dof(top, order)
    kind = "reference" if ref.available() else "port"
    R = ref.RefHex(order) if kind == "reference" else None
    state = dict(lv=lv, top=top, rp=rp, ci=ci, dof=dof, kind=kind, R=R, order=order, nthreads=nthreads, H=None)
    return state


def cpu_reference_step(st):
    """One pass of the hot path on the host cores: assembly by the reference's own FE kernel
    (oracle/_ref, OpenMP over elements) or the numpy port, Galerkin chain (OpenMP C port), level
    setup (penalty, Jacobi diagonal), one V-cycle with the OpenMP C port of the CSR kernels."""
    import scipy.sparse as sp
    from oracle import mesh_box as mb, mg, cpu_port
    top, order, nt = st["top"], st["order"], st["nthreads"]
    n = mb.ndofs(top, order)
    ptap = lambda P, Af, prp, pci: cpu_port.ptap(P, Af, prp, pci, nt)
    t0 = time.perf_counter()
    if st["kind"] == "reference":
        vals, rhs, _ = st["R"].assemble_csr(top.conn, st["dof"], top.xyz, np.zeros(n), st["rp"], st["ci"], 1.0, nt)
        A = sp.csr_matrix((vals, st["ci"], st["rp"]), shape=(n, n))
    else:
        A, rhs = mb.assemble(top, order)
    t1 = time.perf_counter()
    if st["H"] is None:       # first call: also builds the static parts (prolongators, patterns) -- warm-up only
        st["H"] = mg.Hierarchy(st["lv"], order, A_top=A, rhs=rhs, coarse_lu=False, ptap=ptap)
        t1 = time.perf_counter()
    H = st["H"]
    H.set_operator(A, rhs, ptap)                                             # Galerkin chain + penalty + diagonal
    t2 = time.perf_counter()
    M = cpu_port.PortMG(H, nt)
    t3 = time.perf_counter()
    res, eps = M.mg_solve(rhs.copy(), np.zeros(n))
    t4 = time.perf_counter()
    return dict(total=(t1 - t0) + (t2 - t1) + (t4 - t3), assembly=t1 - t0, galerkin_setup=t2 - t1, vcycle=t4 - t3,
                ndofs=n, nel=top.nel, resnorm=float(np.linalg.norm(res)))


def run_cpu_sample(n0, nlevels, order, steps, warmup):
    nthreads = os.cpu_count() or 1
    st = cpu_reference_problem(n0, nlevels, order, nthreads)
    for _ in range(max(warmup, 1)):      # the first pass also builds the static hierarchy (system.init())
        cpu_reference_step(st)
    rs = [cpu_reference_step(st) for _ in range(steps)]
    tot = float(np.mean([r["total"] for r in rs]))
    nve = 27 if order == "biquadratic" else 8
    out = {
        "value": rs[0]["ndofs"] / tot, "unit": UNIT, "cores": nthreads, "kind": st["kind"],
        "sample": (f"{n0 * 2 ** (nlevels - 1)}^3 {'Hex27' if nve == 27 else 'Hex8'} elements, {nlevels}-level MG, "
                   f"{rs[0]['ndofs']} dofs: assembly by "
                   f"{'the compiled reference FE kernel (oracle/_ref, OpenMP)' if st['kind'] == 'reference' else 'the numpy port'}"
                   f", Galerkin by the OpenMP C port (two Gustavson sweeps), V-cycle by the OpenMP C port; {steps} step(s)"),
        "ms_per_step": tot * 1e3,
        "assembly_elem_dof_per_s": rs[0]["nel"] * nve / float(np.mean([r["assembly"] for r in rs])),
        "assembly_ms": float(np.mean([r["assembly"] for r in rs])) * 1e3,
        "galerkin_setup_ms": float(np.mean([r["galerkin_setup"] for r in rs])) * 1e3,
        "vcycle_ms": float(np.mean([r["vcycle"] for r in rs])) * 1e3,
        "resnorm": rs[-1]["resnorm"],
    }
    return out


# ---------------------------------------------------------------------------------------------
def workload_for(ngpus, n0, nlevels):
    """Weak scaling: per-GPU work fixed at n0^3 coarse elements; boxes 1:(n,n,n) 2:(n,n,2n)
    4:(n,2n,2n) 8:(2n,2n,2n) -> 128^3 at N=1 and 256^3 at N=8 for n0=16, 4 levels."""
    mul = {1: (1, 1, 1), 2: (1, 1, 2), 4: (1, 2, 2), 8: (2, 2, 2)}[ngpus]
    return n0 * mul[0], n0 * mul[1], n0 * mul[2]


def bounds_for(nx, ny, nz):
    """The domain grows with the box so that the elements stay cubes at every N (unit cube at N=1 and N=8)."""
    m = float(min(nx, ny, nz))
    return (0.0, nx / m, 0.0, ny / m, 0.0, nz / m)


def ncu_capture(kernel, workload_key):
    """Per-launch DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum) and pipe figures of `kernel` from the
    committed ncu --set full capture of this workload (profiles/ncu_kernels.json, written by tools/ncu_raw_extract.py
    from the .ncu-rep; nothing is typed in by hand).  None when no capture of this kernel / workload is on file."""
    try:
        db = json.load(open(os.path.join(ROOT, "profiles", "ncu_kernels.json")))
    except Exception:
        return None
    for rec in db.get("kernels", []):
        if rec.get("workload") == workload_key and rec.get("kernel") == kernel:
            return rec
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n0", type=int, default=16, help="coarsest-level elements per side per GPU")
    ap.add_argument("--levels", type=int, default=4)
    ap.add_argument("--order", default="biquadratic", choices=["linear", "biquadratic"])
    ap.add_argument("--cpu-n0", type=int, default=8, help="coarsest-level size of the CPU sample (8: 64^3 elements with 4 levels)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--halo", default="peer", choices=["peer", "nccl"], help="interface sums: peer-memory exchange or packed ncclAllReduce")
    args = ap.parse_args()
    W = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    nve = 27 if args.order == "biquadratic" else 8
    nx, ny, nz = workload_for(args.gpus, args.n0, args.levels)
    f = 2 ** (args.levels - 1)
    workload = (f"3D Poisson, {'Hex27' if nve == 27 else 'Hex8 (Q1 on Hex27 geometry)'}, {nx * f}x{ny * f}x{nz * f} elements, "
                f"{args.levels}-level geometric MG V-cycle (Richardson 0.5 + Jacobi, 1 pre / 1 post, coarse Jacobi-PCG), fp64")
    config = {"workload": workload, "coarse_box": [nx, ny, nz], "levels": args.levels, "fe_order": args.order,
              "domain": list(bounds_for(nx, ny, nz)),
              "partition": "single GPU" if args.gpus == 1 else f"z-slabs over {args.gpus} GPUs",
              "interface_sums": None if args.gpus == 1 else ("peer-memory exchange over NVLink (remote stores + flags)" if args.halo == "peer"
                                                               else "packed ncclAllReduce"),
              "l2": "inputs larger than L2 (finest CSR >> 126 MB); no flush needed"}

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        r = run_cpu_sample(args.cpu_n0, args.levels, args.order, max(args.steps, 1), args.warmup)
        fs = args.cpu_n0 * f
        config = dict(config, reference_sample=f"every step of this arm is a BOUNDED SAMPLE of the workload: {fs}x{fs}x{fs} elements "
                      f"({r['sample'].split(':')[0]}), not {nx * f}x{ny * f}x{nz * f}; DOF/s is per finest-level dof of the sample")
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}, "residual_after_one_vcycle": r["resnorm"],
                "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "assembly_elem_dof_per_s": r["assembly_elem_dof_per_s"], "assembly_ms": r["assembly_ms"],
                "galerkin_setup_ms": r["galerkin_setup_ms"], "vcycle_ms": r["vcycle_ms"], "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ B200 arm
    import torch
    import torch.distributed as dist
    from femus_b200 import capi
    from femus_b200.poisson import PoissonMG

    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} needs {args.gpus} ranks (torchrun --nproc-per-node {args.gpus}), got WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    ctx = capi.Context(local_rank)
    dist_arg = None
    if world > 1:
        # one rank per GPU: torch.distributed carries the rendezvous, the library owns its NCCL communicator
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        uid = [capi.Context.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(world, rank, uid[0])
        from femus_b200.dist import torch_allgather
        dist_arg = (rank, world, torch_allgather())

    def barrier():
        ctx.sync()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    t_setup = time.time()
    pb = PoissonMG(ctx, nx, ny, nz, args.levels, args.order, bounds=bounds_for(nx, ny, nz), dist=dist_arg, peer=args.halo == "peer")
    ctx.sync()
    t_setup = time.time() - t_setup
    top = pb.hier.levels[-1]
    n_loc = pb.n
    n, nel = pb.n_global, int(sum_over_ranks(pb.nel))

    # pinned host copies of the per-step inputs / outputs for the end-to-end leg
    def pinned(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t
    near = HostMemoryNearGpu(torch, local_rank)
    with near:
        h_xyz, h_conn = pinned(top.xyz), pinned(top.conn)
        h_sol = torch.zeros(n_loc, dtype=torch.float64).pin_memory()
        h_eps = torch.zeros(n_loc, dtype=torch.float64).pin_memory()
    config["host_buffers"] = dict(near.info, cpus=len(os.sched_getaffinity(0)))
    h2d = sum_over_ranks(h_xyz.numel() * 8 + h_conn.numel() * 4 + n_loc * 8)
    d2h = sum_over_ranks(n_loc * 8 + 8)

    def step_resident():
        pb.step()

    # End-to-end leg: every step's inputs (mesh coordinates, connectivity, current solution) come from
    # pinned host memory and its outputs (correction EPS, residual norm) go back to the host, all inside
    # the timed region.  The inputs are double buffered on the device: step k+1's upload runs on the
    # copy stream while step k computes (b2_ctx_open_copies / b2_mesh_prefetch / b2_vec_prefetch).
    sol_shadow, eps_shadow = ctx.vector(n_loc), ctx.vector(n_loc)
    if pb.halo[-1] is not None:
        sol_shadow.set_halo(pb.halo[-1])
        eps_shadow.set_halo(pb.halo[-1])
    e2e_state = {"shadow": sol_shadow, "eps_shadow": eps_shadow}

    def upload_next():
        ctx.open_copies()
        pb.mesh.prefetch(h_xyz.data_ptr(), h_conn.data_ptr())
        e2e_state["shadow"].prefetch(h_sol.data_ptr(), n_loc)
        ctx.mark_copies()

    def step_e2e(more):
        ctx.wait_marked()                  # the compute stream waits for this step's inputs (not for the last download)
        pb.mesh.swap()
        pb.SOL, e2e_state["shadow"] = e2e_state["shadow"], pb.SOL
        if more:
            upload_next()                  # next step's inputs, behind this step's compute
        pb.EPS, e2e_state["eps_shadow"] = e2e_state["eps_shadow"], pb.EPS     # the other buffer may still be downloading
        pb.step()
        r = pb.residual_norm()             # D2H of the scalar, synchronises the compute stream
        # the 136 MB download goes on the copy stream AFTER the scalar was read: both are device->host transfers, and an
        # 8-byte read queued behind it on the same copy engine would wait for all of it (measured: +1.9 ms per step)
        pb.EPS.fetch(h_eps.data_ptr(), n_loc)      # overlaps the next step's assembly
        return r

    # ---- warm-up
    for _ in range(W):
        step_resident()
    barrier()
    Afine = pb.KK[-1]

    # ---- timed region: K steps, device time on the library stream
    sampler = ClockSampler(local_rank)
    sampler.start()
    ctx.profile_only(Afine)
    ctx.launches(reset=True)
    barrier()
    ctx.timer_start()
    for _ in range(args.steps):
        step_resident()
    ms_total_local = ctx.timer_stop_ms()
    barrier()
    ms_total = max_over_ranks(ms_total_local)          # device time, max over ranks
    launches = int(sum_over_ranks(ctx.launches(reset=True)))
    nsp, ms_sp = ctx.profile_read(Afine)
    ctx.profile_only(None)
    # phase breakdown (separate passes, same data); the assembly kernel is event-timed on its own
    phases = {}
    kernel_ms = {}
    for name, fn in (("assembly", pb.assemble), ("galerkin", pb.galerkin), ("mg_set_levels", pb.mg_set_levels),
                     ("vcycle", pb.mg_solve)):
        ts = []
        for _ in range(3):
            ctx.sync()
            if name == "assembly":
                ctx.profile_only(pb.asm)
            ctx.timer_start()
            fn()
            ts.append(ctx.timer_stop_ms())
            if name == "assembly":
                nk, msk = ctx.profile_read(pb.asm)
                ctx.profile_only(None)
                kernel_ms.setdefault("assembly", []).append(msk / max(nk, 1))
        phases[name] = float(np.median(ts))
    asm_kernel_ms = float(np.median(kernel_ms["assembly"]))
    # V-cycle by level and phase (events around every phase; one extra cycle)
    pb.mg.set_timing(True)
    pb.mg_solve()
    tv = pb.mg.get_timing(args.levels)
    pb.mg.set_timing(False)
    vcycle_phases = {f"level{l}": {k: round(float(tv[l][j]), 4) for j, k in enumerate(("pre_smooth", "residual", "restrict", "coarse_solve",
                                                                                  "prolong", "post_smooth")) if tv[l][j] > 0}
                     for l in range(args.levels)}
    if world > 1 and args.halo == "peer":      # the same V-cycle with the interface sums as a packed ncclAllReduce, for comparison
        ctx.set_option("halo_peer", 0)
        ts = []
        for _ in range(3):
            barrier()
            ctx.timer_start()
            pb.mg_solve()
            ts.append(max_over_ranks(ctx.timer_stop_ms()))
        phases["vcycle_with_nccl_allreduce"] = float(np.median(ts))
        ctx.set_option("halo_peer", 1)
        ts = []
        for _ in range(3):
            barrier()
            ctx.timer_start()
            pb.mg_solve()
            ts.append(max_over_ranks(ctx.timer_stop_ms()))
        phases["vcycle_max_over_ranks"] = float(np.median(ts))
    # ---- end-to-end: host buffers in, host buffers out
    upload_next()
    for k in range(2):
        step_e2e(k < 1)
    ctx.join_copies()
    barrier()
    ctx.timer_start()
    upload_next()
    for k in range(args.steps):
        resnorm = step_e2e(k + 1 < args.steps)
    ctx.join_copies()                      # the last correction has reached the host inside the timed region
    ms_e2e = ctx.timer_stop_ms()
    barrier()
    ms_e2e = max_over_ranks(ms_e2e)
    clocks = sampler.stop()
    # residual trace of a short solve (sanity: the timed path really converges)
    pb.EPS.zero()
    pb.assemble(); pb.galerkin(); pb.mg_set_levels()
    trace = []
    for _ in range(4):
        pb.mg_solve()
        trace.append(pb.residual_norm())

    ms_step = ms_total / args.steps
    value = n / (ms_step * 1e-3)
    e2e_value = n / (ms_e2e / args.steps * 1e-3)
    # ---- rooflines.  `roofline` = the kernel with the largest share of the step (the fused assembly); the finest-level
    # SpMV family of the V-cycle is reported next to it as `roofline_spmv`.
    peak, peak_src = measured_peaks()
    wl_key = f"{args.order}:{nx * f}x{ny * f}x{nz * f}:{args.levels}lev:n{args.gpus}"
    # (1) SpMV family on the finest CSR (3 launches per step: 2 x r = b - A x, 1 x Jacobi sweep); algorithmic bytes per
    # launch as SURVEY section 8(d) defines them
    b_y = pb.spmv_bytes(-1)
    bytes_per_launch = (2 * (b_y + 8 * n_loc) + (b_y + 24 * n_loc)) / 3.0 if world == 1 else b_y + 16 * n_loc
    ach = bytes_per_launch / (ms_sp / max(nsp, 1) * 1e-3) / 1e9 if nsp else None
    cap = ncu_capture("spmv_tma_kernel", wl_key)
    roofline_spmv = {"bound": "hbm", "kernel": "spmv_tma_kernel on the finest-level CSR (2 x r = b - A x, 1 x Jacobi sweep per step)",
                     "achieved": ach, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                     "frac": (ach / peak) if ach else None,
                     "traffic": cap["dram_bytes_per_launch"] if cap else None,
                     "traffic_source": cap["source"] if cap else None,
                     "algorithmic_bytes_per_launch": bytes_per_launch, "launches_timed": nsp,
                     "avg_launch_ms": ms_sp / max(nsp, 1), "share_of_step": ms_sp / ms_total_local}
    if world > 1:
        roofline_spmv["kernel"] = "spmv_tma_kernel on rank 0's finest-level partial CSR (weighted resid x3 per step)"
    # (2) the assembly kernel (fused with the finest Galerkin product).  SURVEY section 8(d): algorithmic work per Q2
    # element 4.5e5 flop (the element loop as the reference writes it: 64 Gauss points x (Jacobian, gradients, 27 x 27
    # x 8 stiffness, residual)) and 7020 B (972 read + 6048 written); reported as flop/s against the chip's fp64 peak,
    # MEASURED in this run by two issue-rate probes (chains of independent DFMA / of mma.sync.m8n8k4.f64 on every SM,
    # a few ms each: b2_ctx_measure_fp64_fma / _tensor; MEASURED_PEAKS.json carries no fp64 figure), with the HBM view
    # beside it.  The sum-factorised kernel performs the same contraction with ~1/5 of the multiply-adds, so the
    # algorithmic figure can exceed the pipe peak; `executed` is what the kernel really issues (ncu instruction counts).
    nel_loc = pb.nel
    kname = ctx.L.b2_asm_kernel_name(pb.asm.h).decode()
    dmma_peak = ctx.measure_fp64_tensor()
    dfma_peak = ctx.measure_fp64_fma()
    on_tensor = "mma" in kname
    fp64_peak = dmma_peak if on_tensor else dfma_peak
    flop_el = 4.5e5 if nve == 27 else 4.5e5 * (8 * 8) / (27 * 27)
    bytes_el = (27 * 3 * 8 + 27 * 4 + nve * 8) + (nve * nve + nve) * 8
    cap = ncu_capture(kname, wl_key)
    ach_tf = nel_loc * flop_el / (asm_kernel_ms * 1e-3) / 1e12
    roofline = {"bound": "tensor" if on_tensor else "fp64",
                "bound_note": None if on_tensor else "FP64 FMA pipe of the CUDA cores (same peak as the FP64 tensor pipe on B200); neither "
                                                     "tensor-core nor HBM bound: see hbm_view",
                "kernel": kname + " (element assembly fused with the finest Galerkin product; one launch per step)",
                "achieved": ach_tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach_tf / fp64_peak,
                "peak_source": "measured in this run: issue-rate probe of " + ("mma.sync.m8n8k4.f64" if on_tensor else "DFMA") +
                               f" chains on every SM (DFMA {dfma_peak:.1f}, DMMA {dmma_peak:.1f} TFLOP/s)",
                "algorithmic_flop_per_launch": nel_loc * flop_el, "algorithmic_flop_per_element": flop_el,
                "avg_launch_ms": asm_kernel_ms, "share_of_step": asm_kernel_ms / ms_step,
                "traffic": cap["dram_bytes_per_launch"] if cap else None,
                "traffic_source": cap["source"] if cap else None,
                "hbm_view": {"algorithmic_bytes_per_launch": nel_loc * bytes_el, "algorithmic_bytes_per_element": bytes_el,
                             "achieved_gbs": nel_loc * bytes_el / (asm_kernel_ms * 1e-3) / 1e9, "peak_gbs": peak,
                             "frac": nel_loc * bytes_el / (asm_kernel_ms * 1e-3) / 1e9 / peak}}
    if cap:
        for k in ("fp64_pipe_pct", "l1tex_data_pipe_pct", "issue_active_pct", "registers_per_thread", "dfma_flop_per_launch"):
            if k in cap:
                roofline.setdefault("ncu", {})[k] = cap[k]
        if cap.get("dfma_flop_per_launch"):
            roofline["executed_flop_per_launch"] = cap["dfma_flop_per_launch"]
            roofline["frac_executed"] = cap["dfma_flop_per_launch"] / (asm_kernel_ms * 1e-3) / 1e12 / fp64_peak

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": W,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "roofline_spmv": roofline_spmv,
            "assembly_elem_dof_per_s": nel * nve / (phases["assembly"] * 1e-3),
            "spmv_gbs": ach, "phases_ms": phases, "vcycle_phases_ms": vcycle_phases, "finest_dofs": n, "finest_nnz": Afine.nnz, "elements": nel,
            "setup_s": t_setup, "residual_trace": trace, "coarse_pcg_iterations": pb.mg.coarse_iterations(),
            "device_bytes": ctx.bytes_in_use(), "dofs_per_rank": n_loc}
    if world > 1:
        barrier()
        dist.destroy_process_group()
        if rank != 0:
            return 0
    if not args.no_cpu_baseline and world == 1:
        try:
            cb = run_cpu_sample(args.cpu_n0, args.levels, args.order, 1, 1)
            line["cpu_baseline"] = {k: v for k, v in cb.items() if k != "resnorm"}
            # in-run parity: the SAME sample through the GPU path (assembly, Galerkin chain, level setup, one V-cycle from
            # a zero guess) against the CPU run's residual norm after that cycle
            ps = PoissonMG(ctx, args.cpu_n0, args.cpu_n0, args.cpu_n0, args.levels, args.order)
            ps.EPS.zero()
            ps.assemble()
            r0 = ps.RES.norm(2)              # the residual the cycle starts from: the scale of the comparison (as in the parity tests)
            ps.galerkin()
            ps.mg_set_levels()
            ps.mg_solve()
            g = ps.RES.norm(2)
            line["parity"] = {"what": "||RES||_2 after assembly + one V-cycle on the CPU baseline's sample, GPU path vs CPU reference path; "
                                      "rel_err = |gpu - cpu| / ||RES||_2 before the cycle",
                              "gpu": g, "cpu": cb["resnorm"], "residual_before_the_cycle": r0, "rel_err": abs(g - cb["resnorm"]) / r0, "tol": 1e-10}
            del ps
        except Exception as e:      # the baseline is reported, never required for the GPU number
            line["cpu_baseline"] = {"error": repr(e)}
    print(json.dumps(line))
    return 0


if __name__ == "__main__":
    sys.exit(main())
