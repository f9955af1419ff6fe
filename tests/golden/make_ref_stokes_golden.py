"""Generate tests/golden/ref_stokes_*.npz from the REFERENCE ITSELF: tests/cpp/ref_stokes.cpp -- a Taylor-Hood system
U, V, W (SECOND) + P (FIRST) on a HEX27 box, driven through the reference's own MultiLevelMesh / MultiLevelSolution /
LinearImplicitSystem classes with the assembly callback of applications/003_NavierStokes/SteadyStokes/main.cpp
(AssembleMatrixResSteadyStokes, :290-598) compiled in place -- on the single-process host backend of oracle/ref_build
(no PETSc / MPI).  Every integer (system dofs of every variable, KKoffset, sparsity counts and pattern, prolongator
structure, Dirichlet flags per variable from the application's own SetBoundaryCondition) and every value (assembled
matrix, Galerkin operator, prolongator, assembled residual at the initial fields of ref_stokes.cpp) is reference output.
The "ns" case swaps in the library routine femus::AssembleNavierStokes_AD (03_navier_stokes.hpp:21-413) on a
NonLinearImplicitSystem: residual of the Galerkin form, the Jacobian adept recorded, the boundary pressure block.

Run in the build container (needs /root/reference):   python tests/golden/make_ref_stokes_golden.py"""
import os
import re
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle.ref_build import build as rb  # noqa: E402

CASES = {"box211_q2q1_2lev": (2, 1, 1, 2), "ns_box211_q2q1_2lev": (2, 1, 1, 2, "ns")}
# Dirichlet flags only: the same system with MultiLevelSolution::FixSolutionAtOnePoint("P") (which level, which dof)
BDC_CASES = {"ns_fix_box211_q2q1_3lev": (2, 1, 1, 3, "ns", "fix")}


def run_case(name, args, newton=False):
    exe = os.path.join(rb.OUT, "ref_stokes_host")
    work = tempfile.mkdtemp(prefix="refstokes_")
    try:
        for d in ("input", "output", "dump"):
            os.makedirs(os.path.join(work, d))
        env = dict(os.environ, FEMUS_REF_DUMP=os.path.join(work, "dump"), GLIBC_TUNABLES="glibc.malloc.tcache_count=0")
        r = subprocess.run([exe] + [str(a) for a in args], cwd=work, env=env, capture_output=True, text=True, timeout=3600)
        if r.returncode:
            raise RuntimeError(f"{name}: reference run failed ({r.returncode})\n{r.stdout[-2000:]}\n{r.stderr[-2000:]}")
        ire = float(re.search(r"IReynolds\s+([0-9.eE+-]+)", r.stdout).group(1))
        out = {"box": np.array(args[:3]), "nlevels": np.array(args[3]), "IReynolds": np.array(ire)}      # "ns": the routine's nu = 1
        if newton:      # per Newton iteration and variable (U, V, W, P): ||Eps||_2 and ||Sol||_2 as printed (NonLinearImplicitSystem.cpp:137)
            rows = re.findall(r"Nonlinear Eps_l2norm/Sol_l2norm (\w)=\s*[0-9.eE+-]+\s*\*\* Eps_l2norm=\s*([0-9.eE+-]+)\s*\*\* Sol_l2norm=\s*([0-9.eE+-]+)", r.stdout)
            assert len(rows) % 4 == 0 and [v for v, _, _ in rows[:4]] == ["U", "V", "W", "P"]
            out["newton_eps_l2"] = np.array([float(e) for _, e, _ in rows]).reshape(-1, 4)
            out["newton_sol_l2"] = np.array([float(s) for _, _, s in rows]).reshape(-1, 4)
        dt = {"i4": np.int32, "i8": np.int64, "f8": np.float64}
        for f in sorted(os.listdir(os.path.join(work, "dump"))):
            m = re.match(r"L(\d+)_(\w+)\.(i4|i8|f8)$", f)
            out[f"L{int(m.group(1))}_{m.group(2)}"] = np.fromfile(os.path.join(work, "dump", f), dtype=dt[m.group(3)])
        return out
    finally:
        shutil.rmtree(work, ignore_errors=True)


# the reference's own Newton loop (NonLinearImplicitSystem::MGsolve) on ONE level, where every linear solve is the host
# backend's exact LU: the "Nonlinear Eps_l2norm" lines it prints are the Newton updates
NEWTON_CASES = {"ns_newton_box221_q2q1_1lev": (2, 2, 1, 1, "ns", "-", "newton", 5)}

PARTITION_CASES = {"box322_3lev": ["box", 3, 2, 2, 3, 1, 3, 8, 1000000],
                   "cube_mixed_3groups_2lev": ["file", "input/cube_mixed_3groups.neu", 2, 1, 3, 8]}


def run_partition(args):
    """MeshASMPartitioning::DoPartition on every level (tests/cpp/ref_partition.cpp) -> dict"""
    import json
    exe = os.path.join(rb.OUT, "ref_partition_host")
    work = tempfile.mkdtemp(prefix="refpart_")
    try:
        for d in ("input", "output"):
            os.makedirs(os.path.join(work, d))
        for f in os.listdir(HERE):
            if f.endswith(".neu"):
                shutil.copy(os.path.join(HERE, f), os.path.join(work, "input", f))
        env = dict(os.environ, GLIBC_TUNABLES="glibc.malloc.tcache_count=0")
        r = subprocess.run([exe] + [str(a) for a in args], cwd=work, env=env, capture_output=True, text=True, timeout=3600)
        if r.returncode:
            raise RuntimeError(f"reference run failed ({r.returncode})\n{r.stdout[-2000:]}\n{r.stderr[-2000:]}")
        line = [ln for ln in r.stdout.splitlines() if ln.startswith("REF_PARTITION_JSON ")][-1]
        return dict(json.loads(line[len("REF_PARTITION_JSON "):]), args=[str(a) for a in args])
    finally:
        shutil.rmtree(work, ignore_errors=True)


def main():
    import json
    rb.build()
    for name, args in NEWTON_CASES.items():
        out = run_case(name, args, newton=True)
        top = int(args[3]) - 1
        keep = {k: v for k, v in out.items() if k.startswith(("newton_", f"L{top}_SOL_", f"L{top}_Bdc", f"L{top}_KKoffset")) or k in ("box", "nlevels", "IReynolds")}
        np.savez_compressed(os.path.join(HERE, f"ref_stokes_{name}.npz"), **keep)
        print(f"{name}: Newton updates (U)", keep["newton_eps_l2"][:, 0])
    for name, args in BDC_CASES.items():
        out = run_case(name, args)
        keep = {k: v for k, v in out.items() if k.endswith(("_Bdc", "_bdcIndex", "_KKoffset")) or k in ("box", "nlevels")}
        np.savez_compressed(os.path.join(HERE, f"ref_stokes_{name}_bdc.npz"), **keep)
        print(f"{name}: Dirichlet flags of {len(keep)} arrays")
    parts = {name: run_partition(args) for name, args in PARTITION_CASES.items()}
    with open(os.path.join(HERE, "ref_partition.json"), "w") as f:
        json.dump(parts, f, separators=(",", ":"))
    print("ref_partition.json:", {k: [lv["nel"] for lv in v["levels"]] for k, v in parts.items()})
    for name, args in CASES.items():
        out = run_case(name, args)
        path = os.path.join(HERE, f"ref_stokes_{name}.npz")
        np.savez_compressed(path, **out)
        print(f"{name}: {len(out)} arrays, {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
