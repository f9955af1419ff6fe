"""Generate tests/golden/ref_poisson_*.npz from the REFERENCE ITSELF: applications/001_Poisson/main.cpp, unmodified,
compiled with the reference's own mesh / solution / system sources on the single-process host backend of
oracle/ref_build (no PETSc / MPI), run here on small inputs.  Every integer in the fixtures (node numbering, element
dofs, dof offsets, KKoffset, sparsity counts and patterns, prolongator structure, Dirichlet rows) and every value
(coordinates, assembled matrices, prolongators, right-hand sides, the residual norms printed by
LinearImplicitSystem.cpp:426) is reference output, not a restatement.

Run in the build container (needs /root/reference):   python tests/golden/make_ref_golden.py

The JSON "box" branch of main.cpp walks a std::map returned BY VALUE (MultiLevelSolution.cpp:611-621 over
Mesh::GetBoundaryInfo(), Mesh.hpp:377): the iterator dangles, and with glibc's tcache the freed nodes are overwritten
before they are read.  The run therefore disables the tcache (GLIBC_TUNABLES), which leaves the freed nodes intact as
on the allocators the reference was developed with; the sources stay untouched."""
import json
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

FACES = ("bottom", "top", "left", "right", "front", "behind")       # MeshGeneration.cpp:1038-1071 names the six box faces


def box_input(n, nlevels, fe_order, source="1.", dirichlet=FACES, neumann=None, smoother="gmres"):
    bcs = [{"facename": f, "bdc_type": "dirichlet"} for f in dirichlet]
    for f, v in (neumann or {}).items():
        bcs.append({"facename": f, "bdc_type": "neumann", "bdc_func": v})
    return {
        "multilevel_mesh": {"first": {"type": {"box": {"nx": n[0], "ny": n[1], "nz": n[2], "xa": 0.0, "xb": 1.0, "ya": 0.0, "yb": 1.0,
                                                        "za": 0.0, "zb": 1.0, "elem_type": "Hex27"}}}},
        "multilevel_solution": {"multilevel_mesh": {"first": {"variable": {"first": {
            "name": "T", "fe_order": fe_order, "init_func": "0.", "func_source": source, "boundary_conditions": bcs}}}}},
        "multilevel_problem": {"multilevel_mesh": {"first": {"system": {"poisson": {"linear_solver": {
            "max_number_linear_iteration": 6, "abs_conv_tol": 1.e-30,
            "type": {"multigrid": {"nlevels": nlevels, "npresmoothing": 1, "npostsmoothing": 1, "mgtype": "V_cycle",
                                   "smoother": {"type": {smoother: {"ksp": "gmres", "precond": "ilu", "rtol": 1.e-12, "atol": 1.e-20,
                                                                    "divtol": 1.e+50, "max_its": 4}}}}}}}}}}},
    }


def file_input(neu, nlevels, fe_order):
    d = box_input((1, 1, 1), nlevels, fe_order, source="0.")
    d["multilevel_mesh"] = {"first": {"type": {"filename": "input/" + neu}}}
    # the shipped 3-D inputs: Dirichlet on "top", flux 0.2 on "right" (input3D_Hex_second.json); with a mesh file the
    # application takes its boundary conditions from its own SetBoundaryCondition (main.cpp:22-44), not from here
    return d


CASES = {
    # name: (input, keep matrix values up to this level)
    "box222_q2_3lev": (box_input((2, 2, 2), 3, "second"), 1),
    "box222_q1_3lev": (box_input((2, 2, 2), 3, "first"), 2),
    "box324_q2_2lev_neumann": (box_input((3, 2, 4), 2, "second", source="0.", dirichlet=("top",), neumann={"right": "0.2"}), 1),
    "box232_q1_2lev_source_xyz": (box_input((2, 3, 2), 2, "first", source="1.+x*y-z^2"), 1),
    "cube_hex_q2_2lev": (file_input("cube_hex27_2x2x2.neu", 2, "second"), 1),
    "cube_tet_q2_2lev": (file_input("cube_tet10.neu", 2, "second"), 1),
    "cube_tet_serendipity_2lev": (file_input("cube_tet10.neu", 2, "serendipity"), 1),
    "cube_wedge_q2_2lev": (file_input("cube_wedge18.neu", 2, "second"), 1),
    "cube_mixed_q2_2lev": (file_input("cube_mixed.neu", 2, "second"), 1),
    "cube_mixed_q1_2lev": (file_input("cube_mixed.neu", 2, "first"), 1),
}


def run_case(name, spec, keep_values_to):
    exe = os.path.join(rb.OUT, "ref_poisson_host")
    work = tempfile.mkdtemp(prefix="refgold_")
    try:
        os.makedirs(os.path.join(work, "input"))
        os.makedirs(os.path.join(work, "output"))
        os.makedirs(os.path.join(work, "dump"))
        for f in os.listdir(HERE):
            if f.endswith(".neu"):
                shutil.copy(os.path.join(HERE, f), os.path.join(work, "input", f))
        with open(os.path.join(work, "input", "in.json"), "w") as f:
            json.dump(spec, f)
        env = dict(os.environ, FEMUS_REF_DUMP=os.path.join(work, "dump"), GLIBC_TUNABLES="glibc.malloc.tcache_count=0")
        r = subprocess.run([exe, "-i", "input/in.json"], cwd=work, env=env, capture_output=True, text=True, timeout=3600)
        if r.returncode:
            raise RuntimeError(f"{name}: reference run failed ({r.returncode})\n{r.stdout[-2000:]}\n{r.stderr[-2000:]}")
        trace = [float(x) for x in re.findall(r"Linear Res\s+L2norm Sol\s*=\s*([0-9.eE+-]+)", r.stdout)]
        out = {"residual_trace": np.array(trace), "input_json": np.array(json.dumps(spec))}
        dt = {"i4": np.int32, "i8": np.int64, "f8": np.float64}
        for f in sorted(os.listdir(os.path.join(work, "dump"))):
            m = re.match(r"L(\d+)_(\w+)\.(i4|i8|f8)$", f)
            level, key, kind = int(m.group(1)), m.group(2), m.group(3)
            a = np.fromfile(os.path.join(work, "dump", f), dtype=dt[kind])
            if key == "KK_val" and level > keep_values_to:
                # large matrices: the pattern stays complete, of the values only row sums, absolute row sums and the diagonal
                rp = np.fromfile(os.path.join(work, "dump", f"L{level}_KK_rowptr.i8"), dtype=np.int64)
                ci = np.fromfile(os.path.join(work, "dump", f"L{level}_KK_col.i4"), dtype=np.int32)
                rows = np.repeat(np.arange(len(rp) - 1), np.diff(rp))
                out[f"L{level}_KK_rowsum"] = np.bincount(rows, weights=a, minlength=len(rp) - 1)
                out[f"L{level}_KK_absrowsum"] = np.bincount(rows, weights=np.abs(a), minlength=len(rp) - 1)
                diag = np.zeros(len(rp) - 1)
                diag[rows[ci == rows]] = a[ci == rows]
                out[f"L{level}_KK_diag"] = diag
                continue
            out[f"L{level}_{key}"] = a
        return out
    finally:
        shutil.rmtree(work, ignore_errors=True)


def main(only=None):
    rb.build()
    for name, (spec, keep) in CASES.items():
        if only and name not in only:
            continue
        out = run_case(name, spec, keep)
        path = os.path.join(HERE, f"ref_poisson_{name}.npz")
        np.savez_compressed(path, **out)
        print(f"{name}: {len(out)} arrays, trace {out['residual_trace']}, {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main(sys.argv[1:])
