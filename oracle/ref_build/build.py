"""TEST INFRASTRUCTURE: compiles the reference's OWN sources where they lie (/root/reference, read-only) with g++ into
oracle/_ref/ -- no copy of any reference source enters this repository, the reference's build system is not run:

    libfemus_ref_host.a     src/**.cpp of the mesh / solution / system / equation layers (PETSc-, SLEPc-, FSI-, optimal
                            control-, uncertainty-quantification-specific files left out), jsoncpp, b64, adept
                            + the single-process shims of oracle/ref_shims (mpi.h, FemusConfig.hpp, hdf5 / metis / petsc /
                            boost / fparser stand-ins) + the host algebra backend oracle/ref_build/HostBackend.hpp
    ref_dump                oracle/ref_build/ref_dump.cpp: mesh -> numbering -> sparsity -> prolongators -> Dirichlet flags
                            -> assembled system, through the reference's classes, written as .npz-ready text
    ref_poisson_host        applications/001_Poisson/main.cpp, UNMODIFIED, on the host backend
    ref_amr_poisson_host    tests/cpp/ref_amr_poisson.cpp: selectively refined meshes through the reference's AMR path
    ref_stokes_host         tests/cpp/ref_stokes.cpp: a Taylor-Hood system U, V, W, P with the assembly callback of
                            applications/003_NavierStokes/SteadyStokes/main.cpp compiled in place
    ref_partition_host      tests/cpp/ref_partition.cpp: MeshASMPartitioning::DoPartition on every level, printed

    python -m oracle.ref_build.build [--force]
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE = os.path.dirname(HERE)
OUT = os.path.join(ORACLE, "_ref")
OBJ = os.path.join(OUT, "obj_host")
SHIMS = os.path.join(ORACLE, "ref_shims")
LIB = os.path.join(OUT, "libfemus_ref_host.a")
EXCLUDE = ("gencase_deprecated", "Petsc", "petsc", "09_optimal_control", "Slepc", "slepc", "template",
           "Preconditioner.cpp", "TransientFSI", "ism/Line.cpp")
FACTORIES = ("NumericVector.cpp", "SparseMatrix.cpp", "LinearEquationSolver.cpp")


def reference_sources(ref):
    out = []
    for root, _, files in os.walk(os.path.join(ref, "src")):
        for f in sorted(files):
            p = os.path.join(root, f)
            if f.endswith(".cpp") and not any(x in p for x in EXCLUDE):
                out.append(p)
    out += [os.path.join(ref, "external/jsoncpp/jsoncpp-src-0.5.0/src/lib_json", f) for f in ("json_reader.cpp", "json_value.cpp", "json_writer.cpp")]
    out.append(os.path.join(ref, "external/adept/adept-1.1/adept/adept.cpp"))
    out.append(os.path.join(ref, "external/b64/b64-1.4.2/src/b64.c"))
    return out


def flags(ref, backend_dir=HERE):
    inc = [SHIMS, backend_dir]
    for root, dirs, _ in os.walk(os.path.join(ref, "src")):
        inc.append(root)
    inc += [os.path.join(ref, "external/adept/adept-1.1/include"), os.path.join(ref, "external/jsoncpp/jsoncpp-src-0.5.0/include"),
            os.path.join(ref, "external/jsoncpp/jsoncpp-src-0.5.0/src/lib_json"), os.path.join(ref, "external/b64/b64-1.4.2/include")]
    return ["-std=c++17", "-O2", "-w", "-fPIC"] + ["-I" + d for d in inc]


def available(ref="/root/reference"):
    return os.path.isdir(os.path.join(ref, "src"))


def compile_objects(ref, objdir, backend_header, extra=(), force=False):
    """Every reference source -> objdir/*.o; the three factory files get the backend header pre-included."""
    os.makedirs(objdir, exist_ok=True)
    srcs = reference_sources(ref)
    fl = flags(ref) + list(extra)
    jobs = []
    objs = []
    newest_dep = max(os.path.getmtime(os.path.join(d, f)) for d in (SHIMS, HERE, os.path.dirname(backend_header)) for f in os.listdir(d)
                     if f.endswith((".h", ".hpp", ".hh")))
    for s in srcs:
        o = os.path.join(objdir, os.path.relpath(s, ref).replace(os.sep, "_") + ".o")
        objs.append(o)
        if force or not os.path.exists(o) or os.path.getmtime(o) < max(os.path.getmtime(s), newest_dep):
            jobs.append((s, o))

    def run(job):
        s, o = job
        pre = ["-include", backend_header] if os.path.basename(s) in FACTORIES else []
        if s.endswith(".c"):
            r = subprocess.run(["gcc", "-O2", "-w", "-fPIC"] + [f for f in fl if f.startswith("-I")] + ["-c", s, "-o", o], capture_output=True, text=True)
        else:
            r = subprocess.run(["g++"] + fl + pre + ["-c", s, "-o", o], capture_output=True, text=True)
        if r.returncode:
            raise RuntimeError(f"g++ failed on {s}:\n{r.stderr[-3000:]}")

    with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        list(ex.map(run, jobs))
    return objs, bool(jobs)


def build(ref="/root/reference", force=False):
    if not available(ref):
        raise RuntimeError("the reference tree is not present: oracle/_ref is used as built")
    backend = os.path.join(HERE, "HostBackend.hpp")
    objs, changed = compile_objects(ref, OBJ, backend, force=force)
    if changed or not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(o) for o in objs):
        changed = True
        if os.path.exists(LIB):
            os.unlink(LIB)
        subprocess.run(["ar", "rcs", LIB] + objs, check=True)
    fl = flags(ref)
    exes = {"ref_dump": os.path.join(HERE, "ref_dump.cpp"), "ref_poisson_host": os.path.join(ref, "applications/001_Poisson/main.cpp"),
            "ref_amr_poisson_host": os.path.join(os.path.dirname(ORACLE), "tests", "cpp", "ref_amr_poisson.cpp"),
            "ref_stokes_host": os.path.join(os.path.dirname(ORACLE), "tests", "cpp", "ref_stokes.cpp"),
            "ref_partition_host": os.path.join(os.path.dirname(ORACLE), "tests", "cpp", "ref_partition.cpp")}
    for name, src in exes.items():
        if not os.path.exists(src):
            continue
        exe = os.path.join(OUT, name)
        if force or changed or not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(src), os.path.getmtime(backend)):
            r = subprocess.run(["g++"] + fl + ["-I" + ref, "-include", backend, src, "-o", exe, LIB, "-lpthread"], capture_output=True, text=True)
            if r.returncode:
                raise RuntimeError(f"link of {name} failed:\n{r.stderr[-4000:]}")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
