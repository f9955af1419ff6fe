"""TEST INFRASTRUCTURE (it compiles reference objects shared with the oracle's host build).  The reference's UNMODIFIED application applications/001_Poisson/main.cpp on the femus_b200 backend:

    femus_b200/ref_poisson_b200    main.cpp + the reference's own mesh / solution / system sources (compiled where they
                                   lie under /root/reference, objects shared with oracle/ref_build) + the three factory
                                   translation units compiled with femus_b200/host/RefBackend.hpp pre-included, linked
                                   against libfemus_b200.so

    femus_b200/ref_amr_poisson_b200    tests/cpp/ref_amr_poisson.cpp the same way: selectively refined meshes through the
                                   reference's own AMR path (hanging-node constraint matrices, non-homogeneous levels)

Built in the container that has /root/reference (python tests/ref_apps_build.py); the binary is in-tree (git-ignored,
like the library) so that it travels to the GPU box, where tests/test_zz_reference_app_gpu.py runs it."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HERE = os.path.join(ROOT, "femus_b200")       # the binaries sit next to libfemus_b200.so (rpath $ORIGIN)
EXE = os.path.join(HERE, "ref_poisson_b200")
AMR_EXE = os.path.join(HERE, "ref_amr_poisson_b200")


def build(ref="/root/reference", force=False):
    sys.path.insert(0, ROOT)
    from oracle.ref_build import build as rb
    from femus_b200 import build as lib
    if not rb.available(ref):
        raise RuntimeError("the reference tree is not present")
    lib.build()
    backend = os.path.join(HERE, "host", "RefBackend.hpp")
    objdir = os.path.join(rb.OUT, "obj_b200")
    # every object but the three factories is backend-independent: compile those once (oracle/ref_build) and reuse them
    host_objs, _ = rb.compile_objects(ref, rb.OBJ, os.path.join(rb.HERE, "HostBackend.hpp"))
    os.makedirs(objdir, exist_ok=True)
    fl = rb.flags(ref) + ["-I" + os.path.join(HERE, "host"), "-I" + os.path.join(ROOT, "include")]
    objs = []
    for s, o in zip(rb.reference_sources(ref), host_objs):
        if os.path.basename(s) in rb.FACTORIES:
            o2 = os.path.join(objdir, os.path.basename(o))
            deps = [os.path.join(HERE, "host", f) for f in os.listdir(os.path.join(HERE, "host")) if f.endswith(".hpp")] + [s]
            if force or not os.path.exists(o2) or os.path.getmtime(o2) < max(os.path.getmtime(d) for d in deps):
                r = subprocess.run(["g++"] + fl + ["-include", backend, "-c", s, "-o", o2], capture_output=True, text=True)
                if r.returncode:
                    raise RuntimeError(f"g++ failed on {s}:\n{r.stderr[-4000:]}")
            objs.append(o2)
        else:
            objs.append(o)
    apps = {EXE: os.path.join(ref, "applications/001_Poisson/main.cpp"),            # the reference's application, unmodified
            AMR_EXE: os.path.join(ROOT, "tests", "cpp", "ref_amr_poisson.cpp")}      # the reference's AMR path (see the file header)
    for exe, main in apps.items():
        r = subprocess.run(["g++"] + fl + ["-I" + ref, "-DB2_REF_DEVICE_ASSEMBLY", "-include", backend, main, "-o", exe] + objs + ["-L" + HERE, "-lfemus_b200", "-Wl,-rpath,$ORIGIN", "-lpthread"],
                           capture_output=True, text=True)
        if r.returncode:
            raise RuntimeError(f"link of {os.path.basename(exe)} failed:\n{r.stderr[-4000:]}")
    return EXE


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
