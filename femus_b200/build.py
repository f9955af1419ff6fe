"""Build femus_b200/libfemus_b200.so (CUDA kernels + C ABI) in-tree for sm_100a.

    python -m femus_b200.build [--force]

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libfemus_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
         "-Xcompiler", "-fPIC", "-ccbin", "/usr/bin/g++", "-Xptxas", "-v"]


def sources():
    out = []
    for root, _, files in os.walk(CSRC):
        for f in sorted(files):
            if f.endswith((".cu", ".cpp")):
                out.append(os.path.join(root, f))
    return out


def _deps_mtime():
    m = 0.0
    for root, _, files in os.walk(CSRC):
        for f in files:
            if f.endswith((".cuh", ".h", ".hpp")):
                m = max(m, os.path.getmtime(os.path.join(root, f)))
    inc = os.path.join(os.path.dirname(HERE), "include")
    for f in os.listdir(inc):
        m = max(m, os.path.getmtime(os.path.join(inc, f)))
    host = os.path.join(HERE, "host")           # b2h_capi.cpp includes the host layer's headers
    for f in os.listdir(host):
        if f.endswith(".hpp"):
            m = max(m, os.path.getmtime(os.path.join(host, f)))
    return m


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    hdr = _deps_mtime()
    jobs = []
    objs = []
    for src in sources():
        rel = os.path.relpath(src, CSRC).replace(os.sep, "_")
        obj = os.path.join(OBJ, rel + ".o")
        objs.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr):
            jobs.append((src, obj))

    def run(job):
        src, obj = job
        extra = []
        r = subprocess.run([NVCC] + FLAGS + extra + ["-c", src, "-o", obj], capture_output=True, text=True)
        log = r.stdout + r.stderr
        with open(obj + ".log", "w") as f:
            f.write(log)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{log}")
        return log

    with ThreadPoolExecutor(max_workers=8) as ex:
        logs = list(ex.map(run, jobs))
    if verbose:
        for l in logs:
            print(l)
    if jobs or not os.path.exists(LIB):
        r = subprocess.run([NVCC, "-shared", "-o", LIB] + objs + ["-ldl", "-ccbin", "/usr/bin/g++"],
                           capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    build_driver(force or bool(jobs))
    return LIB


DRIVER_SRC = os.path.join(os.path.dirname(HERE), "tests", "cpp", "poisson_driver.cpp")
DRIVER = os.path.join(HERE, "poisson_driver")      # in-tree: travels to the GPU box (femus_b200/build/ does not)
STOKES_DRIVER_SRC = os.path.join(os.path.dirname(HERE), "tests", "cpp", "stokes_driver.cpp")
STOKES_DRIVER = os.path.join(HERE, "stokes_driver")


def build_driver(force=False):
    """C++ drivers of the adapter classes (femus_b200/host/*.hpp), linked against the library."""
    deps = [os.path.join(HERE, "host", f) for f in os.listdir(os.path.join(HERE, "host")) if f.endswith(".hpp")]
    for src, exe in ((DRIVER_SRC, DRIVER), (STOKES_DRIVER_SRC, STOKES_DRIVER)):
        if not os.path.exists(src):
            continue
        newest = max(os.path.getmtime(d) for d in deps + [src])
        if force or not os.path.exists(exe) or os.path.getmtime(exe) < newest:
            r = subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-o", exe, src, "-L" + HERE, "-lfemus_b200",
                                "-Wl,-rpath,$ORIGIN"], capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError("adapter driver failed to build:\n" + r.stdout + r.stderr)
    return DRIVER


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
