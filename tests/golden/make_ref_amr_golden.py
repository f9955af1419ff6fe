"""Generate tests/golden/ref_amr_poisson.json: the residual norms and per-level solution norms the AMR application
tests/cpp/ref_amr_poisson.cpp prints when it runs on the reference's own mesh / solution / system sources with the
single-process HOST backend of oracle/ref_build (everything above the algebra is the reference's code: selective
refinement, hanging-node constraint matrices, KK <- Pamr^T KKamr Pamr, F-cycle).  The GPU test
tests/test_zz_reference_amr_gpu.py runs the same binary source on the B200 backend against these numbers.

Run in the build container (needs /root/reference):   python tests/golden/make_ref_amr_golden.py"""
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle.ref_build import build as rb  # noqa: E402

# nx, uniform levels, selective levels, cycle, level preconditioner, cycles
CASES = {
    "box2_1u2s_V_jacobi": ["2", "1", "2", "V", "jacobi", "6"],
    "box2_1u3s_F_sor": ["2", "1", "3", "F", "sor", "4"],
    "box4_1u2s_V_sor": ["4", "1", "2", "V", "sor", "6"],
    "box4_2u1s_V_jacobi": ["4", "2", "1", "V", "jacobi", "6"],
    # 21 568 elements, 186 337 dofs on the finest of 4 levels (2.6 minutes on the host backend, mostly its dense coarse LU)
    "box8_2u2s_V_jacobi": ["8", "2", "2", "V", "jacobi", "4"],
}


def parse(stdout):
    return {"residual_trace": [float(x) for x in re.findall(r"Linear Res\s+L2norm Sol\s*=\s*([0-9.eE+-]+)", stdout)],
            "elements": [int(x) for x in re.findall(r"Number of elements\s*:\s*(\d+)", stdout)],
            "levels": [[int(a), float(b), float(c)] for a, b, c in
                       re.findall(r"AMR level \d+ dofs (\d+) Sol l2 ([0-9.eE+-]+) linf ([0-9.eE+-]+)", stdout)]}


def run(exe, args):
    work = tempfile.mkdtemp(prefix="refamr_")
    try:
        os.makedirs(os.path.join(work, "input"))
        os.makedirs(os.path.join(work, "output"))
        env = dict(os.environ, GLIBC_TUNABLES="glibc.malloc.tcache_count=0")
        r = subprocess.run([exe] + args, cwd=work, env=env, capture_output=True, text=True, timeout=3600)
        if r.returncode:
            raise RuntimeError(f"{exe} {args}: rc {r.returncode}\n{r.stdout[-2000:]}\n{r.stderr[-2000:]}")
        return parse(r.stdout)
    finally:
        shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    rb.build()
    path = os.path.join(HERE, "ref_amr_poisson.json")
    out = json.load(open(path)) if os.path.exists(path) and "--all" not in sys.argv else {}
    for name, args in CASES.items():
        if name in out:
            continue
        out[name] = dict(run(os.path.join(rb.OUT, "ref_amr_poisson_host"), args), args=args)
        print(name, out[name]["elements"][:len(out[name]["levels"])], out[name]["residual_trace"])
    json.dump(out, open(os.path.join(HERE, "ref_amr_poisson.json"), "w"), indent=1)
