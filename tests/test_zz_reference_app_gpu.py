"""GPU: the reference's UNMODIFIED application applications/001_Poisson/main.cpp, compiled with the reference's own
mesh / solution / system sources, running on libfemus_b200.so: every NumericVector / SparseMatrix /
LinearEquationSolver the application touches is a device object of this backend behind the reference's factories
(femus_b200/host/RefBackend.hpp, tests/ref_apps_build.py).  The residual norms the reference prints after every
V-cycle (LinearImplicitSystem.cpp:426) must equal, to the 7 digits printed, what the SAME application printed on the
host backend of the oracle build (tests/golden/ref_poisson_*.npz).  The binary is built where /root/reference exists
(__graft_entry__.build()) and travels in-tree."""
import json
import os
import re
import shutil
import subprocess

import numpy as np
import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900, method="thread")]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
EXE = os.path.join(ROOT, "femus_b200", "ref_poisson_b200")


@pytest.mark.parametrize("case", ["box222_q2_3lev", "box222_q1_3lev", "box324_q2_2lev_neumann", "cube_hex_q2_2lev", "cube_tet_q2_2lev", "cube_mixed_q2_2lev"])
def test_unmodified_reference_application_on_the_b200_backend(case, tmp_path):
    if not os.path.exists(EXE):
        pytest.fail("femus_b200/ref_poisson_b200 is missing: build it with `python tests/ref_apps_build.py` where /root/reference exists")
    g = np.load(os.path.join(GOLDEN, f"ref_poisson_{case}.npz"))
    work = str(tmp_path)
    os.makedirs(os.path.join(work, "input"))
    os.makedirs(os.path.join(work, "output"))
    for f in os.listdir(GOLDEN):
        if f.endswith(".neu"):
            shutil.copy(os.path.join(GOLDEN, f), os.path.join(work, "input", f))
    with open(os.path.join(work, "input", "in.json"), "w") as f:
        f.write(str(g["input_json"]))
    # (the reference's JSON "box" branch iterates a temporary std::map: see tests/golden/make_ref_golden.py)
    env = dict(os.environ, GLIBC_TUNABLES="glibc.malloc.tcache_count=0")
    r = subprocess.run([EXE, "-i", "input/in.json"], cwd=work, env=env, capture_output=True, text=True, timeout=800)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    trace = np.array([float(x) for x in re.findall(r"Linear Res\s+L2norm Sol\s*=\s*([0-9.eE+-]+)", r.stdout)])
    ref = g["residual_trace"]
    assert len(trace) == len(ref) == 6
    assert np.all(np.abs(trace - ref) <= 2e-6 * ref), (trace, ref)
