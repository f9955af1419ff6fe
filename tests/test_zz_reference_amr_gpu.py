"""GPU: the AMR row through the reference's OWN code on this backend.  tests/cpp/ref_amr_poisson.cpp --
an application against the reference's public API whose element loop and boundary function are those of
applications/001_Poisson/main.cpp compiled in place -- is linked with the reference's unmodified mesh / solution / system
sources and libfemus_b200.so (tests/ref_apps_build.py).  Selective refinement, Mesh::GetAMRRestrictionAndAMRSolidMark,
LinearImplicitSystem::BuildAmrProlongatorMatrix, KK <- Pamr^T KKamr Pamr on the non-homogeneous levels, the prolongators
multiplied by the constraint matrices and the V- / F-cycle all run as the reference wrote them; every matrix / vector /
solver object underneath is a device object (insert_row, get_transpose in place, matrix_RightMatMult, matrix_PtAP,
matrix_mult_transpose, SwapMatrices ...).  The printed residual norms and the per-level solution norms must equal what
the SAME application printed on the oracle's host backend (tests/golden/ref_amr_poisson.json, made by
tests/golden/make_ref_amr_golden.py) to the 7 digits printed."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900, method="thread")]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "femus_b200", "ref_amr_poisson_b200")
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_amr_poisson.json")))


@pytest.mark.parametrize("assembly", ["reference_callback", "device_callback"])
@pytest.mark.parametrize("case", sorted(GOLD))
def test_reference_amr_path_on_the_b200_backend(case, assembly, tmp_path):
    """assembly = "device_callback": the element loop itself runs on the GPU too -- femus::AssemblePoissonB200
    (femus_b200/host/RefAssemble.hpp: plans built from the reference's own Mesh / elem_type / boundary-function objects,
    b2_asm_poisson + b2_asm_neumann_faces on every level, the selectively refined ones included) is registered in place
    of 001_Poisson's callback; the golden numbers stay those of the reference's callback on the host backend."""
    if not os.path.exists(EXE):
        pytest.fail("femus_b200/ref_amr_poisson_b200 is missing: build it with `python tests/ref_apps_build.py` where /root/reference exists")
    from make_ref_amr_golden import parse
    g = GOLD[case]
    os.makedirs(tmp_path / "input")
    os.makedirs(tmp_path / "output")
    env = dict(os.environ, GLIBC_TUNABLES="glibc.malloc.tcache_count=0")
    args = g["args"] + (["device"] if assembly == "device_callback" else [])
    r = subprocess.run([EXE] + args, cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=800)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert ("femus::AssemblePoissonB200 (device)" in r.stdout) == (assembly == "device_callback")
    out = parse(r.stdout)
    assert out["elements"] == g["elements"]                                   # the selectively refined hierarchy itself
    assert [l[0] for l in out["levels"]] == [l[0] for l in g["levels"]]      # dofs per level
    trace, ref = np.array(out["residual_trace"]), np.array(g["residual_trace"])
    assert len(trace) == len(ref) and len(ref) >= 4
    # (an F-cycle's first entries are the round-off of the coarsest direct solve: absolute floor)
    assert np.all(np.abs(trace - ref) <= 2e-6 * ref + 1e-13), (trace, ref)
    for a, b in zip(out["levels"], g["levels"]):
        assert abs(a[1] - b[1]) <= 1e-6 * b[1] + 1e-13 and abs(a[2] - b[2]) <= 1e-6 * b[2] + 1e-13, (a, b)
