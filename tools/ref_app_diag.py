"""GPU box diagnostic: run femus_b200/ref_poisson_b200 (the reference's unmodified 001_Poisson on this backend) with
FEMUS_REF_DUMP set and compare, level by level, what the device objects hold with the arrays the SAME application wrote
on the oracle's host backend (tests/golden/ref_poisson_*.npz); then redo the first V-cycle in numpy from the dumped
operators (Richardson 0.5 around one symmetric SOR sweep, exact coarse solve) and compare EPS / RES after it.

    python tools/ref_app_diag.py box222_q1_3lev"""
import os
import re
import shutil
import subprocess
import sys
import tempfile

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
EXE = os.path.join(ROOT, "femus_b200", "ref_poisson_b200")


def load_dump(d):
    out = {}
    dt = {"i4": np.int32, "i8": np.int64, "f8": np.float64}
    for f in sorted(os.listdir(d)):
        m = re.match(r"L(\d+)_(\w+)\.(i4|i8|f8)$", f)
        out[f"L{m.group(1)}_{m.group(2)}"] = np.fromfile(os.path.join(d, f), dtype=dt[m.group(3)])
    return out


def csr(d, key):
    shape = tuple(int(x) for x in d[key + "_shape"])
    return sp.csr_matrix((d[key + "_val"], d[key + "_col"], d[key + "_rowptr"]), shape=shape)


def ssor(A, r):
    D = A.diagonal()
    Lo = sp.tril(A, 0).tocsr()
    Up = sp.triu(A, 0).tocsr()
    z = spla.spsolve_triangular(Lo, r, lower=True)
    return spla.spsolve_triangular(Up, D * z, lower=False)


def vcycle(levels, l, b, w=0.5):
    A, P = levels[l]
    if l == 0:
        return spla.spsolve(A.tocsc(), b)
    x = w * ssor(A, b)
    r = b - A @ x
    x = x + P @ vcycle(levels, l - 1, P.T @ r, w)
    return x + w * ssor(A, b - A @ x)


def main(case):
    g = np.load(os.path.join(GOLDEN, f"ref_poisson_{case}.npz"))
    work = tempfile.mkdtemp(prefix="refdiag_")
    for sub in ("input", "output", "dump"):
        os.makedirs(os.path.join(work, sub))
    for f in os.listdir(GOLDEN):
        if f.endswith(".neu"):
            shutil.copy(os.path.join(GOLDEN, f), os.path.join(work, "input", f))
    with open(os.path.join(work, "input", "in.json"), "w") as f:
        f.write(str(g["input_json"]))
    env = dict(os.environ, FEMUS_REF_DUMP=os.path.join(work, "dump"), GLIBC_TUNABLES="glibc.malloc.tcache_count=0")
    r = subprocess.run([EXE, "-i", "input/in.json"], cwd=work, env=env, capture_output=True, text=True, timeout=800)
    print("rc", r.returncode)
    trace = [float(x) for x in re.findall(r"Linear Res\s+L2norm Sol\s*=\s*([0-9.eE+-]+)", r.stdout)]
    print("trace b200", trace)
    print("trace ref ", list(g["residual_trace"]))
    d = load_dump(os.path.join(work, "dump"))
    nlev = 1 + max(int(k[1]) for k in d)
    levels = []
    for l in range(nlev):
        A = csr(d, f"L{l}_KK")
        P = csr(d, f"L{l}_PP") if f"L{l}_PP_val" in d else None
        print(f"level {l}: KK {A.shape} nnz {A.nnz}", end="")
        if f"L{l}_KK_val" in g.files:
            G = csr(g, f"L{l}_KK")
            print(f"  pattern equal {np.array_equal(G.indptr, A.indptr) and np.array_equal(G.indices, A.indices)}"
                  f"  max|KK - ref| {abs(A - G).max():.3e} of {abs(G).max():.3e}", end="")
        elif f"L{l}_KK_diag" in g.files:
            print(f"  max|diag - ref| {np.abs(A.diagonal() - g[f'L{l}_KK_diag']).max():.3e}"
                  f"  max|rowsum - ref| {np.abs(np.asarray(A.sum(1)).ravel() - g[f'L{l}_KK_rowsum']).max():.3e}", end="")
        if P is not None:
            G = csr(g, f"L{l}_PP")
            print(f"  max|PP - ref| {abs(P - G).max():.3e} (nnz {P.nnz} / {G.nnz})", end="")
        print(f"  bdcIndex equal {np.array_equal(d[f'L{l}_bdcIndex'], g[f'L{l}_bdcIndex'])}")
        # penalty, as MGSetLevel applies it
        A = A.tolil()
        bdc = d[f"L{l}_bdcIndex"]
        for i in bdc:
            A.rows[i] = [int(i)]
            A.data[i] = [1.0]
        levels.append((A.tocsr(), P))
    top = nlev - 1
    res = d[f"L{top}_RES"].copy()
    if f"L{top}_RES" in g.files:
        print(f"RES: max|RES - ref| {np.abs(res - g[f'L{top}_RES']).max():.3e} of {np.abs(res).max():.3e}")
    res[d[f"L{top}_bdcIndex"]] = 0.0
    x = vcycle(levels, top, res)
    res_after = res - levels[top][0] @ x
    print(f"numpy cycle from the dumped operators: |RES_after| {np.linalg.norm(res_after):.6e}")
    print(f"device: |RES_after| {np.linalg.norm(d[f'L{top}_RES_after']):.6e}   max|EPS - numpy| {np.abs(d[f'L{top}_EPS_after'] - x).max():.3e} of {np.abs(x).max():.3e}")
    # which half differs: smoother only / coarse correction only
    A, P = levels[top]
    x1 = 0.5 * ssor(A, res)
    print(f"  after the pre-smoothing sweep alone the numpy residual would be {np.linalg.norm(res - A @ x1):.6e}")
    xj = 0.5 * res / A.diagonal()
    print(f"  (a Jacobi sweep instead: {np.linalg.norm(res - A @ xj):.6e})")
    shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "box222_q1_3lev")
