"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle on the same
seeded inputs.  Bar: integer structure bit-exact; fp64 values within the stated tolerance
(differences come only from summation order / FMA contraction)."""
import numpy as np
import pytest
import scipy.sparse as sp

from femus_b200 import capi
from oracle import fe_hex, mesh_box as mb, mg

pytestmark = pytest.mark.gpu

RTOL = 1e-12        # north-star: matrix entries / residuals to 1e-12 relative


def relerr(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


# ------------------------------------------------------------------------------ vectors
@pytest.mark.parametrize("n", [1, 2, 31, 1000, 100003])
def test_vector_ops(ctx, n):
    rng = np.random.default_rng(n)
    x, y = rng.standard_normal(n), rng.standard_normal(n)
    X, Y, W = ctx.vector(x), ctx.vector(y), ctx.vector(n)
    Y.axpy(0.75, X)
    ref = y + 0.75 * x
    assert relerr(Y.get(), ref) < 1e-15
    Y.aypx(-2.0, X)
    ref = x - 2.0 * ref
    assert relerr(Y.get(), ref) < 1e-15
    Y.scale(3.0)
    ref *= 3.0
    Y.add_scalar(0.5)
    ref += 0.5
    assert relerr(Y.get(), ref) < 1e-15
    W.pointwise_mult(X, Y)
    assert relerr(W.get(), x * ref) < 1e-15
    assert abs(X.dot(Y) - x @ ref) <= 1e-13 * np.abs(x * ref).sum()
    assert abs(X.norm(2) - np.linalg.norm(x)) <= 1e-14 * np.linalg.norm(x)
    assert abs(X.norm(1) - np.abs(x).sum()) <= 1e-14 * np.abs(x).sum()
    assert X.norm(0) == np.abs(x).max()
    assert abs(X.sum() - x.sum()) <= 1e-13 * np.abs(x).sum()
    assert X.minmax() == (x.min(), x.max())
    W.fill(2.5)
    assert np.all(W.get() == 2.5)
    W.zero()
    assert np.all(W.get() == 0.0)
    W.copy_from(X)
    assert np.array_equal(W.get(), x)
    mask = ctx.vector((rng.random(n) > 0.5) * 2.0)
    W.copy_masked(X, mask, 1.1)
    assert np.array_equal(W.get(), np.where(mask.get() > 1.1, x, 0.0))


def test_vector_indexed(ctx):
    n = 5000
    rng = np.random.default_rng(0)
    x = rng.standard_normal(n)
    X = ctx.vector(x)
    idx = rng.permutation(n)[:700].astype(np.int32)
    v = rng.standard_normal(700)
    X.set_indexed(idx, v)
    x[idx] = v
    assert np.array_equal(X.get(), x)
    idx2 = rng.integers(0, n, 3000).astype(np.int32)      # duplicates: add semantics
    v2 = rng.standard_normal(3000)
    X.add_indexed(idx2, v2)
    np.add.at(x, idx2, v2)
    assert relerr(X.get(), x) < 1e-14
    X.fill_indexed(idx, 0.0)
    x[idx] = 0.0
    assert relerr(X.get(), x) < 1e-14
    assert np.array_equal(X.get_indexed(idx2), X.get()[idx2])
    # duplicates in CALL ORDER (VecSetValues): the last value set stays -- MultiLevelSolution::GenerateBdc writes 2., then
    # 1., then 0. to one entry before close() (MultiLevelSolution.cpp:737-835); sums of a zeroed entry are bit-exact
    idx3 = rng.integers(0, n, 4000).astype(np.int32)
    v3 = rng.standard_normal(4000)
    x = rng.standard_normal(n)
    X = ctx.vector(x)
    X.set_indexed(idx3, v3)
    for i, a in zip(idx3, v3):
        x[i] = a
    assert np.array_equal(X.get(), x)
    X.zero()
    X.add_indexed(idx3, v3)
    x[:] = 0.0
    for i, a in zip(idx3, v3):
        x[i] += a
    assert np.array_equal(X.get(), x)


def test_empty_vector(ctx):
    X = ctx.vector(0)
    X.zero()
    X.scale(2.0)
    assert X.get().shape == (0,)


# ------------------------------------------------------------------------------ CSR algebra
def random_csr(rng, m, n, density):
    A = sp.random(m, n, density=density, random_state=np.random.RandomState(int(rng.integers(1 << 30))), format="csr")
    A.sort_indices()
    return A


@pytest.mark.parametrize("m,n,density", [(1, 1, 1.0), (50, 70, 0.02), (3000, 3000, 0.001), (2000, 1500, 0.05), (400, 400, 0.4)])
def test_spmv_family(ctx, m, n, density):
    rng = np.random.default_rng(m * 7 + n)
    A = random_csr(rng, m, n, density)
    # ragged input: some empty rows
    A = A.tolil()
    if m > 10:
        A[3, :] = 0
    A = A.tocsr()
    A.eliminate_zeros()
    A.sort_indices()
    dA = ctx.csr_from_scipy(A)
    x, b = rng.standard_normal(n), rng.standard_normal(m)
    X, Bv, Y = ctx.vector(x), ctx.vector(b), ctx.vector(m)
    scale = np.abs(A) @ np.abs(x) + 1e-300
    dA.spmv(X, Y)
    assert np.all(np.abs(Y.get() - A @ x) <= 4e-16 * scale * 8)
    Y.put(b)
    dA.spmv_add(X, Y)
    assert np.all(np.abs(Y.get() - (b + A @ x)) <= 1e-14 * (scale + np.abs(b)))
    dA.resid(Bv, X, Y)
    assert np.all(np.abs(Y.get() - (b - A @ x)) <= 1e-14 * (scale + np.abs(b)))
    xt = rng.standard_normal(m)
    XT, YT = ctx.vector(xt), ctx.vector(n)
    dA.spmv_t(XT, YT)
    assert np.all(np.abs(YT.get() - A.T @ xt) <= 1e-13 * (np.abs(A.T) @ np.abs(xt) + 1e-300))


@pytest.mark.parametrize("variant", [0, 1, 2])
@pytest.mark.parametrize("kind", ["fe_like", "long_rows", "mostly_empty", "short_rows"])
def test_spmv_kernels_and_edge_shapes(ctx, variant, kind):
    """Every SpMV kernel (register-streaming, TMA-staged, staged + pipelined gathers) on shapes that
    stress the row chunking: FE-like banded rows, rows too long to stage (streaming fallback), long
    runs of empty rows, very short rows; all epilogues including r aliased with b (MGSolve does it)."""
    rng = np.random.default_rng(11)
    if kind == "fe_like":
        m = n = 6000
        offs = sorted(set(int(o) for o in rng.integers(-300, 300, 60)))
        A = sp.dia_matrix((rng.standard_normal((len(offs), n)), offs), shape=(m, n)).tocsr()
    elif kind == "long_rows":
        m, n = 40, 5000
        A = random_csr(rng, m, n, 0.5)                      # ~2500 entries per row
    elif kind == "mostly_empty":
        m, n = 5000, 300
        A = sp.lil_matrix((m, n))
        for r in (0, 7, 2500, 2501, 4999):
            A[r, rng.integers(0, n, 40)] = rng.standard_normal(40)
        A = A.tocsr()
    else:
        m, n = 20000, 9000
        A = random_csr(rng, m, n, 3.0 / n)
    A = A.tocsr()
    A.sort_indices()
    ctx.set_option("spmv_variant", variant)
    try:
        dA = ctx.csr_from_scipy(A)
        x, b, dinv = rng.standard_normal(n), rng.standard_normal(m), rng.random(m) + 0.5
        X, Bv, Y = ctx.vector(x), ctx.vector(b), ctx.vector(m)
        scale = np.abs(A) @ np.abs(x) + np.abs(b) + 1e-300
        dA.spmv(X, Y)
        assert np.all(np.abs(Y.get() - A @ x) <= 1e-14 * scale)
        Y.put(b)
        dA.spmv_add(X, Y)
        assert np.all(np.abs(Y.get() - (b + A @ x)) <= 1e-14 * scale)
        dA.resid(Bv, X, Y)
        assert np.all(np.abs(Y.get() - (b - A @ x)) <= 1e-14 * scale)
        R = ctx.vector(b)
        dA.resid(R, X, R)                                   # r aliased with b
        assert np.all(np.abs(R.get() - (b - A @ x)) <= 1e-14 * scale)
        if m == n:
            D = ctx.vector(dinv)
            dA.jacobi_sweep(D, Bv, X, Y, 0.5)
            assert np.all(np.abs(Y.get() - (x + 0.5 * dinv * (b - A @ x))) <= 1e-14 * (scale + np.abs(x)))
    finally:
        ctx.set_option("spmv_variant", 1)


@pytest.mark.parametrize("order,shape,nl", [("biquadratic", (2, 2, 2), 2), ("biquadratic", (3, 2, 4), 3), ("linear", (2, 3, 2), 3)])
def test_fused_assembly_galerkin_matches_separate(ctx, order, shape, nl, asm_variant):
    """b2_asm_poisson_galerkin: same fine matrix and residual as b2_asm_poisson, and its coarse matrix
    equals both the element-gather product and the oracle's scipy P^T A P (Dirichlet rows/columns of
    P zeroed), on a deformed mesh with a non-zero solution."""
    from femus_b200.poisson import PoissonMG
    from oracle import mesh_box as mb, mg
    pf = PoissonMG(ctx, *shape, nl, order, fused=True)
    ps = PoissonMG(ctx, *shape, nl, order, fused=False)
    rng = np.random.default_rng(3)
    sol = rng.standard_normal(pf.n)
    for pb in (pf, ps):
        pb.SOL.put(sol)
        pb.assemble()
        pb.galerkin()
    Af, As = pf.KK[-1].to_scipy(), ps.KK[-1].to_scipy()
    # (fp64 atomics: the order of the element contributions differs from run to run)
    assert np.array_equal(Af.indices, As.indices)
    assert np.abs(Af.data - As.data).max() <= 1e-14 * np.abs(As.data).max()
    assert np.abs(pf.RES.get() - ps.RES.get()).max() <= 1e-13 * np.abs(ps.RES.get()).max()
    for l in range(nl - 1):
        Cf, Cs = pf.KK[l].to_scipy(), ps.KK[l].to_scipy()
        assert np.array_equal(Cf.indices, Cs.indices)
        assert np.abs(Cf.data - Cs.data).max() <= 1e-13 * np.abs(Cs.data).max()
    # oracle: P^T A P with scipy on the oracle's own hierarchy
    lv = mb.build_hierarchy(*shape, nl)
    H = mg.Hierarchy(lv, order, A_top=As, rhs=np.zeros(pf.n), coarse_lu=False)
    for l in range(nl - 1):
        C, Co = pf.KK[l].to_scipy(), H.A_raw[l]
        assert np.array_equal(C.indptr, Co.indptr) and np.array_equal(C.indices, Co.indices)
        assert np.abs(C.data - Co.data).max() <= 1e-12 * np.abs(Co.data).max()
    del pf, ps


def test_transpose_zero_rows_cols_diag(ctx):
    rng = np.random.default_rng(5)
    A = random_csr(rng, 700, 500, 0.03)
    dA = ctx.csr_from_scipy(A)
    T = dA.transpose().to_scipy()
    At = A.T.tocsr()
    At.sort_indices()
    assert np.array_equal(T.indptr, At.indptr) and np.array_equal(T.indices, At.indices)
    assert np.array_equal(T.data, At.data)
    S = random_csr(rng, 600, 600, 0.02) + sp.identity(600, format="csr") * 3.0
    S = S.tocsr()
    S.sort_indices()
    dS = ctx.csr_from_scipy(S)
    d = ctx.vector(600)
    dS.diag(d)
    assert np.array_equal(d.get(), S.diagonal())
    rows = np.sort(rng.permutation(600)[:50]).astype(np.int32)
    dS.zero_rows(rows, 1.0)
    ref = mg.penalty_fast(S, rows)
    got = dS.to_scipy()
    assert np.array_equal(got.indices, ref.indices) and np.array_equal(got.data, ref.data)
    cols = np.sort(rng.permutation(500)[:40]).astype(np.int32)
    dA.zero_cols(cols)
    ref = A.copy()
    ref.data[np.isin(ref.indices, cols)] = 0.0
    assert np.array_equal(dA.to_scipy().data, ref.data)


def test_add_blocks_and_set_rows(ctx):
    """Compat path of add_matrix_blocked / insert_row against the oracle's scatter."""
    lv = mb.build_hierarchy(2, 2, 2, 1)
    L = lv[0]
    d = mb.system_dof(L, "biquadratic")
    rp, ci = mb.sparsity(L, "biquadratic")
    A = ctx.csr(rp.shape[0] - 1, rp.shape[0] - 1, rp, ci)
    X = L.xyz[:, L.conn].transpose(1, 0, 2)
    F, B = fe_hex.poisson_elements("biquadratic", X, np.zeros((L.nel, 27)))
    A.add_blocks(d, d, B.reshape(L.nel, -1))
    ref, _ = mb.assemble(L, "biquadratic")
    got = A.to_scipy()
    assert np.array_equal(got.indices, ref.indices)
    assert np.abs(got.data - ref.data).max() <= RTOL * np.abs(ref.data).max()
    with pytest.raises(capi.B2Error):
        bad = d.copy()
        A2 = ctx.csr(3, 3, np.array([0, 1, 2, 3]), np.array([0, 1, 2]))
        A2.add_blocks(np.array([[0, 1]]), np.array([[0, 1]]), np.ones((1, 4)))
    # insert_row
    P = mb.prolongator(lv[0], mb.refine(lv[0]), "linear").tocsr()
    dP = ctx.csr(P.shape[0], P.shape[1], P.indptr, P.indices)
    rows = np.arange(P.shape[0], dtype=np.int32)
    dP.set_rows(rows, P.indptr, P.indices, P.data)
    assert np.array_equal(dP.to_scipy().data, P.data)


# ------------------------------------------------------------------------------ pattern + assembly
@pytest.mark.parametrize("order,shape", [("linear", (3, 2, 4)), ("biquadratic", (3, 2, 2)), ("biquadratic", (4, 4, 4))])
def test_pattern_from_elements_bit_exact(ctx, order, shape):
    L = mb.build_box(*shape)
    d = mb.system_dof(L, order)
    rp, ci = mb.sparsity(L, order)
    A = capi.Csr.from_elements(ctx, mb.ndofs(L, order), d)
    grp, gci, _ = A.get(values=False)
    assert np.array_equal(grp, rp)
    assert np.array_equal(gci, ci)


def distorted(L, amp, seed):
    """Smoothly distorted copy of the box so that the Jacobian varies inside every element."""
    x, y, z = L.xyz
    rng = np.random.default_rng(seed)
    a = rng.uniform(0.5, 1.5, 3)
    out = L.xyz.copy()
    out[0] += amp * np.sin(a[0] * np.pi * y) * np.sin(np.pi * z) * x * (1 - x)
    out[1] += amp * np.sin(a[1] * np.pi * z) * np.sin(np.pi * x) * y * (1 - y)
    out[2] += amp * np.sin(a[2] * np.pi * x) * np.sin(np.pi * y) * z * (1 - z)
    return out


@pytest.fixture(params=[3, 1, 0], ids=["sumfac", "tensorcore", "cudacore"])
def asm_variant(request, ctx):
    """The three triquadratic assembly kernels: sum factorisation (default), FP64 tensor cores, CUDA-core register tiles."""
    ctx.set_option("asm_variant", request.param)
    yield request.param
    ctx.set_option("asm_variant", 3)


@pytest.mark.parametrize("order", ["linear", "biquadratic"])
@pytest.mark.parametrize("amp", [0.0, 0.08])
def test_assembly_matches_oracle(ctx, order, amp, asm_variant):
    lv = mb.build_hierarchy(2, 3, 2, 2)
    L = lv[-1]
    L.xyz = distorted(L, amp, 3)
    n = mb.ndofs(L, order)
    d = mb.system_dof(L, order)
    rng = np.random.default_rng(11)
    u = rng.standard_normal(n)
    Aref, rhs_ref = mb.assemble(L, order, u, fsrc=1.5)
    A = capi.Csr.from_elements(ctx, n, d)
    mesh = capi.Mesh(ctx, L.xyz, L.conn)
    asm = capi.Assembler(mesh, A, d, fe_hex.tables(order))
    U, R = ctx.vector(u), ctx.vector(n)
    asm.poisson(U, R, nu=1.0, fsrc=1.5)
    got = A.to_scipy()
    assert np.array_equal(got.indptr, Aref.indptr) and np.array_equal(got.indices, Aref.indices)
    assert np.abs(got.data - Aref.data).max() <= RTOL * np.abs(Aref.data).max()
    # residual: tolerance relative to the magnitude of the summed terms
    mag = np.abs(Aref) @ np.abs(u) + np.abs(rhs_ref)
    assert np.all(np.abs(R.get() - rhs_ref) <= RTOL * mag.max())
    # accumulate semantics: a second pass doubles the matrix (the app zeroes it first)
    asm.poisson(U, None, nu=1.0, fsrc=1.5)
    assert np.abs(A.to_scipy().data - 2 * Aref.data).max() <= 2 * RTOL * np.abs(Aref.data).max()


@pytest.mark.parametrize("order", ["linear", "biquadratic"])
def test_neumann_faces_match_oracle(ctx, order):
    """b2_asm_neumann on a distorted mesh against the oracle's restatement of main.cpp:495-548
    (elem_type_2D::JacobianSur pinned bit-exactly to the compiled reference), two flux values."""
    from femus_b200 import hostapi
    from femus_b200.poisson import PoissonMG
    neumann = {3: 0.2, 6: -1.5}
    pb = PoissonMG(ctx, 2, 3, 2, 2, order, dirichlet_faces=(1,), neumann=neumann, fsrc=0.0)
    lv = mb.build_hierarchy(2, 3, 2, 2)
    L = lv[-1]
    L.xyz = distorted(L, 0.06, 2)
    xyz_host = np.ascontiguousarray(L.xyz)
    pb.mesh.update(xyz_host.ctypes.data, None)
    ctx.sync()
    pb.assemble()
    _, rhs0 = mb.assemble(L, order, None, 0.0)
    ref_rhs = rhs0 + mb.neumann_rhs(L, order, neumann)
    got = pb.RES.get()
    assert np.abs(got - ref_rhs).max() <= 1e-13 * np.abs(ref_rhs).max()
    assert abs(got.sum() - (mb.neumann_rhs(L, order, neumann)).sum()) <= 1e-12
    del pb


@pytest.mark.parametrize("order", ["linear", "biquadratic"])
def test_vcycle_trace_with_neumann_and_partial_dirichlet(ctx, order):
    """The shipped 3-D input of applications/001_Poisson (input3D_Hex_first.json): source 0, Dirichlet on
    "top" only, Neumann flux 0.2 on "right", natural elsewhere; residual trace against the oracle."""
    from femus_b200.poisson import PoissonMG
    from oracle import mg
    neumann, dirichlet = {3: 0.2}, (6,)
    pb = PoissonMG(ctx, 2, 2, 2, 3, order, dirichlet_faces=dirichlet, neumann=neumann, fsrc=0.0, coarse_rtol=1e-15)
    pb.assemble(); pb.galerkin(); pb.mg_set_levels()
    lv = mb.build_hierarchy(2, 2, 2, 3)
    H = mg.Hierarchy(lv, order, fsrc=0.0, dirichlet_faces=dirichlet, neumann=neumann)
    trace, eps = H.mg_solve_trace(3)
    r0 = float(np.linalg.norm(np.where(H.bdc[-1] > 1.1, H.rhs, 0.0)))
    assert abs(pb.residual_norm() - r0) <= 1e-12 * r0
    for k in range(3):
        pb.mg_solve()
        assert abs(pb.residual_norm() - trace[k]) <= 1e-11 * r0, (k, pb.residual_norm(), trace[k])
    assert np.abs(pb.EPS.get() - eps).max() <= 1e-10 * np.abs(eps).max()
    del pb


@pytest.mark.parametrize("order,npre", [("linear", 2), ("biquadratic", 3)])
def test_vcycle_trace_chebyshev_smoother(ctx, order, npre):
    """Chebyshev + Jacobi smoothing (KSPCHEBYSHEV + PCJACOBI) with our stated eigenvalue bounds
    ([0.1, 1.1] x a 10-step power-iteration estimate from a fixed start vector): bounds and V-cycle
    residual trace against the oracle."""
    from femus_b200.poisson import PoissonMG
    from oracle import mg
    pb = PoissonMG(ctx, 2, 2, 2, 3, order, npre=npre, npost=npre, smoother="chebyshev", coarse_rtol=1e-15)
    pb.assemble(); pb.galerkin(); pb.mg_set_levels()
    lv = mb.build_hierarchy(2, 2, 2, 3)
    H = mg.Hierarchy(lv, order, smoother="chebyshev")
    for l in (1, 2):
        a, b = pb.mg.level_bounds(l)
        assert abs(a - H.ebounds[l][0]) <= 1e-12 * a and abs(b - H.ebounds[l][1]) <= 1e-12 * b
    trace, eps = H.mg_solve_trace(3, npre, npre)
    r0 = float(np.linalg.norm(np.where(H.bdc[-1] > 1.1, H.rhs, 0.0)))
    for k in range(3):
        pb.mg_solve()
        assert abs(pb.residual_norm() - trace[k]) <= 1e-11 * r0, (k, pb.residual_norm(), trace[k])
    assert np.abs(pb.EPS.get() - eps).max() <= 1e-10 * np.abs(eps).max()
    if order == "biquadratic":      # degree 3 Chebyshev smoothing contracts faster than 3 damped-Jacobi sweeps
        Hj = mg.Hierarchy(lv, order)
        tj, _ = Hj.mg_solve_trace(3, npre, npre)
        assert trace[-1] < tj[-1]
    del pb


def test_assembly_golden_elements(ctx, asm_variant):
    """Single elements of the committed golden fixture (values of the compiled reference)."""
    import os
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "fe_hex_ref.npz"))
    for order, nve in (("linear", 8), ("biquadratic", 27)):
        for k in range(G[f"{order}_X"].shape[0]):
            X = G[f"{order}_X"][k]
            xyz = np.zeros((3, 27))
            xyz[:, :nve] = X
            conn = np.arange(27, dtype=np.int32)[None, :]
            d = np.arange(nve, dtype=np.int32)[None, :]
            A = capi.Csr.from_elements(ctx, nve, d)
            asm = capi.Assembler(capi.Mesh(ctx, xyz, conn), A, d, fe_hex.tables(order))
            U, R = ctx.vector(G[f"{order}_U"][k]), ctx.vector(nve)
            asm.poisson(U, R, 1.0, 1.0)
            B = A.to_scipy().toarray()
            Bref, Fref = G[f"{order}_B"][k], G[f"{order}_F"][k]
            assert np.abs(B - Bref).max() <= RTOL * np.abs(Bref).max()
            assert np.abs(R.get() - Fref).max() <= RTOL * (np.abs(Bref) @ np.abs(G[f"{order}_U"][k])).max()


# ------------------------------------------------------------------------------ Galerkin + V-cycle
@pytest.mark.parametrize("order", ["linear", "biquadratic"])
def test_ptap_matches_oracle(ctx, order):
    lv = mb.build_hierarchy(2, 2, 3, 3)
    H = mg.Hierarchy(lv, order)
    for l in (2, 1):
        A = ctx.csr_from_scipy(H.A_raw[l])
        P = ctx.csr_from_scipy(H.P[l])
        rp, ci = mb.sparsity(lv[l - 1], order)
        C = ctx.csr(rp.shape[0] - 1, rp.shape[0] - 1, rp, ci)
        C.ptap(P, A)
        got = C.to_scipy()
        ref = H.A_raw[l - 1]
        assert np.array_equal(got.indices, ref.indices)
        assert np.abs(got.data - ref.data).max() <= RTOL * np.abs(ref.data).max()


@pytest.mark.parametrize("shape", [(40, 30, 50, 0.2), (1, 1, 1, 1.0), (300, 257, 129, 0.03), (64, 64, 64, 0.0)])
def test_matmat_and_axpy_match_scipy(ctx, shape):
    """General products / sums of the AMR path (matrix_RightMatMult, matrix_LeftMatMult, matrix_ABC, matrix_add):
    structure identical to scipy's (structural zeros kept), values to round-off, run-to-run bit-identical."""
    m, k, n, dens = shape
    rng = np.random.default_rng(m + n)
    A, B = random_csr(rng, m, k, dens), random_csr(rng, k, n, dens)
    dA, dB = ctx.csr_from_scipy(A), ctx.csr_from_scipy(B)
    C = dA.matmat(dB)
    got = C.to_scipy()
    # scipy drops nothing structurally either: pattern of the product of the patterns
    ref = (A @ B).tocsr()
    pat = (abs(A).sign() @ abs(B).sign()).tocsr()
    pat.sort_indices()
    assert got.shape == ref.shape and np.array_equal(got.indptr, pat.indptr) and np.array_equal(got.indices, pat.indices)
    assert abs(got - ref).max() <= 1e-14 * max(abs(ref).max(), 1e-300) if ref.nnz else got.nnz == 0
    again = dA.matmat(dB).to_scipy()
    assert np.array_equal(again.data, got.data)
    if m == k == n:
        # Y += a X: the pattern of X inside the pattern of Y
        Y = (A + B).tocsr()
        Y.sort_indices()
        dY = ctx.csr_from_scipy(Y)
        assert dY.pattern_contains(dA) and not ctx.csr_from_scipy(sp.csr_matrix((m, m))).pattern_contains(ctx.csr_from_scipy(sp.eye(m, format="csr")))
        dY.axpy(-0.5, dA)
        assert abs(dY.to_scipy() - (Y - 0.5 * A)).max() <= 1e-15 * max(abs(Y).max(), 1.0)
        with pytest.raises(capi.B2Error):
            ctx.csr_from_scipy(sp.csr_matrix((m, m))).axpy(1.0, ctx.csr_from_scipy(sp.eye(m, format="csr")))


@pytest.mark.parametrize("order", ["linear", "biquadratic"])
@pytest.mark.parametrize("shape", [(2, 2, 3), (1, 1, 1), (3, 1, 2)])
def test_galerkin_element_gather_matches_oracle(ctx, order, shape):
    """Fast path of matrix_PtAP (coarse-element gather) against the oracle's scipy P^T A P and
    against the general device triple product, with the Dirichlet rows/columns of P zeroed."""
    from femus_b200 import hostapi
    nl = 3
    hh = hostapi.HostHierarchy(*shape, nl)
    lv = mb.build_hierarchy(*shape, nl)
    H = mg.Hierarchy(lv, order)
    ploc, fent = hostapi.galerkin_element(order)
    for l in (2, 1):
        A = ctx.csr_from_scipy(H.A_raw[l])
        rp, ci = mb.sparsity(lv[l - 1], order)
        C = ctx.csr(rp.shape[0] - 1, rp.shape[0] - 1, rp, ci)
        fd, val = hh.galerkin_maps(l - 1, order)
        # host layer == oracle on the element -> dof maps the plan is built from
        assert np.array_equal(hh.levels[l - 1].system_dofs(order), mb.system_dof(lv[l - 1], order))
        G = capi.Galerkin(A, C, fd, mb.system_dof(lv[l - 1], order), ploc, fent, val, H.bdc[l] < 1.5, H.bdc[l - 1] < 1.5)
        G.apply()
        got, ref = C.to_scipy(), H.A_raw[l - 1]
        assert np.array_equal(got.indices, ref.indices)
        assert np.abs(got.data - ref.data).max() <= RTOL * np.abs(ref.data).max()
        G.apply()                                   # overwrite semantics (MAT_REUSE_MATRIX)
        assert np.abs(C.to_scipy().data - ref.data).max() <= RTOL * np.abs(ref.data).max()
        # without masks: plain P (no Dirichlet zeroing)
        Pfull = mb.prolongator(lv[l - 1], lv[l], order).tocsr()
        ref2 = (Pfull.T @ H.A_raw[l] @ Pfull).tocsr()
        G2 = capi.Galerkin(A, C, fd, mb.system_dof(lv[l - 1], order), ploc, fent, val)
        G2.apply()
        got2 = C.to_scipy()
        assert np.abs((got2 - ref2)).max() <= RTOL * np.abs(ref2.data).max()


def build_device_hierarchy(ctx, lv, order, fsrc=1.0, npre=1, npost=1, omega=0.5):
    """Mirror of LinearImplicitSystem::MGsolve on the device through the C ABI."""
    nl = len(lv)
    H = mg.Hierarchy(lv, order, fsrc)          # oracle: used for P (host-side setup) and as the checker
    top = lv[-1]
    n = mb.ndofs(top, order)
    d = mb.system_dof(top, order)
    A = [None] * nl
    A[-1] = capi.Csr.from_elements(ctx, n, d)
    asm = capi.Assembler(capi.Mesh(ctx, top.xyz, top.conn), A[-1], d, fe_hex.tables(order))
    res = ctx.vector(n)
    asm.poisson(None, res, 1.0, fsrc)
    P = [None] + [ctx.csr_from_scipy(H.P[l]) for l in range(1, nl)]
    for l in range(nl - 1, 0, -1):
        A[l - 1] = capi.Csr.from_elements(ctx, mb.ndofs(lv[l - 1], order), mb.system_dof(lv[l - 1], order))
        A[l - 1].ptap(P[l], A[l])
    M = capi.Multigrid(ctx, nl)
    for l in range(nl):
        M.set_level(l, A[l], P[l], H.bdc_idx[l], npre, npost, omega)
    M.set_coarse(1e-15, 5000)
    return H, M, A, res


@pytest.mark.parametrize("order,shape,nl", [("linear", (2, 2, 2), 3), ("biquadratic", (2, 2, 2), 3), ("biquadratic", (2, 3, 2), 2)])
def test_vcycle_residual_trace(ctx, order, shape, nl):
    lv = mb.build_hierarchy(*shape, nl)
    H, M, A, res = build_device_hierarchy(ctx, lv, order)
    # level operators after Galerkin + penalty
    for l in range(nl):
        got, ref = A[l].to_scipy(), H.A[l]
        assert np.array_equal(got.indices, ref.indices)
        assert np.abs(got.data - ref.data).max() <= RTOL * np.abs(ref.data).max()
    ncyc = 6
    trace_ref, eps_ref = H.mg_solve_trace(ncyc)
    n = res.n
    eps = ctx.vector(n)
    free = ctx.vector(H.bdc[-1])
    masked = ctx.vector(n)
    trace = []
    for _ in range(ncyc):
        M.solve(res, eps)
        masked.copy_masked(res, free, 1.1)              # UpdateRes
        trace.append(masked.norm(2))
    assert M.coarse_iterations() > 0
    for a, b in zip(trace, trace_ref):
        assert abs(a - b) <= RTOL * trace_ref[0], (trace, trace_ref)
    assert np.abs(eps.get() - eps_ref).max() <= 1e-11 * np.abs(eps_ref).max()


@pytest.mark.parametrize("order,n0", [("biquadratic", 3), ("linear", 12)])
def test_coarse_pcg_persistent_kernel_equals_the_host_driven_loop(ctx, order, n0):
    """The coarse Jacobi-PCG as one cooperative kernel (b2_cg.cu: grid-wide barriers between the phases of an iteration)
    against the host-driven loop of five launches per iteration (option coarse_persistent = 0): same solution to
    round-off, a residual at the tolerance, run-to-run bit-identical, and the V-cycle traces of both against the oracle."""
    from femus_b200.poisson import PoissonMG
    pb = PoissonMG(ctx, n0, n0, n0, 2, order, coarse_rtol=1e-14)
    pb.assemble(); pb.galerkin(); pb.mg_set_levels()
    res0 = pb.RES.get()              # (the assembly sums with atomics: the SAME assembled system goes through every mode)
    sols, its = [], []
    for mode in (1, 1, 0):
        ctx.set_option("coarse_persistent", mode)
        pb.RES.put(res0)
        pb.EPS.zero()
        pb.mg_solve()
        sols.append(pb.EPS.get())
        its.append(pb.mg.coarse_iterations())
    ctx.set_option("coarse_persistent", 1)
    assert np.array_equal(sols[0], sols[1])                                    # deterministic reductions
    assert np.abs(sols[0] - sols[2]).max() <= 1e-12 * np.abs(sols[2]).max()
    assert its[0] >= 5 and its[2] - 8 <= its[0] <= its[2]                        # the host loop looks at the residual every 8th iteration
    lv = mb.build_hierarchy(n0, n0, n0, 2)
    H = mg.Hierarchy(lv, order)
    trace_ref, eps_ref = H.mg_solve_trace(1)
    assert np.abs(sols[0] - eps_ref).max() <= 1e-11 * np.abs(eps_ref).max()


def test_single_level_is_a_direct_solve(ctx):
    """BASELINE config 1 (1 level): MGSolve degenerates to the coarse solver (reference: LU)."""
    lv = mb.build_hierarchy(4, 4, 4, 1)
    H, M, A, res = build_device_hierarchy(ctx, lv, "linear")
    n = res.n
    eps = ctx.vector(n)
    M.solve(res, eps)
    rhs = H.rhs.copy()
    rhs[H.bdc_idx[0]] = 0.0
    ref = H.lu.solve(rhs)
    assert np.abs(eps.get() - ref).max() <= 1e-12 * np.abs(ref).max()


@pytest.mark.gpu
def test_full_size_properties(ctx):
    """BASELINE config 2 (128^3 Hex27, 4 levels, 16.97 M dofs, 1.08e9 nonzeros) is too large for the
    oracle, so the full-size run is checked through size-independent properties of the path:
    the stiffness matrix annihilates constants and is symmetric, every SpMV kernel agrees, the fused
    Galerkin product equals the element-gather product, the V-cycle contracts monotonically and the
    converged solution reproduces the analytic solution of -Laplace(u) = 1 on the unit cube."""
    from femus_b200.poisson import PoissonMG
    pb = PoissonMG(ctx, 16, 16, 16, 4, "biquadratic")
    n = pb.n
    assert n == 257 ** 3 and pb.KK[-1].nnz == 1025 ** 3          # SURVEY section 8 size table
    pb.assemble()
    A = pb.KK[-1]
    rng = np.random.default_rng(5)
    one, y, x, z = ctx.vector(np.ones(n)), ctx.vector(n), ctx.vector(rng.standard_normal(n)), ctx.vector(rng.standard_normal(n))
    dmax = 0.0355555555555555 * 2.0      # ~ largest entry at this mesh size (centre-node diagonal, h = 1/128) x margin
    A.spmv(one, y)
    assert y.norm(0) <= 1e-12 * 125 * dmax                                  # A 1 = 0 (un-penalised stiffness matrix)
    A.spmv(z, y)
    xAz = x.dot(y)
    A.spmv(x, y)
    zAx = z.dot(y)
    assert abs(xAz - zAx) <= 1e-12 * x.norm(2) * z.norm(2) * 125 * dmax   # symmetry
    yref = ctx.vector(n)
    ctx.set_option("spmv_variant", 0)
    A.spmv(x, yref)
    for var in (1, 2):
        ctx.set_option("spmv_variant", var)
        A.spmv(x, y)
        y.axpy(-1.0, yref)
        assert y.norm(0) <= 1e-13 * yref.norm(0)
    ctx.set_option("spmv_variant", 1)
    # fused Galerkin product (inside the assembly) == element-gather product on the same fine matrix
    C2 = pb.KK[-2]
    xc, yc1, yc2 = ctx.vector(rng.standard_normal(C2.shape[0])), ctx.vector(C2.shape[0]), ctx.vector(C2.shape[0])
    C2.spmv(xc, yc1)
    pb.gal[-1].apply()
    C2.spmv(xc, yc2)
    yc2.axpy(-1.0, yc1)
    assert yc2.norm(0) <= 1e-12 * yc1.norm(0)
    # V-cycles: monotone contraction, then the analytic value at the cube centre
    pb.galerkin(); pb.mg_set_levels()
    trace = [pb.residual_norm()]
    for _ in range(45):
        pb.mg_solve()
        trace.append(pb.residual_norm())
    assert all(b < a for a, b in zip(trace, trace[1:])) and trace[-1] < 1e-9 * trace[0], trace[::5]
    top = pb.hier.levels[-1]
    centre = int(np.nonzero((np.abs(top.xyz - 0.5) < 1e-12).all(axis=0))[0][0])     # biquadratic dof = node id
    k = np.arange(1, 200, 2)
    sgn = np.where(((k - 1) // 2) % 2 == 0, 1.0, -1.0)
    I, J, K = np.meshgrid(k, k, k, indexing="ij")
    S = sgn[:, None, None] * sgn[None, :, None] * sgn[None, None, :]
    exact = float((64.0 / np.pi ** 5 * S / (I * J * K * (I * I + J * J + K * K))).sum())
    got = float(pb.EPS.get_indexed(np.array([centre], dtype=np.int32))[0])
    assert abs(got - exact) <= 2e-6 * exact, (got, exact)
    del pb


@pytest.mark.gpu
def test_more_than_2_to_31_nonzeros_on_one_gpu(ctx):
    """north_star's single-GPU target (256^3 Hex27) implies 64-bit entry offsets on one device: 192^3 elements, 57 M
    dofs, 1537^3 = 3.63e9 non-zeros (93 GB of device memory) through the fused assembly, the Galerkin chain, level
    setup and 45 V-cycles, checked by the size-independent properties of test_full_size_properties
    (tools/big_run.py; first run on B200: profiles/r2_big_run_192cube_3.6e9_nnz.json)."""
    import gc
    import importlib.util
    import os
    import torch
    gc.collect()
    free, _ = torch.cuda.mem_get_info(0)
    if free < 110e9:
        pytest.skip(f"needs 110 GB of free device memory, {free / 1e9:.0f} GB are free")
    spec = importlib.util.spec_from_file_location("big_run", os.path.join(os.path.dirname(os.path.dirname(__file__)), "tools", "big_run.py"))
    big = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(big)
    out = big.run(ctx, 24)
    assert out["nnz"] == 1537 ** 3 > 2 ** 31
    gc.collect()


NEU = __import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "cube_hex27_2x2x2.neu")


@pytest.mark.gpu
@pytest.mark.parametrize("order", ["linear", "biquadratic"])
def test_neu_mesh_single_level_matches_oracle(ctx, order):
    """The reference's shipped coarse mesh (cube_hex27_2x2x2.neu, unstructured element order and orientations):
    pattern bit-exact, matrix and residual to 1e-12, single-level solve against a sparse direct solve."""
    import scipy.sparse.linalg as spla
    from femus_b200 import hostapi
    from femus_b200.poisson import PoissonMG
    from oracle import gambit, mg
    H = hostapi.HostHierarchy.from_neu(NEU, 1)
    pb = PoissonMG(ctx, 0, 0, 0, 1, order, hier=H, coarse_rtol=1e-15)
    pb.assemble()
    L = gambit.read_hex27(NEU)
    Aref, rhs = mb.assemble(L, order)
    A = pb.KK[-1].to_scipy()
    assert np.array_equal(A.indptr, Aref.indptr) and np.array_equal(A.indices, Aref.indices)
    assert np.abs(A.data - Aref.data).max() <= RTOL * np.abs(Aref.data).max()
    assert np.abs(pb.RES.get() - rhs).max() <= RTOL * np.abs(rhs).max()
    pb.galerkin(); pb.mg_set_levels(); pb.mg_solve()
    idx = np.nonzero(mb.bdc_flags(L, order) < 1.5)[0]
    Ap = mg.penalty_fast(Aref, idx)
    b = rhs.copy(); b[idx] = 0.0
    x = spla.spsolve(Ap.tocsc(), b)
    assert np.abs(pb.EPS.get() - x).max() <= 1e-10 * np.abs(x).max()
    del pb


@pytest.mark.gpu
def test_neu_mesh_multilevel_solution_equals_box_solution(ctx):
    """3 levels refined from cube_hex27_2x2x2.neu by the host's topological refinement, fused assembly + element-
    matrix Galerkin chain + V-cycles to convergence: the discrete solution must coincide, node by node
    (matched through coordinates), with the oracle's solution on the generated 2x2x2 box hierarchy --
    same FE space, different element order, orientations and numbering."""
    import scipy.sparse.linalg as spla
    from femus_b200 import hostapi
    from femus_b200.poisson import PoissonMG
    from oracle import mg
    order = "biquadratic"
    H = hostapi.HostHierarchy.from_neu(NEU, 3)
    pb = PoissonMG(ctx, 0, 0, 0, 3, order, hier=H, npre=2, npost=2, coarse_rtol=1e-15)
    pb.assemble(); pb.galerkin(); pb.mg_set_levels()
    r0 = pb.residual_norm()
    for _ in range(60):
        pb.mg_solve()
    assert pb.residual_norm() <= 1e-11 * r0
    lv = mb.build_hierarchy(2, 2, 2, 3)
    O = mg.Hierarchy(lv, order)
    b = O.rhs.copy(); b[O.bdc_idx[-1]] = 0.0
    xo = spla.spsolve(O.A[-1].tocsc(), b)
    key = lambda xyz: [tuple(k) for k in np.rint(xyz.T * 64).astype(np.int64)]
    where = {k: i for i, k in enumerate(key(lv[-1].xyz))}
    perm = np.array([where[k] for k in key(H.levels[-1].xyz)])
    got = pb.EPS.get()
    assert np.abs(got - xo[perm]).max() <= 1e-9 * np.abs(xo).max()
    del pb


# ------------------------------------------------------------------------------ BASELINE sizes against the oracle
def _oracle_system(lv, order, nthreads=None):
    """Finest matrix and right-hand side of the oracle at sizes its numpy loops would take minutes for: the element
    loop by the compiled reference FE kernel (oracle/_ref) where that was built, else the numpy restatement."""
    import os
    from oracle import ref
    top = lv[-1]
    n = mb.ndofs(top, order)
    if ref.available():
        rp, ci = mb.sparsity(top, order)
        vals, rhs, _ = ref.RefHex(order).assemble_csr(top.conn, mb.system_dof(top, order), top.xyz, np.zeros(n), rp, ci, 1.0,
                                                      nthreads or os.cpu_count() or 1)
        return sp.csr_matrix((vals, ci, rp), shape=(n, n)), rhs
    return mb.assemble(top, order)


def test_baseline_config1_32cube_hex8_one_level(ctx):
    """BASELINE configs[0] at its FULL size: applications/001_Poisson on the 3-D unit box, 32^3 elements, trilinear
    (Hex8) unknown, ONE level -- 35 937 dofs.  Dof -> row map and CSR structure bit-exact, matrix entries and the
    right-hand side to 1e-12 against the oracle, MGSolve on one level (the reference: PREONLY + LU) against the solution
    of the oracle's penalised system, and the residual after it at round-off level."""
    import scipy.sparse.linalg as spla
    from femus_b200.poisson import PoissonMG
    lv = mb.build_hierarchy(32, 32, 32, 1)
    pb = PoissonMG(ctx, 32, 32, 32, 1, "linear", coarse_rtol=1e-15)
    assert pb.n == 33 ** 3
    assert np.array_equal(pb.dofs[-1], mb.system_dof(lv[-1], "linear"))                 # dof -> row map
    rp, ci = mb.sparsity(lv[-1], "linear")
    pb.assemble()
    got = pb.KK[-1].to_scipy()
    assert np.array_equal(got.indptr, rp) and np.array_equal(got.indices, ci)            # CSR structure
    Aref, rhs = _oracle_system(lv, "linear")
    Aref.sort_indices()
    assert np.abs(got.data - Aref.data).max() <= RTOL * np.abs(Aref.data).max()
    assert np.abs(pb.RES.get() - rhs).max() <= RTOL * np.abs(rhs).max()
    pb.mg_set_levels()
    pb.mg_solve()
    H = mg.Hierarchy(lv, "linear", A_top=Aref, rhs=rhs, coarse_lu=False)
    assert np.array_equal(pb.bdc_idx[-1], H.bdc_idx[-1])
    b = rhs.copy()
    b[H.bdc_idx[-1]] = 0.0
    # (a sparse LU of 36 k dofs in natural 3-D ordering takes half a minute: the oracle's solve is scipy's CG run to round-off)
    xref, info = spla.cg(H.A[-1], b, rtol=1e-15, atol=0.0, maxiter=5000, M=sp.diags(1.0 / H.A[-1].diagonal()))
    assert info == 0 and np.linalg.norm(b - H.A[-1] @ xref) <= 1e-13 * np.linalg.norm(b)
    assert np.abs(pb.EPS.get() - xref).max() <= 1e-11 * np.abs(xref).max()
    assert pb.residual_norm() <= 1e-12 * np.linalg.norm(b)


def test_baseline_config2_sample_32cube_hex27_4_levels_against_the_cpu_port(ctx):
    """BASELINE configs[1] (Hex27, 4-level V-cycle) on the sample bench.py's CPU arm runs: 32^3 elements, 274 625 dofs.
    Every level operator after the Galerkin chain + penalty and the residual norms of 4 V-cycles against the oracle
    (oracle/mg.Hierarchy for the operators, the OpenMP C port oracle/cpu_port for the cycles), 1e-12 relative."""
    import os
    from femus_b200.poisson import PoissonMG
    from oracle import cpu_port
    nt = os.cpu_count() or 1
    lv = mb.build_hierarchy(4, 4, 4, 4)
    Aref, rhs = _oracle_system(lv, "biquadratic", nt)
    H = mg.Hierarchy(lv, "biquadratic", A_top=Aref, rhs=rhs, coarse_lu=False, ptap=lambda P, Af, prp, pci: cpu_port.ptap(P, Af, prp, pci, nt))
    pb = PoissonMG(ctx, 4, 4, 4, 4, "biquadratic")
    pb.assemble(); pb.galerkin(); pb.mg_set_levels()
    for l in range(4):
        got, ref = pb.KK[l].to_scipy(), H.A[l]
        ref.sort_indices()
        assert np.array_equal(got.indptr, ref.indptr) and np.array_equal(got.indices, ref.indices)
        assert np.abs(got.data - ref.data).max() <= RTOL * np.abs(ref.data).max(), l
    M = cpu_port.PortMG(H, nt)
    res, eps = rhs.copy(), np.zeros(pb.n)
    free = H.bdc[-1] > 1.1
    for _ in range(4):
        res, eps = M.mg_solve(res, eps)
        pb.mg_solve()
        a, b = pb.residual_norm(), float(np.linalg.norm(res[free]))
        assert abs(a - b) <= 1e-11 * float(np.linalg.norm(rhs[free])), (a, b)
    assert np.abs(pb.EPS.get() - eps).max() <= 1e-10 * np.abs(eps).max()
