"""Oracle (TEST INFRASTRUCTURE ONLY): the multigrid solve the reference CONFIGURES in PETSc,
restated with scipy.sparse in fp64.

PARITY: the operators (assembled matrix, prolongators with Dirichlet rows / columns zeroed, Galerkin chain,
Dirichlet rows) and the residual norms of six V-cycles with the application's own smoother (Richardson 0.5 + SOR,
main.cpp:239-242) are pinned to REFERENCE OUTPUT (tests/golden/ref_poisson_*.npz, tests/test_reference_pin.py): the
reference's unmodified sources run here on the host backend of oracle/ref_build.  The cycle arithmetic itself lives
in PETSc 3.20.2 (contrib/scripts/install_petsc.sh:12, absent): in that run it is the host backend's C++ restatement
of PCMG, an implementation independent of this one; the two agree to the 7 digits the reference prints.  Restates, with paths relative to /root/reference/src:

  08_equations/00_stationary/LinearImplicitSystem.cpp:347-370     Galerkin chain A_{l-1} = P^T A_l P
  08_algebra.../LinearEquationSolverPetsc.cpp:53-90, 428-436      BuildBdcIndex, SetPenalty
                                                                   (MatZeroRows, diag 1, pattern kept)
  .../LinearEquationSolverPetsc.cpp:185-290                       PCMG multiplicative V, Richardson+Jacobi
  .../LinearEquationSolverPetsc.cpp:294-353, 417-424              MGSolve: RES[bdc]=0, cycle, RESC=A EPSC,
                                                                   RES-=RESC, EPS+=EPSC
  06_solution/00_single_level/00_definition/Solution.cpp:595-628  UpdateRes (zeros where Bdc<=1.1)
  08_equations/00_stationary/LinearImplicitSystem.cpp:415-449     HasLinearConverged: ||Res||_2

The outer Krylov solver is PREONLY (as applications/MGAMR/ex5/ex5.cpp:141-204 sets it), so one
MGSolve is exactly one V-cycle; the coarse solver is a sparse LU (the reference: MUMPS).
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from . import mesh_box as mb


def penalty(A, bdc_idx):
    """MatZeroRows(A, idx, 1.0) keeping the nonzero pattern: rows zeroed, diagonal 1."""
    A = A.tocsr(copy=True)
    for r in bdc_idx:
        A.data[A.indptr[r]:A.indptr[r + 1]] = 0.0
    A = A.tolil()
    for r in bdc_idx:
        A[r, r] = 1.0
    A = A.tocsr()
    A.sort_indices()
    return A


def penalty_fast(A, bdc_idx):
    A = A.tocsr(copy=True)
    A.sort_indices()
    mask = np.zeros(A.shape[0], dtype=bool)
    mask[bdc_idx] = True
    rows = np.repeat(np.arange(A.shape[0]), np.diff(A.indptr))
    isd = mask[rows]
    A.data[isd] = 0.0
    A.data[isd & (A.indices == rows)] = 1.0
    return A


def on_pattern(A, rowptr, col):
    """Copy A's entries onto the (larger) CSR pattern (rowptr, col); missing entries become
    explicit zeros.  Every entry of A must lie in the pattern."""
    A = A.tocoo()
    n = A.shape[1]
    rows = np.repeat(np.arange(len(rowptr) - 1, dtype=np.int64), np.diff(rowptr))
    keys = rows * n + col.astype(np.int64)
    ka = A.row.astype(np.int64) * n + A.col.astype(np.int64)
    pos = np.searchsorted(keys, ka)
    assert np.array_equal(keys[pos], ka), "entry outside the pattern"
    vals = np.zeros(col.shape[0])
    np.add.at(vals, pos, A.data)
    return sp.csr_matrix((vals, col.copy(), rowptr.copy()), shape=A.shape)


class Hierarchy:
    """Per-level operators exactly as the reference sets them up for one MGsolve."""

    def __init__(self, levels, order, fsrc=1.0, dirichlet_faces=(1, 2, 3, 4, 5, 6), A_top=None, rhs=None,
                 coarse_lu=True, ptap=None, neumann=None, smoother="richardson", mesh=None, asm_blocks=None, asm_sub="lu",
                 asm_orders=None, ksp="richardson", nullspace=None):
        """mesh: the oracle module the levels come from (mesh_box by default, mesh_tet for tetrahedra).
        smoother "asm": asm_blocks[l] = the overlapping index sets of level l >= 1 (oracle.asm.level_blocks),
        asm_orders[l] = the order they are swept in (None: as listed)."""
        self.asm_blocks, self.asm_sub, self.asm_orders = asm_blocks, asm_sub, asm_orders
        # nullspace[l]: vector spanning the null space of the level-l operator, l >= 1 (RemoveNullSpace,
        # LinearEquationSolverPetsc.cpp:357-414), or None
        self.nullspace = nullspace
        self.ksp = ksp          # "gmres": KSPGMRES (left preconditioning) around the Jacobi / element-block preconditioner
        mb = mesh if mesh is not None else globals()["mb"]
        self.levels = levels
        self.order = order
        self.smoother = smoother
        nl = len(levels)
        self.bdc = [mb.bdc_flags(L, order, dirichlet_faces) for L in levels]
        self.bdc_idx = [np.nonzero(b < 1.5)[0] for b in self.bdc]
        # prolongators with Dirichlet rows/cols zeroed (built at init(), before assembly)
        self.P = [None] * nl
        for l in range(1, nl):
            P = mb.prolongator(levels[l - 1], levels[l], order)
            self.P[l] = mb.zero_dirichlet(P, self.bdc[l], self.bdc[l - 1])
        # coarse patterns = coarse element coupling patterns (explicit zeros kept); fixed at init()
        self.patterns = [mb.sparsity(levels[l], order) for l in range(nl - 1)]
        self.coarse_lu = coarse_lu
        # assembly on the finest level (V_CYCLE: only the top level is assembled)
        if A_top is None:
            A_top, rhs = mb.assemble(levels[-1], order, None, fsrc)
            if neumann:
                rhs = rhs + mb.neumann_rhs(levels[-1], order, neumann)
        self.set_operator(A_top, rhs, ptap)

    def set_operator(self, A_top, rhs, ptap=None):
        """What one MGsolve does with a freshly assembled finest matrix: Galerkin chain on the
        un-penalised matrices, then MGSetLevel's penalty and the Jacobi diagonal on every level.
        ptap: optional threaded triple product (oracle.cpu_port.ptap) for the CPU baseline."""
        nl = len(self.levels)
        self.A_raw = [None] * nl
        self.A_raw[-1], self.rhs = A_top, rhs
        for l in range(nl - 1, 0, -1):
            rp, ci = self.patterns[l - 1]
            if ptap is not None:
                self.A_raw[l - 1] = ptap(self.P[l], self.A_raw[l], rp, ci)
                continue
            Ac = (self.P[l].T @ self.A_raw[l] @ self.P[l]).tocsr()
            self.A_raw[l - 1] = on_pattern(Ac, rp, ci)
        self.A = [penalty_fast(self.A_raw[l], self.bdc_idx[l]) for l in range(nl)]
        with np.errstate(divide="ignore"):          # velocity-pressure systems: zero pressure diagonal (Jacobi is not used there)
            self.dinv = [1.0 / A.diagonal() for A in self.A]
        self.lu = spla.splu(self.A[0].tocsc()) if self.coarse_lu else None
        self.asm = [None] * nl
        if self.smoother == "asm":          # PCASM sub-matrices are extracted from the penalised level operator
            from . import asm as _asm
            for l in range(1, nl):
                self.asm[l] = _asm.BlockSmoother(self.A[l], self.asm_blocks[l], self.asm_sub,
                                                 self.asm_orders[l] if self.asm_orders else None)
        # Chebyshev bounds: our own stated ones, [0.1, 1.1] x the power-iteration estimate of lambda_max(D^-1 A)
        self.ebounds = [None] * nl
        if self.smoother == "chebyshev":
            for l in range(1, nl):
                lam = self.estimate_emax(l)
                self.ebounds[l] = (0.1 * lam, 1.1 * lam)

    def estimate_emax(self, l, its=10):
        """Largest eigenvalue of D^-1 A by power iteration from the fixed start vector
        v_i = 1 + 0.5 sin(i mod 1000), zero on Dirichlet rows (b2_mg.cu estimate_emax)."""
        n = self.A[l].shape[0]
        v = 1.0 + 0.5 * np.sin((np.arange(n) % 1000).astype(np.float64))
        v[self.bdc_idx[l]] = 0.0
        lam = 0.0
        for it in range(its + 1):
            nrm = float(np.sqrt(v @ v))
            if it > 0:
                lam = nrm
            if it == its or nrm == 0.0:
                break
            v = self.dinv[l] * (self.A[l] @ (v / nrm))
        return lam

    def pc_apply(self, l, r):
        return self.asm[l].apply(r) if self.smoother == "asm" else self.dinv[l] * r

    def gmres(self, l, x, b, k):
        """k iterations of left-preconditioned GMRES from x (PETSc's KSPGMRES defaults; no restart within the call):
        minimises ||M^-1 (b - A x)|| over x + K_k(M^-1 A, M^-1 r0).  Modified Gram-Schmidt + Givens rotations."""
        A = self.A[l]
        z = self.pc_apply(l, b - A @ x)
        beta = float(np.sqrt(z @ z))
        if k <= 0 or not beta > 0.0:
            return x
        V = [z / beta]
        H = np.zeros((k + 1, k))
        cs, sn, g = np.zeros(k), np.zeros(k), np.zeros(k + 1)
        g[0] = beta
        m = 0
        for j in range(k):
            w = self.pc_apply(l, A @ V[j])
            for i in range(j + 1):
                H[i, j] = w @ V[i]
                w = w - H[i, j] * V[i]
            hn = float(np.sqrt(w @ w))
            H[j + 1, j] = hn
            for i in range(j):
                a, c = H[i, j], H[i + 1, j]
                H[i, j], H[i + 1, j] = cs[i] * a + sn[i] * c, -sn[i] * a + cs[i] * c
            a, c = H[j, j], H[j + 1, j]
            d = np.sqrt(a * a + c * c)
            m = j + 1
            if not d > 0.0:
                m = j
                break
            cs[j], sn[j] = a / d, c / d
            H[j, j], H[j + 1, j] = d, 0.0
            g[j + 1] = -sn[j] * g[j]
            g[j] = cs[j] * g[j]
            if not hn > 0.0:
                break
            V.append(w / hn)
        y = np.zeros(m)
        for i in range(m - 1, -1, -1):
            y[i] = (g[i] - H[i, i + 1:m] @ y[i + 1:]) / H[i, i]
        for i in range(m):
            x = x + y[i] * V[i]
        return x

    def smooth(self, l, x, b, nsweeps, omega):
        """KSPRICHARDSON (scale omega) + PCJACOBI: x <- x + omega D^-1 (b - A x); or Chebyshev + Jacobi
        on the stated interval (Saad, alg. 12.1), restarted at every call like a PETSc smoother."""
        if self.nullspace is not None and self.nullspace[l] is not None:
            # KSPSolve with MatSetNullSpace + MatSetTransposeNullSpace: the right-hand side is projected (a copy), and so
            # is every preconditioned residual (KSP_PCApply -> KSP_RemoveNullSpace); Richardson around the preconditioner
            assert self.ksp == "richardson" and self.smoother != "chebyshev"
            nv = self.nullspace[l] / np.linalg.norm(self.nullspace[l])
            bp = b - (nv @ b) * nv
            for _ in range(nsweeps):
                z = self.pc_apply(l, bp - self.A[l] @ x)
                x = x + omega * (z - (nv @ z) * nv)
            return x
        if self.ksp == "gmres" and self.smoother != "chebyshev":
            return self.gmres(l, x, b, nsweeps)
        if self.smoother == "asm":           # KSPRICHARDSON (scale omega) + PCASM (basic, multiplicative)
            return self.asm[l].richardson(x, b, nsweeps, omega)
        if self.smoother == "chebyshev":
            emin, emax = self.ebounds[l]
            theta, delta = 0.5 * (emax + emin), 0.5 * (emax - emin)
            sigma1 = theta / delta
            rho = 1.0 / sigma1
            d = None
            for k in range(nsweeps):
                z = self.dinv[l] * (b - self.A[l] @ x)
                if k == 0:
                    d = z / theta
                else:
                    rho_new = 1.0 / (2.0 * sigma1 - rho)
                    d = (rho_new * rho) * d + (2.0 * rho_new / delta) * z
                    rho = rho_new
                x = x + d
            return x
        for _ in range(nsweeps):
            x = x + omega * (self.dinv[l] * (b - self.A[l] @ x))
        return x

    def vcycle(self, l, b, npre=1, npost=1, omega=0.5):
        """PCMG multiplicative V-cycle, zero initial guess on every level."""
        if l == 0:
            return self.lu.solve(b)
        x = np.zeros_like(b)
        x = self.smooth(l, x, b, npre, omega)
        r = b - self.A[l] @ x
        bc = self.P[l].T @ r
        xc = self.vcycle(l - 1, bc, npre, npost, omega)
        x = x + self.P[l] @ xc
        x = self.smooth(l, x, b, npost, omega)
        return x

    def mg_solve_trace(self, ncycles, npre=1, npost=1, omega=0.5):
        """Vcycle() of LinearImplicitSystem.cpp:468-497 with outer PREONLY: returns the list of
        ||_Res||_2 after each MGSolve+UpdateRes, and the final EPS."""
        top = len(self.levels) - 1
        res = self.rhs.copy()
        eps = np.zeros_like(res)
        free = self.bdc[top] > 1.1
        trace = []
        for _ in range(ncycles):
            res[self.bdc_idx[top]] = 0.0                      # ZerosBoundaryResiduals
            epsc = self.vcycle(top, res, npre, npost, omega)
            resc = self.A[top] @ epsc
            res = res - resc
            eps = eps + epsc
            trace.append(float(np.linalg.norm(np.where(free, res, 0.0))))
        return trace, eps
