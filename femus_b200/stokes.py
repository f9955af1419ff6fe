"""Harness-side mirror of the reference's driver for the steady Stokes application
(applications/003_NavierStokes/SteadyStokes/main.cpp: system "Navier-Stokes" with U, V(, W), P; assembly callback
AssembleMatrixResNS; LinearImplicitSystem::MGsolve with the ASM / Vanka level solver), in three dimensions, expressed
as calls into the C++ host layer and the device C ABI -- SURVEY 8f row 3, first vertical slice:

    system.init()       per-level system matrices on the multi-variable pattern, system prolongators with the
                        Dirichlet rows / columns zeroed                       hostapi.SystemOnLevel
    assembly            b2_stokes_assemble on the finest level                 capi.StokesAssembler
    Galerkin chain      the general triple product (MatPtAP)                   Csr.ptap
    level solver        Richardson around velocity-pressure Vanka blocks       hostapi.AsmIndex(nschur=1), capi.Schwarz
    coarsest level      a direct solve (PREONLY + LU in the reference)         Multigrid.set_coarse_schwarz

No arithmetic happens here: every number is produced by libfemus_b200.so.  One rank; meshes of several element types
get one assembly plan per type."""
import numpy as np

from . import capi, hostapi


class StokesMG:
    def __init__(self, ctx, hier, order_v="biquadratic", order_p="linear", IRe=1.0, velocity_dirichlet=(1, 2, 3, 4, 5, 6),
                 pressure_dirichlet=(), npre=1, npost=1, omega=1.0, block_elems=1, schedule="colours", equation="stokes",
                 block_sub="lu", boundary_pressure=None, fix_pressure_at_one_point=False):
        """equation: "stokes" (SteadyStokes/main.cpp, IRe = the viscosity factor) or "navier_stokes" (the library routine
        03_navier_stokes.hpp: Galerkin residual + exact Newton Jacobian, IRe = nu).  boundary_pressure: {boundary set:
        prescribed pressure tau} for the faces whose normal velocity is not Dirichlet (the routine's boundary block,
        :196-300; Navier-Stokes only).  fix_pressure_at_one_point: MultiLevelSolution::FixSolutionAtOnePoint("P") for
        enclosed flows (every velocity Dirichlet: the pressure is defined up to a constant) -- the first pressure dof of the
        COARSEST level becomes a Dirichlet row (MultiLevelSolution.cpp:826-830), and on the levels above the constant
        pressure is removed as the null space of the operator (RemoveNullSpace, LinearEquationSolverPetsc.cpp:357-414)."""
        self.ctx, self.hier, self.IRe, self.equation = ctx, hier, IRe, equation
        self.fams = [order_v] * 3 + [order_p]
        lv = hier.levels
        nl = self.nlevels = len(lv)
        self.npre, self.npost, self.omega = npre, npost, omega
        top = lv[-1]
        self.sys = [hostapi.SystemOnLevel(L, self.fams) for L in lv]
        self.n = self.sys[-1].n
        dirichlet = [velocity_dirichlet] * 3 + [pressure_dirichlet]
        self.bdc = [S.bdc(dirichlet) for S in self.sys]
        self.fix_pressure = fix_pressure_at_one_point
        if fix_pressure_at_one_point:
            self.bdc[0][int(self.sys[0].offsets[3, 0])] = 0.0
        self.bdc_idx = [np.nonzero(b < 1.5)[0].astype(np.int32) for b in self.bdc]
        self.pattern = [S.sparsity() for S in self.sys]
        self.KK = [ctx.csr(S.n, S.n, *pat) for S, pat in zip(self.sys, self.pattern)]
        self.PP = [None] * nl
        for l in range(1, nl):
            rp, ci, v, shape = self.sys[l].prolongator()
            P = ctx.csr(shape[0], shape[1], rp, ci, v)
            P.zero_rows(self.bdc_idx[l], 0.0)
            P.zero_cols(self.bdc_idx[l - 1])
            self.PP[l] = P
        # one plan per element type present (the tables are the element type), all accumulating into KK and RES
        edofs = self.sys[-1].elem_dofs()
        self.plans = []
        self.pressure_groups = []
        for t in ([top.elem_type] if top.elem_type >= 0 else sorted(set(top.elem_types.tolist()))):
            sel = slice(None) if top.elem_type >= 0 else np.nonzero(top.elem_types == t)[0]
            mesh_t = capi.Mesh(ctx, top.xyz, np.ascontiguousarray(top.conn[sel]))
            self.plans.append((mesh_t, capi.StokesAssembler(mesh_t, self.KK[-1], np.ascontiguousarray(edofs[sel]), hostapi.elem_tables(t, order_v),
                                                            hostapi.elem_tables(t, order_p), navier_stokes=(equation == "navier_stokes"))))
            if boundary_pressure:
                from .poisson import neumann_face_groups
                for faces, tabs, fnodes in neumann_face_groups(top, order_v, dict(boundary_pressure), t, sel):
                    self.pressure_groups.append((self.plans[-1][1], faces, tabs, fnodes))
        self.mesh, self.asm = self.plans[0]
        self.RES, self.EPS, self.SOL = ctx.vector(self.n), ctx.vector(self.n), ctx.vector(self.n)
        self.BDC, self.RESM = ctx.vector(self.bdc[-1]), ctx.vector(self.n)
        self.mg = capi.Multigrid(ctx, nl)
        # coarsest level: one block with every dof, exact solve
        n0 = self.sys[0].n
        self.coarse = capi.Schwarz(ctx, self.KK[0], np.array([0, n0], dtype=np.int64), np.arange(n0, dtype=np.int32),
                                   np.array([0, 1], dtype=np.int64), np.zeros(1, dtype=np.int32))
        self.mg.set_coarse_schwarz(self.coarse)
        # levels above: Vanka blocks, the pressure being the Schur variable (velocities of the near elements)
        self.asm_index, self.asm_groups, self.schwarz = [None] * nl, [None] * nl, [None] * nl
        for l in range(1, nl):
            ix = hostapi.AsmIndex(lv[l], self.fams, block_elems, nschur=1)
            grp, gptr, gblocks = hostapi.asm_schedule(*self.pattern[l], ix.overlap_ptr, ix.overlap, schedule)
            self.asm_index[l], self.asm_groups[l] = ix, grp
            self.schwarz[l] = capi.Schwarz(ctx, self.KK[l], ix.overlap_ptr, ix.overlap, gptr, gblocks)
            self.schwarz[l].set_subsolver(block_sub)     # "lu": exact (MLU_PRECOND), "ilu": ILU(0) in system-dof order (ILU_PRECOND)
            self.mg.set_level_schwarz(l, self.schwarz[l])
            if fix_pressure_at_one_point:                 # GetNullSpaceBase: 1 on the pressure dofs with Bdc > 1.9
                self.mg.set_level_nullspace(l, ctx.vector(self.nullspace_base(l)))

    def nullspace_base(self, l):
        nv = np.zeros(self.sys[l].n)
        p0, p1 = int(self.sys[l].offsets[3, 0]), int(self.sys[l].offsets[4, 0])
        nv[p0:p1] = self.bdc[l][p0:p1] > 1.9
        return nv

    def assemble(self):
        self.RES.zero()
        self.KK[-1].zero()
        for _, plan in self.plans:
            if self.equation == "navier_stokes":
                plan.assemble_ns(self.SOL, self.RES, self.IRe)
            else:
                plan.assemble(self.SOL, self.RES, self.IRe)
        if self.equation == "navier_stokes":
            for plan, faces, tabs, fnodes in self.pressure_groups:
                plan.pressure_faces(*faces, tabs, fnodes, self.RES)

    def galerkin(self):
        for l in range(self.nlevels - 1, 0, -1):
            self.KK[l - 1].ptap(self.PP[l], self.KK[l])

    def mg_set_levels(self):
        for l in range(self.nlevels):
            self.mg.set_level(l, self.KK[l], self.PP[l], self.bdc_idx[l], self.npre, self.npost, self.omega)

    def mg_solve(self):
        self.mg.solve(self.RES, self.EPS)

    def newton_step(self, ncycles=1):
        """One iteration of NonLinearImplicitSystem::solve on the finest level (NonLinearImplicitSystem.cpp:157-361,
        reduced to its V-cycle path): assemble residual and Jacobian at Sol, Galerkin chain, level setup, `ncycles`
        MGSolve, Sol += EPS.  Returns ||RES||_2 over the free rows BEFORE the update."""
        self.EPS.zero()
        self.assemble()
        r0 = self.residual_norm()
        self.galerkin()
        self.mg_set_levels()
        for _ in range(ncycles):
            self.mg_solve()
        self.SOL.axpy(1.0, self.EPS)
        return r0

    def residual_norm(self):
        self.RESM.copy_masked(self.RES, self.BDC, 1.1)
        return self.RESM.norm(2)
