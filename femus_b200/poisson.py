"""Harness-side mirror of the reference's driver for the in-scope application: the sequence
applications/001_Poisson/main.cpp performs through MultiLevelMesh / MultiLevelSolution /
LinearImplicitSystem, expressed as calls into the C++ host layer and the device C ABI.

    GenerateCoarseBoxMesh + RefineMesh          main.cpp:133-141
    AddSolution / GenerateBdc("All")            main.cpp:149-184
    system.init()                               LinearImplicitSystem.cpp:138-282
        per-level _KK with exact sparsity       LinearEquation.cpp:196-342
        BuildProlongatorMatrix, ZeroInterpolatorDirichletNodes   :826-909, :1032-1120
    system.MGsolve()                            LinearImplicitSystem.cpp:288-411
        SetResZero, assemble, PtAP chain, MGInit, MGSetLevel, Vcycle

No arithmetic happens here: every number is produced by libfemus_b200.so."""
import numpy as np

from . import capi, hostapi
from . import dist as distlayout


def neumann_face_groups(level, order, neumann, elem_type, sel, bfaces=None):
    """Neumann faces of the elements of one assembly plan (element type `elem_type`, rows `sel` of the level's
    connectivity), one group per face kind: the reference picks the face element per face,
    _finiteElement[GetElementFaceType(iel, jface)][order_ind] (main.cpp:507-525).  Returns a list of
    ((plan-local element, local face, flux), face tables, face nodes[6][9]) ready for b2_asm_neumann_faces."""
    fe, fl, fb = bfaces if bfaces is not None else level.boundary_faces()
    nm = np.isin(fb, list(neumann))
    fe, fl, fb = fe[nm], fl[nm], fb[nm]
    pos = np.full(level.nel, -1, dtype=np.int64)         # element -> row of the plan's connectivity
    pos[sel] = np.arange(level.nel if isinstance(sel, slice) else len(sel))
    kinds = hostapi.elem_face_kinds(elem_type)
    mine = pos[fe] >= 0
    out = []
    for kind in (hostapi.QUAD_FACE, hostapi.TRI_FACE):
        g = mine & (kinds[fl] == kind)
        if g.any():
            faces = (pos[fe[g]].astype(np.int32), fl[g].astype(np.int32),
                     np.array([neumann[int(b)] for b in fb[g]], dtype=np.float64))
            out.append((faces, hostapi.face_kind_tables(kind, order), hostapi.elem_face_nodes(elem_type)))
    return out


class PoissonMG:
    """dist=None: the whole mesh on one GPU.  dist=(rank, world, allgather): this rank's z-slab of
    the mesh (local hierarchy, partial matrices, interface sums through b2_halo); the context must
    already hold the NCCL communicator (Context.comm_init)."""

    def __init__(self, ctx, nx, ny, nz, nlevels, order="biquadratic", bounds=None, npre=1, npost=1, omega=0.5,
                 dirichlet_faces=(1, 2, 3, 4, 5, 6), fsrc=1.0, coarse_rtol=1e-14, hier=None, dist=None, fused=True,
                 neumann=None, smoother="richardson", asm_block_elems=8, asm_schedule="colours",
                 asm_sub="lu", ksp="richardson", asm_row_levels=False, peer=True):
        self.ctx = ctx
        self.order = order
        self.fam = hostapi.FAMILY[order]
        self.nlevels = nlevels
        self.npre, self.npost, self.omega, self.fsrc = npre, npost, omega, fsrc
        self.dist = dist
        # neumann: {boundary index: constant flux} (JSON "bdc_type": "neumann", "bdc_func": value)
        self.neumann = dict(neumann) if neumann else {}
        # fused: the finest Galerkin product is formed inside the assembly kernel (b2_asm_poisson_galerkin)
        self.fused = fused and nlevels > 1
        if hier is not None:
            self.hier = hier
        elif dist is None:
            self.hier = hostapi.HostHierarchy(nx, ny, nz, nlevels, bounds)
        else:
            self.hier = hostapi.HostHierarchy(nx, ny, nz, nlevels, bounds, nprocs=dist[1], local_rank=dist[0])
        lv = self.hier.levels
        top = lv[-1]
        self.ndofs = [L.ndofs(order) for L in lv]
        self.n = self.ndofs[-1]
        self.nel = top.nel
        # element type of the mesh (hexahedra from the box generator or a .neu file, tetrahedra from a .neu
        # file): the reference dispatches on it through _finiteElement[ielGeom][solType] (main.cpp:438)
        self.elem_type = top.elem_type             # -1: the mesh mixes hexahedra, tetrahedra and wedges
        self.mixed = self.elem_type < 0
        # fast Galerkin paths: hexahedra with 8 or 27 dofs (the 20-node family uses the general triple product)
        self.hex = self.elem_type == hostapi.HEX and order != "quadratic"
        self.nve = None if self.mixed else hostapi.elem_nve(self.elem_type, order)
        if not self.hex:       # the element-gather / fused Galerkin products are kernels for refined hexahedra
            self.fused = False
            if dist is not None:
                raise NotImplementedError("the sharded run is implemented for hexahedra with 8 or 27 dofs")
        # --- system.init(): per-level matrices with the exact element-coupling pattern
        # (a mesh of several element types has ragged element rows: pattern built on the host, as the
        # reference's GetSparsityPatternSize does, LinearEquation.cpp:407-548)
        if self.mixed:
            self.dofs = [L.system_dofs27(order) for L in lv]
            self.KK = [ctx.csr(self.ndofs[l], self.ndofs[l], *lv[l].sparsity(order)) for l in range(nlevels)]
        else:
            self.dofs = [L.system_dofs(order) for L in lv]
            self.KK = [capi.Csr.from_elements(ctx, self.ndofs[l], self.dofs[l]) for l in range(nlevels)]
        self.bdc = [L.bdc(order, dirichlet_faces) for L in lv]
        self.bdc_idx = [np.nonzero(b < 1.5)[0].astype(np.int32) for b in self.bdc]
        # --- prolongators, Dirichlet rows (fine) and columns (coarse) zeroed
        self.PP = [None] * nlevels
        for l in range(1, nlevels):
            rp, ci, v, shp = self.hier.prolongator(l, order)
            P = ctx.csr(shp[0], shp[1], rp, ci, v)
            P.zero_rows(self.bdc_idx[l], 0.0)
            P.zero_cols(self.bdc_idx[l - 1])
            self.PP[l] = P
        # --- Galerkin plans: element-gather P^T A P per level pair (fast path of matrix_PtAP)
        # (hexahedra; other element types use the general triple product b2_csr_ptap, like MatPtAP)
        self.gal = [None] * nlevels
        if self.hex:
            ploc, fent = hostapi.galerkin_element(order)
        for l in range(1, nlevels if self.hex else 0):
            fd, val = self.hier.galerkin_maps(l - 1, order)
            self.gal[l] = capi.Galerkin(self.KK[l], self.KK[l - 1], fd, self.dofs[l - 1], ploc, fent, val,
                                        self.bdc[l] < 1.5, self.bdc[l - 1] < 1.5)
        # --- finest-level mesh + assembly plan
        # one plan per element type present (the tables are the element type: _finiteElement[ielGeom][solType])
        self.plans = []
        self.nm_groups = []         # Neumann faces per (plan, face kind): (assembler, faces, tables, face nodes)
        if self.neumann:
            fe, fl, fb = top.boundary_faces()
            nm = np.isin(fb, list(self.neumann))
            fe, fl, fb = fe[nm], fl[nm], fb[nm]
        for t in ([self.elem_type] if not self.mixed else sorted(set(top.elem_types.tolist()))):
            sel = slice(None) if not self.mixed else np.nonzero(top.elem_types == t)[0]
            mesh_t = capi.Mesh(ctx, top.xyz, np.ascontiguousarray(top.conn[sel]))
            dof_t = self.dofs[-1] if not self.mixed else np.ascontiguousarray(self.dofs[-1][sel][:, :hostapi.elem_nve(t, order)])
            tables_t = hostapi.elem_tables(t, order)
            self.plans.append((mesh_t, capi.Assembler(mesh_t, self.KK[-1], dof_t, tables_t), tables_t))
            if self.neumann and not self.hex:
                for faces, tabs, fnodes in neumann_face_groups(top, order, self.neumann, t, sel, (fe, fl, fb)):
                    self.nm_groups.append((self.plans[-1][1], faces, tabs, fnodes))
        self.mesh, self.asm, self.tables = self.plans[0]
        if self.neumann and self.hex:    # Neumann faces of the finest level: (element, local face, flux)
            self.nm_faces = (fe, fl, np.array([self.neumann[int(b)] for b in fb], dtype=np.float64))
            self.nm_tables = hostapi.face_tables(order)
            self.nm_face_nodes = hostapi.hex_face_nodes()
        if self.fused:      # element-matrix Galerkin chain: every plan but the coarsest records its element matrices
            for l in range(2, nlevels):
                self.gal[l].record_elements(True)
        # --- vectors of the finest LinearEquation: _RES, _EPS; solution Sol and its Bdc mask
        self.RES = ctx.vector(self.n)
        self.EPS = ctx.vector(self.n)
        self.SOL = ctx.vector(self.n)
        self.BDC = ctx.vector(self.bdc[-1])
        self.RESM = ctx.vector(self.n)
        self.mg = capi.Multigrid(ctx, nlevels)
        self.mg.set_coarse(coarse_rtol, 10000)
        self.smoother = smoother
        self.asm_index = [None] * nlevels
        self.asm_groups = [None] * nlevels
        self.schwarz = [None] * nlevels
        if smoother == "asm":
            # "smoother": "asm" of 001_Poisson (main.cpp:234-250): element blocks of every level above the coarsest
            # (DoPartition + BuildASMIndex on the host), swept in `asm_schedule` order: "levels" = the reference's
            # block order exactly, "colours" = the same sweep with the blocks stably sorted by colour
            if dist is not None:
                raise NotImplementedError("the element-block smoother runs on one rank")
            # asm_block_elems: elements per block (LinearEquationSolverPetscAsm::SetElementBlockNumber(n); the system-level
            # SetElementBlockNumber(d) of LinearImplicitSystem.cpp:1191-1201 passes 8^d, capped by the level's element
            # count); "all": one block with every element of the level, which on one rank is what
            # SetElementBlockNumber("All", overlap) and FEMuS_DEFAULT amount to (the sub-solver preconditions the level)
            nb_elems = 2 ** 30 if asm_block_elems == "all" else int(asm_block_elems)
            for l in range(1, nlevels):
                ix = hostapi.AsmIndex(lv[l], order, nb_elems)
                rp, ci = lv[l].sparsity(order)
                grp, gptr, gblocks = hostapi.asm_schedule(rp, ci, ix.overlap_ptr, ix.overlap, asm_schedule)
                self.asm_index[l], self.asm_groups[l] = ix, grp
                self.schwarz[l] = capi.Schwarz(ctx, self.KK[l], ix.overlap_ptr, ix.overlap, gptr, gblocks)
                self.schwarz[l].set_subsolver(asm_sub)      # "lu": MLU_PRECOND on the blocks, "ssor": SOR_PRECOND (main.cpp:242)
                if asm_row_levels:                          # large blocks: rows of a dependency level in parallel
                    self.schwarz[l].set_row_levels(True)
                self.mg.set_level_schwarz(l, self.schwarz[l])
        elif smoother != "richardson":
            for l in range(1, nlevels):
                self.mg.set_smoother(l, smoother)
        # level solver around the Jacobi / element-block preconditioner: SetSolverFineGrids(RICHARDSON | GMRES)
        self.ksp = ksp
        if ksp != "richardson":
            if dist is not None or smoother == "chebyshev":
                raise NotImplementedError("GMRES as level solver: one rank, Jacobi or element-block preconditioner")
            for l in range(1, nlevels):
                self.mg.set_level_ksp(l, ksp)
        # --- distributed layout: interface dofs of every level, ownership; reductions over owned dofs
        self.layout = [None] * nlevels
        self.halo = [None] * nlevels
        self.n_global = self.n
        if dist is not None:
            rank, world, gather = dist
            for l in range(nlevels):
                self.layout[l] = distlayout.level_layout(lv[l], self.ndofs[l], rank, gather)
            # interface sums through peer memory (NVLink / NVSwitch stores + flags) instead of a packed ncclAllReduce:
            # one inbox block per rank, sized for the largest message of the run, opened by every rank (once per context)
            if peer and getattr(ctx, "peer_slot", None) is None:
                need = max(lay.exchange[4] for lay in self.layout)
                ctx.peer_slot = max(max(gather(need)) + 8, 1 << 16)
                ctx.peer_init(ctx.peer_slot, gather)
            for l in range(nlevels):
                lay = self.layout[l]
                self.halo[l] = capi.Halo(ctx, lay.n_local, lay.idx, lay.pos, lay.n_packed, lay.owned, lay.mult)
                if peer:
                    if lay.exchange[4] + 8 > ctx.peer_slot:
                        raise ValueError("the peer inbox of this context is too small for this mesh")
                    self.halo[l].set_exchange(*lay.exchange[:4])
                self.mg.set_level_halo(l, self.halo[l])
            self.n_global = int(sum(gather(self.layout[-1].n_owned)))
            for v in (self.RES, self.EPS, self.SOL, self.RESM):
                v.set_halo(self.halo[-1])

    # ---- pieces of MGsolve -----------------------------------------------------------------
    def assemble(self):
        """SetResZero + the assembly callback (KK->zero(); element loop; close())."""
        self.RES.zero()
        self.KK[-1].zero()
        if self.fused:
            self.asm.poisson_galerkin(self.gal[-1], self.SOL, self.RES, 1.0, self.fsrc)
        else:
            for _, asm, _ in self.plans:       # one launch per element type, accumulating into KK and RES
                asm.poisson(self.SOL, self.RES, 1.0, self.fsrc)
        if self.neumann and self.hex and self.nm_faces[0].size:
            self.asm.neumann(*self.nm_faces, self.nm_tables, self.nm_face_nodes, self.RES)
        for asm, faces, tabs, fnodes in self.nm_groups:
            asm.neumann_faces(*faces, tabs, fnodes, self.RES)
        if self.halo[-1] is not None:          # close(): contributions of the other ranks' elements
            self.halo[-1].sum(self.RES)

    def galerkin(self, algebraic=False):
        """A_{l-1} = P_l^T A_l P_l down the hierarchy, on the un-penalised matrices: element-gather
        plans by default, the general sparse triple product (b2_csr_ptap) on request."""
        top = self.nlevels - 1
        algebraic = algebraic or not self.hex
        for l in range(top, 0, -1):
            if l == top and self.fused and not algebraic:
                continue            # already formed by the fused assembly
            if algebraic:
                self.KK[l - 1].ptap(self.PP[l], self.KK[l])
            elif self.fused:
                self.gal[l].apply_from_elements(self.gal[l + 1])
            else:
                self.gal[l].apply()

    def mg_set_levels(self):
        """MGInit + MGSetLevel on every level (SetPenalty, smoother setup)."""
        for l in range(self.nlevels):
            self.mg.set_level(l, self.KK[l], self.PP[l], self.bdc_idx[l], self.npre, self.npost, self.omega)

    def mg_solve(self):
        """One MGSolve (outer PREONLY => one V-cycle) + UpdateRes; returns nothing (no sync)."""
        self.mg.solve(self.RES, self.EPS)

    def residual_norm(self):
        """||_Res||_2 after UpdateRes (zeros where Bdc <= 1.1); synchronises."""
        self.RESM.copy_masked(self.RES, self.BDC, 1.1)
        return self.RESM.norm(2)

    def step(self):
        """One pass of the hot path: assembly, Galerkin chain, level setup, one V-cycle."""
        self.EPS.zero()
        self.assemble()
        self.galerkin()
        self.mg_set_levels()
        self.mg_solve()

    def update_sol(self):
        """UpdateSol: Sol += EPS."""
        self.SOL.axpy(1.0, self.EPS)

    # algorithmic bytes of one y = A x on level l
    def spmv_bytes(self, l=-1):
        A = self.KK[l]
        n, nnz = A.shape[0], A.nnz
        w = 4 if nnz < 2 ** 31 else 8
        return nnz * 12 + n * (16 + w)
