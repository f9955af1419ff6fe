"""Oracle (TEST INFRASTRUCTURE ONLY): the ASM / Vanka element-block smoother of the reference, restated with
numpy and plain loops for ONE Lagrange variable (the Poisson system of applications/001_Poisson with
"smoother": "asm": SetNumberOfSchurVariables(0), SetElementBlockNumber(n), main.cpp:234-250).

PARITY: the element blocks (DoPartition, with the material flags in the reference's element order on refined boxes and
on the mixed mesh with three material groups) are PINNED to the reference's own output (tests/cpp/ref_partition.cpp on
the host backend of oracle/ref_build -> tests/golden/ref_partition.json, tests/test_asm_cpu.py).  BuildASMIndex and the
application of the smoother stay UNPINNED: they live in the PETSc solver classes (PCASM), which cannot be built here, and
the reference ships no expected index sets or residuals for them.  What is restated (paths relative to /root/reference/src)
  06_mesh/00_single_level/02_partitioning/MeshASMPartitioning.cpp:89-148      DoPartition: consecutive owned
      elements of one material, block_size at a time; materials in the order 4 (solid), 3 (porous), 2 (fluid)
  08_algebra.../03_solvers_with_preconditioner/petsc_asm/LinearEquationSolverPetscAsm.cpp:91-262   BuildASMIndex:
      per block the sorted "overlapping" index set (dofs of the block's elements owned by this rank, then the
      ghost dofs) and the sorted "local" one (dofs no earlier block owns)
  LinearEquationSolverPetscAsm.cpp:266-340, 03_algebra/02_preconditioners/PetscPreconditioner.cpp:179-184
      PCASM, PC_ASM_BASIC, local type PC_COMPOSITE_MULTIPLICATIVE, the user's index sets, overlap 0
and the published algorithm of PETSc 3.20's PCApply_ASM for that configuration on one rank:
      y = 0;  for i = 0 .. nblocks-1:  y[B_i] += solve(A[B_i, B_i], (r - A y)[B_i])
with B_i the overlapping set (PC_ASM_BASIC adds the whole overlapping correction; the local sets only matter to
the RESTRICT / INTERPOLATE variants).  Sub-solves: exact (what MLU_PRECOND on the blocks gives; stated choice --
001_Poisson's own SOR_PRECOND sub-preconditioner is the "ssor" option: PCSOR's default local symmetric sweep,
omega 1, one iteration, zero initial guess).  The level smoother is KSPRICHARDSON with the reference's scale
factor around that preconditioner (LinearEquationSolverPetsc.cpp:516-519)."""
import numpy as np
import scipy.sparse as sp

FLAG_BLOCK = (4, 3, 2)          # MeshASMPartitioning.cpp:100


def do_partition(material, elem_offset, iproc, block_size):
    """MeshASMPartitioning::DoPartition.  material[nel], elem_offset[nprocs+1], block_size[3] per material class
    (solid, porous, fluid).  Returns (block_elements: list of lists, block_type_range[3])."""
    e0, e1 = int(elem_offset[iproc]), int(elem_offset[iproc + 1])
    owned = e1 - e0
    counter = [0, 0, 0]
    for iel in range(e0, e1):
        if material[iel] == FLAG_BLOCK[0]:
            counter[0] += 1
        elif material[iel] == FLAG_BLOCK[1]:
            counter[1] += 1
    counter[2] = owned - counter[0] - counter[1]
    block_elements = []
    block_type_range = [0, 0, 0]
    block_start = 0
    for im in range(3):
        if counter[im] != 0:
            bs = int(block_size[im])
            rem = counter[im] % bs
            blocks = counter[im] // bs if rem == 0 else counter[im] // bs + 1
            for i in range(blocks):
                block_elements.append([0] * bs)
            if rem != 0:
                block_elements[block_start + blocks - 1] = [0] * rem
            c = 0
            for iel in range(e0, e1):
                if material[iel] == FLAG_BLOCK[im]:
                    block_elements[block_start + c // bs][c % bs] = iel
                    c += 1
            block_type_range[im] = block_start + blocks
            block_start += blocks
        else:
            block_type_range[im] = block_start
    return block_elements, block_type_range


def build_asm_index(elem_dofs, dof_offset, iproc, block_elements):
    """BuildASMIndex for one non-Schur variable (NSchurVar = 0 => FastVankaBlock, near elements = the element
    itself).  elem_dofs[e] = system dofs of element e (one variable: system dof = solution dof), dof_offset
    [nprocs+1] of the variable's family.  Returns (local_is, overlapping_is): lists of sorted int64 arrays."""
    d0, d1 = int(dof_offset[iproc]), int(dof_offset[iproc + 1])
    size = d1 - d0
    NONE = size
    indexa = [NONE] * size
    indexb = [NONE] * size
    owned = [False] * size
    local_is, over_is = [], []
    for elems in block_elements:
        loc, ovl = [], []
        ghosts = {}
        seen_el = set()
        for iel in elems:
            jel = iel                                   # GetElementNearElementSize(iel, 0) == 1: the element itself
            if jel in seen_el:
                continue
            seen_el.add(jel)
            for kk in elem_dofs[jel]:
                kk = int(kk)
                if d0 <= kk < d1:
                    if indexa[kk - d0] == NONE and not owned[kk - d0]:
                        owned[kk - d0] = True
                        indexa[kk - d0] = len(loc)
                        loc.append(kk)
                    if indexb[kk - d0] == NONE:
                        indexb[kk - d0] = len(ovl)
                        ovl.append(kk)
                else:
                    ghosts[kk] = True
        for kk in loc:
            indexa[kk - d0] = NONE
        for kk in ovl:
            indexb[kk - d0] = NONE
        ovl = ovl + sorted(ghosts)
        local_is.append(np.array(sorted(loc), dtype=np.int64))
        over_is.append(np.array(sorted(ovl), dtype=np.int64))
    return local_is, over_is


def level_blocks(L, elem_dofs, family_index, block_elems, iproc=0):
    """Blocks of a mesh level: block_elems elements per block for every material class (SetElementBlockNumber),
    material 2 where the level carries none (generated boxes, MeshGeneration: fluid)."""
    material = getattr(L, "material", None)
    if material is None or len(material) == 0:
        material = np.full(L.nel, 2, dtype=np.int64)
    nb = min(int(block_elems), int(L.nel))              # LinearImplicitSystem.cpp:1198
    be, rng = do_partition(material, L.elem_offset, iproc, (nb, nb, nb))
    loc, ovl = build_asm_index(elem_dofs, L.dof_offset[family_index], iproc, be)
    return be, rng, loc, ovl


def schedule(A, blocks):
    """Dependency levels of the multiplicative sweep: block j must run after every earlier block that writes a
    dof j reads or reads a dof j writes; blocks of one level commute, so sweeping level by level (blocks of a
    level in any order, or at once) equals the sequential sweep.  Returns level[nblocks]."""
    A = sp.csr_matrix(A)
    n = A.shape[0]
    wlev = np.full(n, -1, dtype=np.int64)      # highest level of an earlier block writing the dof
    rlev = np.full(n, -1, dtype=np.int64)      # ... reading the dof
    level = np.zeros(len(blocks), dtype=np.int64)
    for j, B in enumerate(blocks):
        cols = np.unique(np.concatenate([A.indices[A.indptr[r]:A.indptr[r + 1]] for r in B] + [np.asarray(B)]))
        lv = max(int(wlev[cols].max()), int(rlev[B].max())) + 1
        level[j] = lv
        wlev[B] = np.maximum(wlev[B], lv)
        rlev[cols] = np.maximum(rlev[cols], lv)
    return level


class BlockSmoother:
    """M^-1 of PCASM (basic, multiplicative, overlap 0) on one rank, and Richardson around it."""

    def __init__(self, A, blocks, sub="lu", order=None):
        """order: the sweep order of the blocks (None: as listed, the reference's)."""
        self.order = order
        self.A = sp.csr_matrix(A)
        self.blocks = [np.asarray(b, dtype=np.int64) for b in blocks]
        self.sub = sub
        self.dense = [self.A[b][:, b].toarray() for b in self.blocks]
        self._ilu = {}

    def _ilu0(self, i):
        """ILU(0) of the block in its sorted dofs on the pattern of A[B, B] (PCILU defaults: levels 0, natural
        ordering): IKJ elimination restricted to the pattern; L (unit) and U share one dense array here."""
        if i not in self._ilu:
            b = self.blocks[i]
            pat = (self.A[b][:, b] != 0).toarray() | (sp.csr_matrix((np.ones(self.A.nnz), self.A.indices, self.A.indptr),
                                                                      shape=self.A.shape)[b][:, b].toarray() != 0)
            F = self.dense[i].copy()
            m = F.shape[0]
            for r in range(m):
                cr = np.nonzero(pat[r])[0]
                for k in cr[cr < r]:
                    F[r, k] = F[r, k] / F[k, k]
                    js = cr[cr > k]
                    js = js[pat[k, js]]
                    F[r, js] -= F[r, k] * F[k, js]
            self._ilu[i] = (F, pat)
        return self._ilu[i]

    def _subsolve(self, i, t):
        M = self.dense[i]
        if self.sub == "lu":
            return np.linalg.solve(M, t)
        if self.sub == "ilu":
            F, pat = self._ilu0(i)
            m = len(t)
            z = t.copy()
            for r in range(m):
                z[r] -= (F[r, :r] * pat[r, :r]) @ z[:r]
            for r in range(m - 1, -1, -1):
                z[r] = (z[r] - (F[r, r + 1:] * pat[r, r + 1:]) @ z[r + 1:]) / F[r, r]
            return z
        # PCSOR default: SOR_LOCAL_SYMMETRIC_SWEEP, omega = 1, its = 1, zero initial guess
        z = np.zeros_like(t)
        m = len(t)
        for r in range(m):
            z[r] = (t[r] - M[r, :r] @ z[:r]) / M[r, r]
        for r in range(m - 1, -1, -1):
            z[r] = (t[r] - M[r, :r] @ z[:r] - M[r, r + 1:] @ z[r + 1:]) / M[r, r]
        return z

    def apply(self, r, order=None):
        y = np.zeros_like(r)
        order = self.order if order is None else order
        for i in (range(len(self.blocks)) if order is None else order):
            b = self.blocks[i]
            t = r[b] - self.A[b] @ y
            y[b] += self._subsolve(i, t)
        return y

    def richardson(self, x, b, nsweeps, scale):
        for _ in range(nsweeps):
            x = x + scale * self.apply(b - self.A @ x)
        return x


# ---- several variables, Schur variables (Vanka blocks of saddle-point systems) -----------------------------------
def kk_offsets(L, families):
    """LinearEquation::InitPde (LinearEquation.cpp:211-237): KKoffset[var][rank] of the system rows
    [rank][variable][dof]; families[k] = FE family index (0 linear, 1 quadratic, 2 biquadratic) of variable k."""
    nprocs = len(L.elem_offset) - 1
    nv = len(families)
    KK = np.zeros((nv + 1, nprocs), dtype=np.int64)
    for j in range(1, nv + 1):
        f = families[j - 1]
        KK[j, 0] = KK[j - 1, 0] + (L.dof_offset[f][1] - L.dof_offset[f][0])
    for i in range(1, nprocs):
        KK[0, i] = KK[nv, i - 1]
        for j in range(1, nv + 1):
            f = families[j - 1]
            KK[j, i] = KK[j - 1, i] + (L.dof_offset[f][i + 1] - L.dof_offset[f][i])
    return KK


def system_dof(L, KK, families, k, sol_dof):
    """LinearEquation::GetSystemDof (LinearEquation.cpp:76-85) of solution dof `sol_dof` of variable k."""
    f = families[k]
    isub = int(np.searchsorted(L.dof_offset[f], sol_dof, side="right") - 1)
    return int(KK[k, isub] + sol_dof - L.dof_offset[f][isub])


def near_elements(elem_vertices):
    """elem::BuildElementNearElement (Elem.cpp:493-526): the element itself, then every other element sharing a
    vertex with it, ascending.  elem_vertices[e] = vertex nodes of element e."""
    near_vertex = {}
    for e, vs in enumerate(elem_vertices):
        for v in vs:
            near_vertex.setdefault(int(v), []).append(e)
    out = []
    for e, vs in enumerate(elem_vertices):
        others = set()
        for v in vs:
            others.update(near_vertex[int(v)])
        others.discard(e)
        out.append([e] + sorted(others))
    return out


def build_asm_index_system(L, elem_sol_dofs, families, nschur, block_elements, near, iproc=0):
    """LinearEquationSolverPetscAsm::BuildASMIndex (LinearEquationSolverPetscAsm.cpp:91-262) for a system of several
    Lagrange variables, the last `nschur` of them Schur variables (pressure-like): a block takes the non-Schur dofs
    of every owned element NEAR its elements (the elements themselves if there is no Schur variable:
    FastVankaBlock) and the Schur dofs of its own elements.  elem_sol_dofs[k][e] = solution dofs of variable k on
    element e.  Returns (local_is, overlapping_is) in system numbering, sorted."""
    nv = len(families)
    KK = kk_offsets(L, families)
    e0, e1 = int(L.elem_offset[iproc]), int(L.elem_offset[iproc + 1])
    dof0 = int(KK[0, iproc])
    size = int(KK[nv, iproc] - KK[0, iproc])
    NONE = size
    indexa, indexb, owned = [NONE] * size, [NONE] * size, [False] * size
    fast = True if nschur == 0 else False           # Lagrange Schur variable (SolType < 3): not a fast Vanka block
    non_schur = [True] * (nv - nschur) + [False] * nschur
    local_is, over_is = [], []
    for elems in block_elements:
        loc, ovl, ghosts, inblock = [], [], {}, set()

        def add(k, jel):
            f = families[k]
            for jdof in elem_sol_dofs[k][jel]:
                kk = system_dof(L, KK, families, k, int(jdof))
                if L.dof_offset[f][iproc] <= jdof < L.dof_offset[f][iproc + 1]:
                    if indexa[kk - dof0] == NONE and not owned[kk - dof0]:
                        owned[kk - dof0] = True
                        indexa[kk - dof0] = len(loc)
                        loc.append(kk)
                    if indexb[kk - dof0] == NONE:
                        indexb[kk - dof0] = len(ovl)
                        ovl.append(kk)
                else:
                    ghosts[kk] = True

        for iel in elems:
            for jel in (near[iel] if not fast else [iel]):
                if e0 <= jel < e1 and jel not in inblock:
                    inblock.add(jel)
                    for k in range(nv):
                        if non_schur[k]:
                            add(k, jel)
            for k in range(nv):
                if not non_schur[k]:
                    add(k, iel)
        for kk in loc:
            indexa[kk - dof0] = NONE
        for kk in ovl:
            indexb[kk - dof0] = NONE
        ovl = ovl + sorted(ghosts)
        local_is.append(np.array(sorted(loc), dtype=np.int64))
        over_is.append(np.array(sorted(ovl), dtype=np.int64))
    return local_is, over_is
