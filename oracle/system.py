"""Oracle (TEST INFRASTRUCTURE ONLY): systems of several Lagrange variables on a mesh level -- row numbering
[rank][variable][dof], element dof lists, sparsity pattern, prolongator, Dirichlet flags -- restated with numpy / scipy
on top of the single-variable oracles (mesh_box, mesh_mixed), whose per-family results it only re-indexes.

PINNED TO THE REFERENCE ITSELF (round 2): system dofs of every variable, KKoffset, sparsity counts, Dirichlet flags per
variable and the system prolongator of a Taylor-Hood system are reference output in tests/golden/ref_stokes_*.npz
(tests/cpp/ref_stokes.cpp on the host backend of oracle/ref_build) and reproduced bit-exactly
(tests/test_reference_pin_stokes.py).  Restates (paths relative to
/root/reference/src/08_algebra.../03_solvers_with_preconditioner and src/08_equations/00_stationary):
  LinearEquation.cpp:76-85, 211-237     GetSystemDof, KKoffset                       (oracle/asm.py: kk_offsets, system_dof)
  LinearEquation.cpp:407-548            GetSparsityPatternSize: every element couples variable i with variable j
  LinearImplicitSystem.cpp:826-909      BuildProlongatorMatrix, variable by variable
"""
import numpy as np
import scipy.sparse as sp

from . import asm


def _solution_dofs(L, mesh, order):
    """list over elements of the solution dofs of family `order`"""
    if hasattr(mesh, "element_dofs"):
        return mesh.element_dofs(L, order)
    return list(mesh.system_dof(L, order))          # one variable: system dof == solution dof (mesh_box)


def row_map(L, mesh, orders, k):
    """system row of every solution dof of variable k"""
    fi = [mesh.FAMILY[o] for o in orders]
    KK = asm.kk_offsets(L, fi)
    n = int(L.dof_offset[fi[k]][-1])
    return np.array([asm.system_dof(L, KK, fi, k, s) for s in range(n)], dtype=np.int64)


def elem_system_dofs(L, mesh, orders):
    maps = [row_map(L, mesh, orders, k) for k in range(len(orders))]
    return [[maps[k][np.asarray(d)] for d in _solution_dofs(L, mesh, orders[k])] for k in range(len(orders))]


def sparsity(L, mesh, orders, pattern=None):
    nv = len(orders)
    d = elem_system_dofs(L, mesh, orders)
    n = int(asm.kk_offsets(L, [mesh.FAMILY[o] for o in orders])[-1, -1])
    rows, cols = [], []
    for e in range(L.nel):
        for i in range(nv):
            for j in range(nv):
                if pattern is not None and not pattern[i][j]:
                    continue
                rows.append(np.repeat(d[i][e], len(d[j][e])))
                cols.append(np.tile(d[j][e], len(d[i][e])))
    rows, cols = np.concatenate(rows), np.concatenate(cols)
    A = sp.csr_matrix((np.ones(len(rows), dtype=np.int8), (rows, cols)), shape=(n, n))
    A.sum_duplicates()
    A.sort_indices()
    return A.indptr.astype(np.int64), A.indices.astype(np.int32)


def prolongator(C, F, mesh, orders):
    fi = [mesh.FAMILY[o] for o in orders]
    nf, nc = int(asm.kk_offsets(F, fi)[-1, -1]), int(asm.kk_offsets(C, fi)[-1, -1])
    P = sp.lil_matrix((nf, nc))
    out = sp.csr_matrix((nf, nc))
    for k, o in enumerate(orders):
        S = mesh.prolongator(C, F, o).tocoo()
        rf, rc = row_map(F, mesh, orders, k), row_map(C, mesh, orders, k)
        out = out + sp.csr_matrix((S.data, (rf[S.row], rc[S.col])), shape=(nf, nc))
    out.sort_indices()
    return out


def bdc(L, mesh, orders, dirichlet_faces_per_var):
    fi = [mesh.FAMILY[o] for o in orders]
    out = np.full(int(asm.kk_offsets(L, fi)[-1, -1]), 2.0)
    for k, o in enumerate(orders):
        out[row_map(L, mesh, orders, k)] = mesh.bdc_flags(L, o, dirichlet_faces_per_var[k])
    return out


class SystemMesh:
    """Adapter that lets oracle.mg.Hierarchy run on a system of several variables: the `mesh` module interface
    (bdc_flags, prolongator, zero_dirichlet, sparsity) for fixed variable families and per-variable Dirichlet sets."""

    def __init__(self, mesh, orders, dirichlet_faces_per_var):
        self.mesh, self.orders, self.dirichlet = mesh, list(orders), list(dirichlet_faces_per_var)

    def bdc_flags(self, L, order, dirichlet_faces=None):
        return bdc(L, self.mesh, self.orders, self.dirichlet)

    def prolongator(self, C, F, order):
        return prolongator(C, F, self.mesh, self.orders)

    def zero_dirichlet(self, P, bdc_f, bdc_c):
        from . import mesh_box
        return mesh_box.zero_dirichlet(P, bdc_f, bdc_c)

    def sparsity(self, L, order):
        return sparsity(L, self.mesh, self.orders)
