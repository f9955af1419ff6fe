"""Oracle (TEST INFRASTRUCTURE ONLY): the steady Stokes assembly of the reference, restated with numpy.

Restates applications/003_NavierStokes/SteadyStokes/main.cpp:290-598 (AssembleMatrixResNS; the loop is written for
dim velocity components U, V(, W) and the pressure P): per Gauss point of the VELOCITY element (Jacobian of the
velocity family's nodes, :448-450; pressure functions phi1 = the pressure element's table at the same point)
    F_u[k][i]  += (-IRe sum_d dphi2_i/dx_d dU_k/dx_d + P dphi2_i/dx_k) w                  (:486-494)
    B[k][k][i][j] += IRe sum_d dphi2_i/dx_d dphi2_j/dx_d w                                (:498-508)
    B[k][p][i][j] -= dphi2_i/dx_k phi1_j w                                                (:511-515)
    F_p[i]     += phi1_i div U w                                                          (:521-528; the PSPG term is x 0.)
    B[p][k][i][j] -= phi1_i dphi2_j/dx_k w                                                (:533-538)
    B[p][p]       -= hk^2 / (4 IRe) alpha ...   with alpha = 0 unless both families are linear (:336-340, 541-552):
                     zero VALUES for Taylor-Hood pairs, but the block is added, so it belongs to the pattern
and scatters them with add_matrix_blocked / add_vector_blocked into the system rows [rank][variable][dof] (:575-590).
PINNED TO THE REFERENCE ITSELF (round 2): the callback, compiled in place from the reference's unmodified main.cpp and run
by the reference's own classes on the host backend of oracle/ref_build (tests/cpp/ref_stokes.cpp), produced
tests/golden/ref_stokes_*.npz -- pattern of the assembled matrix, its values, the residual at given fields, the Galerkin
operator below; assemble() reproduces them to 1e-13 (tests/test_reference_pin_stokes.py).  Equal-order linear pairs
(alpha != 0) are not restated."""
import numpy as np
import scipy.sparse as sp

from . import asm, fe_hex, system as osys


def stokes_elements(X, U, P, IRe, tabs_v, tabs_p):
    """X[nel,3,nv] velocity-node coordinates, U[3][nel,nv], P[nel,np].  Returns K[nel,nv,nv] (= B[k][k] / IRe),
    G[3][nel,nv,np] (= B[k][p]; B[p][k] is its transpose), Fu[3][nel,nv], Fp[nel,np]."""
    nel, nv = X.shape[0], tabs_v[0].shape[1]
    npr = tabs_p[0].shape[1]
    K = np.zeros((nel, nv, nv))
    G = np.zeros((3, nel, nv, npr))
    Fu = np.zeros((3, nel, nv))
    Fp = np.zeros((nel, npr))
    for ig in range(tabs_v[4].shape[0]):
        w, _, g = fe_hex.jacobian(None, X, ig, tabs_v)
        phi1 = tabs_p[0][ig]
        gradU = [np.einsum("eid,ei->ed", g, U[k]) for k in range(3)]
        Pg = P @ phi1
        div = gradU[0][:, 0] + gradU[1][:, 1] + gradU[2][:, 2]
        for k in range(3):
            lap_rhs = np.einsum("eid,ed->ei", g, gradU[k])
            Fu[k] += (-IRe * lap_rhs + Pg[:, None] * g[:, :, k]) * w[:, None]
            G[k] -= g[:, :, None, k] * phi1[None, None, :] * w[:, None, None]
        for d in range(3):
            K += g[:, :, None, d] * g[:, None, :, d] * w[:, None, None]
        Fp += phi1[None, :] * (div * w)[:, None]
    return K, G, Fu, Fp


def assemble(L, mesh, order_v, order_p, sol, IRe, tables_of):
    """System matrix (CSR on the pattern of oracle.system.sparsity, explicit zeros kept) and residual of one level.
    sol: the current solution in system numbering; tables_of(element type, order) -> FE tables."""
    orders = [order_v] * 3 + [order_p]
    d = osys.elem_system_dofs(L, mesh, orders)
    n = sol.shape[0]
    rows, cols, vals = [], [], []
    rhs = np.zeros(n)
    etype = getattr(L, "etype", None)
    types = sorted(set(int(t) for t in etype)) if etype is not None else [0]
    for t in types:
        sel = [e for e in range(L.nel) if etype is None or etype[e] == t]
        tv, tp = tables_of(t, order_v), tables_of(t, order_p)
        nv, npr = tv[0].shape[1], tp[0].shape[1]
        dv = [np.array([d[k][e] for e in sel]) for k in range(3)]
        dp = np.array([d[3][e] for e in sel])
        X = L.xyz[:, L.conn[sel][:, :nv]].transpose(1, 0, 2)
        K, G, Fu, Fp = stokes_elements(X, [sol[dv[k]] for k in range(3)], sol[dp], IRe, tv, tp)
        for k in range(3):
            rows.append(np.repeat(dv[k], nv, axis=1).ravel()); cols.append(np.tile(dv[k], (1, nv)).ravel()); vals.append((IRe * K).ravel())
            rows.append(np.repeat(dv[k], npr, axis=1).ravel()); cols.append(np.tile(dp, (1, nv)).ravel()); vals.append(G[k].ravel())
            rows.append(np.repeat(dp, nv, axis=1).ravel()); cols.append(np.tile(dv[k], (1, npr)).ravel())
            vals.append(G[k].transpose(0, 2, 1).ravel())
            np.add.at(rhs, dv[k].ravel(), Fu[k].ravel())
        rows.append(np.repeat(dp, npr, axis=1).ravel()); cols.append(np.tile(dp, (1, npr)).ravel()); vals.append(np.zeros(len(sel) * npr * npr))
        np.add.at(rhs, dp.ravel(), Fp.ravel())
    A = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n)).tocsr()
    A.sort_indices()
    return A, rhs
