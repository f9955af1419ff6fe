"""Oracle (TEST INFRASTRUCTURE ONLY): the steady incompressible Navier-Stokes assembly of the reference's library
routine, restated with numpy and an ANALYTIC Newton Jacobian.

Restates src/08_equations/assemble/03_navier_stokes.hpp:305-376 (the element loop shared by the Navier-Stokes
tutorials): per Gauss point of the velocity element, with u, grad u, p evaluated from the current solution,
    aResV[k][i] += ( nu sum_j dphi_i/dx_j du_k/dx_j + phi_i sum_j u_j du_k/dx_j - p dphi_i/dx_k ) w      (:339-352)
    aResP[i]    += - div u psi_i w                                                                        (:355-359)
    RES = -aRes                                                                                            (:379-389)
    KK += d aRes / d sol   -- the reference records the loop with adept and extracts the exact Jacobian (:391-413);
                              here it is written out:
        d aResV[k][i] / d u_l[j] = delta_kl ( nu grad phi_i . grad phi_j + phi_i u . grad phi_j ) w + phi_i phi_j du_k/dx_l w
        d aResV[k][i] / d p[j]   = - psi_j dphi_i/dx_k w ,   d aResP[i] / d u_l[j] = - psi_i dphi_j/dx_l w ,   d aResP / d p = 0
PINNED TO THE REFERENCE ITSELF (round 2): femus::AssembleNavierStokes_AD, run by the reference's own classes on the host
backend of oracle/ref_build (tests/cpp/ref_stokes.cpp "ns"), produced tests/golden/ref_stokes_ns_*.npz -- the residual,
the Jacobian ADEPT recorded and the boundary pressure block; assemble() + pressure_boundary_rhs() reproduce them to 1e-12
(tests/test_reference_pin_stokes.py).  The Jacobian is also checked against finite differences of the residual
(tests/test_oracle_ns.py)."""
import numpy as np
import scipy.sparse as sp

from . import fe_hex, system as osys


def ns_elements(X, U, P, nu, tabs_v, tabs_p):
    """X[nel,3,nv], U[3][nel,nv], P[nel,np].  Returns aResV[3][nel,nv], aResP[nel,np], D[nel,nv,nv] (the block every
    velocity component has on its diagonal), N[3][3][nel,nv,nv] (Newton coupling phi_i phi_j du_k/dx_l), G[3][nel,nv,np]."""
    nel, nv = X.shape[0], tabs_v[0].shape[1]
    npr = tabs_p[0].shape[1]
    RV = np.zeros((3, nel, nv))
    RP = np.zeros((nel, npr))
    D = np.zeros((nel, nv, nv))
    N = np.zeros((3, 3, nel, nv, nv))
    G = np.zeros((3, nel, nv, npr))
    for ig in range(tabs_v[4].shape[0]):
        w, phi, g = fe_hex.jacobian(None, X, ig, tabs_v)
        psi = tabs_p[0][ig]
        u = np.stack([U[k] @ phi for k in range(3)], axis=1)                         # [nel,3]
        gu = np.stack([np.einsum("eid,ei->ed", g, U[k]) for k in range(3)], axis=1)     # [nel,k,d]
        p = P @ psi
        adv = np.einsum("ej,eij->ei", u, g)                                           # u . grad phi_j  [nel,nv]
        lap = np.einsum("eid,ejd->eij", g, g)
        D += (nu * lap + phi[None, :, None] * adv[:, None, :]) * w[:, None, None]
        mass = phi[:, None] * phi[None, :]
        for k in range(3):
            conv = np.einsum("ej,ej->e", u, gu[:, k, :])
            RV[k] += (nu * np.einsum("eid,ed->ei", g, gu[:, k, :]) + phi[None, :] * conv[:, None] - p[:, None] * g[:, :, k]) * w[:, None]
            G[k] -= g[:, :, None, k] * psi[None, None, :] * w[:, None, None]
            for l in range(3):
                N[k, l] += mass[None, :, :] * (gu[:, k, l] * w)[:, None, None]
        RP -= psi[None, :] * ((gu[:, 0, 0] + gu[:, 1, 1] + gu[:, 2, 2]) * w)[:, None]
    return RV, RP, D, N, G


def assemble(L, mesh, order_v, order_p, sol, nu, tables_of):
    """Jacobian (CSR, explicit zeros kept) and RES = -aRes of one level at the solution `sol` (system numbering)."""
    orders = [order_v] * 3 + [order_p]
    d = osys.elem_system_dofs(L, mesh, orders)
    n = sol.shape[0]
    rows, cols, vals = [], [], []
    rhs = np.zeros(n)
    etype = getattr(L, "etype", None)
    for t in (sorted(set(int(x) for x in etype)) if etype is not None else [0]):
        sel = [e for e in range(L.nel) if etype is None or etype[e] == t]
        tv, tp = tables_of(t, order_v), tables_of(t, order_p)
        nv, npr = tv[0].shape[1], tp[0].shape[1]
        dv = [np.array([d[k][e] for e in sel]) for k in range(3)]
        dp = np.array([d[3][e] for e in sel])
        X = L.xyz[:, L.conn[sel][:, :nv]].transpose(1, 0, 2)
        RV, RP, D, N, G = ns_elements(X, [sol[dv[k]] for k in range(3)], sol[dp], nu, tv, tp)
        for k in range(3):
            for l in range(3):
                rows.append(np.repeat(dv[k], nv, axis=1).ravel()); cols.append(np.tile(dv[l], (1, nv)).ravel())
                vals.append((N[k, l] + (D if k == l else 0.0)).ravel())
            rows.append(np.repeat(dv[k], npr, axis=1).ravel()); cols.append(np.tile(dp, (1, nv)).ravel()); vals.append(G[k].ravel())
            rows.append(np.repeat(dp, nv, axis=1).ravel()); cols.append(np.tile(dv[k], (1, npr)).ravel())
            vals.append(G[k].transpose(0, 2, 1).ravel())
            np.add.at(rhs, dv[k].ravel(), -RV[k].ravel())
        rows.append(np.repeat(dp, npr, axis=1).ravel()); cols.append(np.tile(dp, (1, npr)).ravel()); vals.append(np.zeros(len(sel) * npr * npr))
        np.add.at(rhs, dp.ravel(), -RP.ravel())
    A = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n)).tocsr()
    A.sort_indices()
    return A, rhs


def pressure_boundary_rhs(L, mesh, order_v, order_p, tau_by_set):
    """Boundary pressure term (03_navier_stokes.hpp:196-300): for every boundary face whose set index is a key of
    tau_by_set (the caller's selection of the faces whose normal velocity component is not Dirichlet, :262-266),
    aResV[k][face node i] += phi_i tau n_k weight_g with the face element of the velocity family and the unit normal of
    JacobianSur at the Gauss point; returned as the RES contribution (-aRes) in system numbering."""
    from . import asm, fe_face, mesh_mixed as mm
    orders = [order_v] * 3 + [order_p]
    fi = [mesh.FAMILY[o] for o in orders]
    KK = asm.kk_offsets(L, fi)
    rhs = np.zeros(int(KK[-1, -1]))
    etype = getattr(L, "etype", None)
    tabs = {k: fe_face.tables(k, order_v) for k in ("tri", "quad")}
    for e in range(L.nel):
        t = int(etype[e]) if etype is not None else mm.HEX
        for f in range(mm.NFACES[t]):
            b = -(int(L.face[e, f]) + 1)
            if b <= 0 or b not in tau_by_set:
                continue
            kind = fe_face.face_kind(mm.FACE_NVERT[t][f])
            loc = mm.FACE_NODES[t][f][:fe_face.ndofs(kind, order_v)]
            nodes = L.conn[e, loc]
            X = L.xyz[:, nodes]
            sdof = mm.node_dof(L, order_v, nodes) if hasattr(L, "etype") else _box_node_dof(L, mesh, order_v, nodes)
            for ig in range(tabs[kind][3].shape[0]):
                wt, phi, nrm = fe_face.jacobian_sur(X, ig, tabs[kind])
                for k in range(3):
                    for i, s in enumerate(sdof):
                        rhs[asm.system_dof(L, KK, fi, k, int(s))] -= phi[i] * float(tau_by_set[b]) * nrm[k] * wt
    return rhs


def _box_node_dof(L, mesh, order, nodes):
    """solution dof of family `order` of the given nodes on a box level (mesh_box numbering)."""
    k = mesh.FAMILY[order]
    nodes = np.asarray(nodes)
    if k == 2:
        return nodes.copy()
    p = np.searchsorted(L.dof_offset[2], nodes, side="right") - 1
    return (nodes - L.dof_offset[2][p]) + L.dof_offset[k][p]
