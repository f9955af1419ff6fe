"""TEST HELPER (synthetic, not a restatement of a reference routine): a Stokes-type saddle-point matrix on a box
level -- triquadratic velocity (3 components), trilinear pressure -- in the reference's system numbering
[rank][variable][dof] (LinearEquation::GetSystemDof), with the velocity Dirichlet rows set to identity as
MGSetLevel's penalty does.  It exists to exercise velocity-pressure Vanka blocks (index sets with a Schur variable)
on the block smoother; the Navier-Stokes assembly of the reference itself is out of scope so far."""
import numpy as np
import scipy.sparse as sp

from oracle import asm, fe_hex, mesh_box as mb

FAMILIES = ["biquadratic", "biquadratic", "biquadratic", "linear"]


def stokes_matrix(L):
    fi = [mb.FAMILY[f] for f in FAMILIES]
    KK = asm.kk_offsets(L, fi)
    dq, dl = mb.system_dof(L, "biquadratic"), mb.system_dof(L, "linear")
    t2, t1 = fe_hex.tables("biquadratic"), fe_hex.tables("linear")
    X = L.xyz[:, L.conn[:, :27]].transpose(1, 0, 2)
    nel = L.nel
    K = np.zeros((nel, 27, 27))
    B = np.zeros((3, nel, 8, 27))
    for ig in range(t2[4].shape[0]):
        w, _, g = fe_hex.jacobian("biquadratic", X, ig, t2)
        for d in range(3):
            K += g[:, :, None, d] * g[:, None, :, d] * w[:, None, None]
            B[d] -= t1[0][ig][None, :, None] * g[:, None, :, d] * w[:, None, None]
    sysq = [np.vectorize(lambda s, k=k: asm.system_dof(L, KK, fi, k, int(s)))(dq) for k in range(3)]
    sysl = np.vectorize(lambda s: asm.system_dof(L, KK, fi, 3, int(s)))(dl)
    rows, cols, vals = [], [], []
    for d in range(3):
        rows.append(np.repeat(sysq[d], 27, axis=1).ravel()); cols.append(np.tile(sysq[d], (1, 27)).ravel()); vals.append(K.ravel())
        rows.append(np.repeat(sysl, 27, axis=1).ravel()); cols.append(np.tile(sysq[d], (1, 8)).ravel()); vals.append(B[d].ravel())
        rows.append(np.tile(sysq[d], (1, 8)).ravel()); cols.append(np.repeat(sysl, 27, axis=1).ravel()); vals.append(B[d].ravel())
    # pressure-pressure zeros belong to the pattern (the reference's sparsity couples every variable pair)
    rows.append(np.repeat(sysl, 8, axis=1).ravel()); cols.append(np.tile(sysl, (1, 8)).ravel()); vals.append(np.zeros(nel * 64))
    n = int(KK[4, -1])
    A = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n)).tocsr()
    A.sort_indices()
    # velocity Dirichlet rows -> identity, pattern kept (SetPenalty)
    bdc = np.nonzero(mb.bdc_flags(L, "biquadratic") < 1.5)[0]
    A = A.tolil()
    for k in range(3):
        for s in bdc:
            r = asm.system_dof(L, KK, fi, k, int(s))
            A[r, A.rows[r]] = 0.0
            A[r, r] = 1.0
    A = A.tocsr()
    A.sort_indices()
    return A
