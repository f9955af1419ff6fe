"""Generate tests/golden/fe_hex_ref.npz from the compiled reference FE kernel (oracle/_ref).
Run in the build container (needs /root/reference):  python tests/golden/make_fe_golden.py"""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref, fe_hex  # noqa: E402

assert ref.build(), "oracle/_ref could not be built (is /root/reference present?)"
out = {}
rng = np.random.default_rng(20261017)
rng20 = np.random.default_rng(20261018)      # the 20-node family was added later: own stream, so that nothing else moves
for order in ("linear", "biquadratic", "quadratic"):
    R = ref.RefHex(order)
    rng_ = rng
    if order == "quadratic":
        rng = rng20
    w, xi = R.gauss()
    out[f"{order}_gauss_w"], out[f"{order}_gauss_xi"] = w, xi
    phi, dxi, deta, dzeta = R.tables()
    out[f"{order}_phi"], out[f"{order}_dxi"], out[f"{order}_deta"], out[f"{order}_dzeta"] = phi, dxi, deta, dzeta
    n = R.n
    # 4 elements: the unit-box elements of BASELINE configs and two distorted ones
    Xs = []
    for h in (1. / 32, 1. / 128):
        Xs.append((fe_hex.XC[:n].T + 1.0) * (h / 2))
    for _ in range(2):
        Xs.append((fe_hex.XC[:n].T * 0.05 + 0.3) + rng.standard_normal((3, n)) * 0.004)
    Xs = np.array(Xs)
    Us = rng.standard_normal((Xs.shape[0], n))
    Fs, Bs, Ws, Gs = [], [], [], []
    for X, U in zip(Xs, Us):
        F, B = R.poisson_element(X, U, 1.0)
        Fs.append(F)
        Bs.append(B)
        wj, gj = [], []
        for ig in range(R.ng):
            wt, _, g = R.jacobian(X, ig)
            wj.append(wt)
            gj.append(g)
        Ws.append(wj)
        Gs.append(gj)
    out[f"{order}_X"], out[f"{order}_U"] = Xs, Us
    out[f"{order}_F"], out[f"{order}_B"] = np.array(Fs), np.array(Bs)
    out[f"{order}_weight"], out[f"{order}_gradphi"] = np.array(Ws), np.array(Gs)
    rows = R.prolongator()
    P = np.zeros((R.nf, n))
    pos = np.zeros((R.nf, 3), dtype=np.int64)
    for i, (ch, nd, idx, val) in enumerate(rows):
        P[i, idx] = val
        pos[i] = fe_hex.XC[ch] + fe_hex.XC[nd]
    out[f"{order}_prol"], out[f"{order}_prol_pos2"] = P, pos
    rng = rng_
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "fe_hex_ref.npz"), **out)
print("wrote fe_hex_ref.npz", {k: v.shape for k, v in out.items()})

# ---- face element (Neumann integrals): tests/golden/fe_quad_ref.npz
from oracle import fe_quad  # noqa: E402
outq = {}
for order in ("linear", "biquadratic"):
    Q = ref.RefQuad(order)
    w, xi = Q.gauss()
    outq[f"{order}_gauss_w"], outq[f"{order}_gauss_xi"] = w, xi
    phi, dxi, deta = Q.tables()
    outq[f"{order}_phi"], outq[f"{order}_dxi"], outq[f"{order}_deta"] = phi, dxi, deta
    base = np.array([[0, 1, 1, 0, 0.5, 1, 0.5, 0, 0.5], [0, 0, 1, 1, 0, 0.5, 1, 0.5, 0.5], [0.3] * 9])
    Xs = [base / 128.0, base + 0.07 * rng.standard_normal(base.shape), base[[2, 0, 1]] * 0.3 + 0.02 * rng.standard_normal(base.shape)]
    Ws, Ns, Fs = [], [], []
    for X in Xs:
        wj, nj = [], []
        F = np.zeros(Q.n)
        for ig in range(Q.ng):
            wt, ph, nrm = Q.jacobian_sur(X[:, :Q.n].copy(), ig)
            wj.append(wt)
            nj.append(nrm)
            for i in range(Q.n):
                F[i] += ph[i] * 0.2 * wt          # main.cpp:541-546 with bdc_func = 0.2
        Ws.append(wj)
        Ns.append(nj)
        Fs.append(F)
    outq[f"{order}_X"] = np.array(Xs)
    outq[f"{order}_weight"], outq[f"{order}_normal"], outq[f"{order}_F02"] = np.array(Ws), np.array(Ns), np.array(Fs)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "fe_quad_ref.npz"), **outq)
print("wrote tests/golden/fe_quad_ref.npz")


# ---- tetrahedra (4 / 10 / 15 dofs, 31-point rule): tests/golden/fe_tet_ref.npz
from oracle import fe_tet  # noqa: E402
outt = {}
for order in ("linear", "quadratic", "biquadratic"):
    R = ref.RefTet(order)
    w, xi = R.gauss()
    outt[f"{order}_gauss_w"], outt[f"{order}_gauss_xi"] = w, xi
    phi, dxi, deta, dzeta = R.tables()
    outt[f"{order}_phi"], outt[f"{order}_dxi"], outt[f"{order}_deta"], outt[f"{order}_dzeta"] = phi, dxi, deta, dzeta
    n = R.n
    # the reference tetrahedron scaled to a 1/64 box cell, and three distorted / rotated ones
    Xs = [fe_tet.XC[:n].T / 64.0]
    for _ in range(3):
        M = np.eye(3) + 0.2 * rng.standard_normal((3, 3))
        if np.linalg.det(M) < 0:
            M[:, 0] = -M[:, 0]
        Xs.append(M @ fe_tet.XC[:n].T * 0.2 + 0.3 + rng.standard_normal((3, n)) * 0.001)
    Xs = np.array(Xs)
    Us = rng.standard_normal((Xs.shape[0], n))
    Fs, Bs, Ws, Gs = [], [], [], []
    for X, U in zip(Xs, Us):
        F, B = R.poisson_element(X, U, 1.0)
        Fs.append(F)
        Bs.append(B)
        wj, gj = [], []
        for ig in range(R.ng):
            wt, _, g = R.jacobian(X, ig)
            wj.append(wt)
            gj.append(g)
        Ws.append(wj)
        Gs.append(gj)
    outt[f"{order}_X"], outt[f"{order}_U"] = Xs, Us
    outt[f"{order}_F"], outt[f"{order}_B"] = np.array(Fs), np.array(Bs)
    outt[f"{order}_weight"], outt[f"{order}_gradphi"] = np.array(Ws), np.array(Gs)
    rows = R.prolongator()                       # fine dof i -> (child, child-local node), coarse columns
    P = np.zeros((R.nf, n))
    kv = np.zeros((R.nf, 2), dtype=np.int64)
    for i, (ch, nd, idx, val) in enumerate(rows):
        P[i, idx] = val
        kv[i] = (ch, nd)
    outt[f"{order}_prol"], outt[f"{order}_prol_kvert"] = P, kv
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "fe_tet_ref.npz"), **outt)
print("wrote tests/golden/fe_tet_ref.npz")

# ---- wedges (6 / 15 / 21 dofs, 52-point rule): tests/golden/fe_wedge_ref.npz
from oracle import fe_wedge  # noqa: E402
rng = np.random.default_rng(20261019)
outw = {}
for order in ("linear", "quadratic", "biquadratic"):
    R = ref.RefElem("wedge", order)
    w, xi = R.gauss()
    outw[f"{order}_gauss_w"], outw[f"{order}_gauss_xi"] = w, xi
    phi, dxi, deta, dzeta = R.tables()
    outw[f"{order}_phi"], outw[f"{order}_dxi"], outw[f"{order}_deta"], outw[f"{order}_dzeta"] = phi, dxi, deta, dzeta
    n = R.n
    # the reference wedge scaled to a 1/64 box cell, and three distorted / rotated ones
    Xs = [fe_wedge.XC[:n].T / 64.0]
    for _ in range(3):
        M = np.eye(3) + 0.2 * rng.standard_normal((3, 3))
        if np.linalg.det(M) < 0:
            M[:, 0] = -M[:, 0]
        Xs.append(M @ fe_wedge.XC[:n].T * 0.2 + 0.3 + rng.standard_normal((3, n)) * 0.001)
    Xs = np.array(Xs)
    Us = rng.standard_normal((Xs.shape[0], n))
    Fs, Bs, Ws, Gs = [], [], [], []
    for X, U in zip(Xs, Us):
        F, B = R.poisson_element(X, U, 1.0)
        Fs.append(F)
        Bs.append(B)
        wj, gj = [], []
        for ig in range(R.ng):
            wt, _, g = R.jacobian(X, ig)
            wj.append(wt)
            gj.append(g)
        Ws.append(wj)
        Gs.append(gj)
    outw[f"{order}_X"], outw[f"{order}_U"] = Xs, Us
    outw[f"{order}_F"], outw[f"{order}_B"] = np.array(Fs), np.array(Bs)
    outw[f"{order}_weight"], outw[f"{order}_gradphi"] = np.array(Ws), np.array(Gs)
    rows = R.prolongator()                       # fine dof i -> (child, child-local node), coarse columns
    P = np.zeros((R.nf, n))
    kv = np.zeros((R.nf, 2), dtype=np.int64)
    for i, (ch, nd, idx, val) in enumerate(rows):
        P[i, idx] = val
        kv[i] = (ch, nd)
    outw[f"{order}_prol"], outw[f"{order}_prol_kvert"] = P, kv
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "fe_wedge_ref.npz"), **outw)
print("wrote tests/golden/fe_wedge_ref.npz")


# ---- face elements of every 3-D element (triangles 3 / 6 / 7, quadrilaterals 4 / 8 / 9; 13- and 16-point rules):
#      tests/golden/fe_face_ref.npz.  Own random stream, so that the fixtures above do not move.
rngf = np.random.default_rng(20261019)
outf = {}
REFPTS = {"tri": np.array([[0, 0], [1, 0], [0, 1], [.5, 0], [.5, .5], [0, .5], [1. / 3, 1. / 3]]),
          "quad": np.array([[-1, -1], [1, -1], [1, 1], [-1, 1], [0, -1], [1, 0], [0, 1], [-1, 0], [0, 0]], dtype=float)}
for geom in ("tri", "quad"):
    for order in ("linear", "quadratic", "biquadratic"):
        Q = ref.RefFace(geom, order)
        k = f"{geom}_{order}"
        w, xi = Q.gauss()
        outf[f"{k}_gauss_w"], outf[f"{k}_gauss_xi"] = w, xi
        phi, dxi, deta = Q.tables()
        outf[f"{k}_phi"], outf[f"{k}_dxi"], outf[f"{k}_deta"] = phi, dxi, deta
        # a flat face of a 1/64 cell, a curved one, a rotated + curved one (face nodes in 3-D)
        P = REFPTS[geom][:Q.n]
        flat = np.array([P[:, 0], P[:, 1], 0.3 + 0 * P[:, 0]])
        Xs = [flat / 64.0, flat + 0.05 * rngf.standard_normal(flat.shape), flat[[2, 0, 1]] * 0.3 + 0.02 * rngf.standard_normal(flat.shape)]
        Ws, Ns, Fs = [], [], []
        for X in Xs:
            wj, nj = [], []
            F = np.zeros(Q.n)
            for ig in range(Q.ng):
                wt, ph, nrm = Q.jacobian_sur(np.ascontiguousarray(X), ig)
                wj.append(wt)
                nj.append(nrm)
                for i in range(Q.n):
                    F[i] += ph[i] * 0.2 * wt          # main.cpp:541-546 with bdc_func = 0.2
            Ws.append(wj)
            Ns.append(nj)
            Fs.append(F)
        outf[f"{k}_X"] = np.array(Xs)
        outf[f"{k}_weight"], outf[f"{k}_normal"], outf[f"{k}_F02"] = np.array(Ws), np.array(Ns), np.array(Fs)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "fe_face_ref.npz"), **outf)
print("wrote tests/golden/fe_face_ref.npz")
