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
for order in ("linear", "biquadratic"):
    R = ref.RefHex(order)
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
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "fe_hex_ref.npz"), **out)
print("wrote fe_hex_ref.npz", {k: v.shape for k, v in out.items()})
