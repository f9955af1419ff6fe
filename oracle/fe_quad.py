"""Oracle (TEST INFRASTRUCTURE ONLY): the quadrilateral face element of a hexahedron, its Gauss
rule, the surface Jacobian and the Neumann boundary integral of the Poisson assembly, restated in
numpy with the reference's operation order.

Pinned against the compiled reference (oracle/_ref, elem_type_2D) by tests/test_oracle_fe.py and
against the committed fixture tests/golden/fe_quad_ref.npz (tests/golden/make_fe_golden.py).

Restates (paths relative to /root/reference/src/02_reference_geom_elements):
  01_fe/2d/Quadrilateral.cpp:22-31, 49-88          node table Xc / IND, tensor-product shape functions
  02_quadrature/2d/quadrature_Quadrangle.cpp:30-33 "seventh" rule: 16 points, 14-digit constants
  03_fe_evaluations_at_quadrature/ElemType.hpp:1330-1379   elem_type_2D::JacobianSur_type
and applications/001_Poisson/main.cpp:495-594 (boundary-face loop of the Poisson assembly).
"""
import numpy as np

from . import fe_hex

XC2 = np.array([(-1, -1), (1, -1), (1, 1), (-1, 1), (0, -1), (1, 0), (0, 1), (-1, 0), (0, 0)], dtype=np.int64)
IND2 = XC2 + 1
NDOFS2 = {"linear": 4, "biquadratic": 9}

# 2-D "seventh" rule (Gauss3 of quad_gauss): tensor 4x4, first coordinate slowest; weights are
# 14-digit truncations of the products (three distinct values)
_G4 = fe_hex._G4
_W16 = {0: 0.1210029932856, 1: 0.22685185185185, 2: 0.42529330301069}


def gauss_quad(name="seventh"):
    if name != "seventh":
        raise NotImplementedError(name)
    w = np.zeros(16)
    xi = np.zeros((16, 2))
    g = 0
    for a in range(4):
        for b in range(4):
            xi[g] = (_G4[a], _G4[b])
            w[g] = _W16[sum(1 for t in (a, b) if t in (1, 2))]
            g += 1
    return w, xi


def shape2(order, pts):
    """phi[npts, n], dphi[npts, n, 2] at reference points pts[npts, 2]."""
    pts = np.asarray(pts, dtype=np.float64)
    n = NDOFS2[order]
    phi = np.zeros((pts.shape[0], n))
    dphi = np.zeros((pts.shape[0], n, 2))
    for a in range(n):
        l = [fe_hex._lag(order, pts[:, d], int(IND2[a, d])) for d in range(2)]
        phi[:, a] = l[0][0] * l[1][0]
        dphi[:, a, 0] = l[0][1] * l[1][0]
        dphi[:, a, 1] = l[0][0] * l[1][1]
    return phi, dphi


def tables2(order, gauss="seventh"):
    """phi[ng,n], dxi[ng,n], deta[ng,n], w[ng] of the face element."""
    w, xi = gauss_quad(gauss)
    phi, dphi = shape2(order, xi)
    return phi, dphi[:, :, 0].copy(), dphi[:, :, 1].copy(), w


def jacobian_sur(order, X, ig, tabs=None):
    """elem_type_2D::JacobianSur: X[3][>=n] face-node coordinates (only the first n are read).
    Returns (weight, phi[n], normal[3]) with the reference's operation order."""
    phi, dxi, deta, w = tabs if tabs is not None else tables2(order)
    n = NDOFS2[order]
    J = np.zeros((3, 3))
    for i in range(n):
        for d in range(3):
            J[d, 0] += dxi[ig, i] * X[d][i]
            J[d, 1] += deta[ig, i] * X[d][i]
    nx = J[1, 0] * J[2, 1] - J[1, 1] * J[2, 0]
    ny = J[0, 1] * J[2, 0] - J[2, 1] * J[0, 0]
    nz = J[0, 0] * J[1, 1] - J[0, 1] * J[1, 0]
    inv = 1. / np.sqrt(nx * nx + ny * ny + nz * nz)
    nrm = np.array([nx * inv, ny * inv, nz * inv])
    J[:, 2] = nrm
    det = (J[0, 0] * (J[1, 1] * J[2, 2] - J[1, 2] * J[2, 1]) + J[0, 1] * (J[1, 2] * J[2, 0] - J[1, 0] * J[2, 2]) +
           J[0, 2] * (J[1, 0] * J[2, 1] - J[1, 1] * J[2, 0]))
    return det * w[ig], phi[ig].copy(), nrm


def neumann_face(order, X, value, tabs=None):
    """F[i] = sum_g (phi_i(g) * value) * weight_g over one face (main.cpp:524-548), i < n face dofs."""
    tabs = tabs if tabs is not None else tables2(order)
    n = NDOFS2[order]
    F = np.zeros(n)
    for ig in range(16):
        wgt, phi, _ = jacobian_sur(order, X, ig, tabs)
        for i in range(n):
            F[i] += phi[i] * value * wgt
    return F
