"""Oracle (TEST INFRASTRUCTURE ONLY): the face elements of the 3-D elements -- triangles with 3 / 6 / 7 nodes
(faces of tetrahedra and wedges) and quadrilaterals with 4 / 8 / 9 nodes (hexahedra, wedges) -- their "seventh"
Gauss rules, and the Neumann boundary integral of the Poisson assembly over any face.

Pinned against the compiled reference (oracle/_ref, elem_type_2D("tri" | "quad", ...)) by
tests/test_oracle_fe.py and the committed fixture tests/golden/fe_face_ref.npz.

Restates (paths relative to /root/reference/src/02_reference_geom_elements):
  01_fe/2d/Triangle.hpp:60-170, Triangle.cpp            3 / 6 / 7-node triangle (vertices, edge midpoints, centre)
  01_fe/2d/Quadrilateral.cpp:22-31, 49-130              4 / 8 (serendipity) / 9-node quadrilateral
  02_quadrature/2d/quadrature_Triangle.cpp              "seventh": 13 points (the triangle part of the wedge rule)
  03_fe_evaluations_at_quadrature/ElemType.hpp:1330-1379   elem_type_2D::JacobianSur_type (fe_quad.jacobian_sur)
and applications/001_Poisson/main.cpp:495-594 (boundary-face loop).  Quadrilaterals with 4 / 9 nodes: fe_quad.py.
"""
import numpy as np

from . import fe_quad, fe_wedge

NDOFS_TRI = {"linear": 3, "quadratic": 6, "biquadratic": 7}
NDOFS_QUAD = {"linear": 4, "quadratic": 8, "biquadratic": 9}
_TRI_W = (-0.074785022233835, 0.087807628716602, 0.026673617804419, 0.038556880445128)


def gauss_tri(name="seventh"):
    if name != "seventh":
        raise NotImplementedError(name)
    xi = np.array([p for p, _ in fe_wedge._TRI13])
    w = np.array([_TRI_W[0]] + [_TRI_W[1]] * 3 + [_TRI_W[2]] * 3 + [_TRI_W[3]] * 6)
    return w, xi


def shape_tri(order, pts):
    """phi[npts, n], dphi[npts, n, 2]: the triangle functions as sums of products of barycentrics."""
    pts = np.asarray(pts, dtype=np.float64)
    n = NDOFS_TRI[order]
    phi = np.zeros((pts.shape[0], n))
    dphi = np.zeros((pts.shape[0], n, 2))
    for a in range(n):
        for coef, fac in fe_wedge._tri_terms(order, a):
            vals = [f[0] + pts @ f[1:3] for f in fac]
            v = np.ones(pts.shape[0])
            for x in vals:
                v = v * x
            phi[:, a] += coef * v
            for i, f in enumerate(fac):
                rest = np.ones(pts.shape[0])
                for m, x in enumerate(vals):
                    if m != i:
                        rest = rest * x
                dphi[:, a, :] += coef * rest[:, None] * f[None, 1:3]
    return phi, dphi


def shape_quad8(pts):
    """8-node (serendipity) quadrilateral: vertices 1/4 (1+x xa)(1+y ya)(x xa + y ya - 1), edge midpoints
    1/2 (1-x^2)(1+y ya) resp. 1/2 (1+x xa)(1-y^2) (Quadrilateral.cpp, QuadQuadratic)."""
    pts = np.asarray(pts, dtype=np.float64)
    x, y = pts[:, 0], pts[:, 1]
    phi = np.zeros((pts.shape[0], 8))
    dphi = np.zeros((pts.shape[0], 8, 2))
    for a in range(8):
        xa, ya = float(fe_quad.XC2[a, 0]), float(fe_quad.XC2[a, 1])
        if xa != 0 and ya != 0:
            c = x * xa + y * ya - 1.
            phi[:, a] = 0.25 * (1. + x * xa) * (1. + y * ya) * c
            dphi[:, a, 0] = 0.25 * (1. + y * ya) * (xa * c + (1. + x * xa) * xa)
            dphi[:, a, 1] = 0.25 * (1. + x * xa) * (ya * c + (1. + y * ya) * ya)
        elif xa == 0:
            phi[:, a] = 0.5 * (1. - x * x) * (1. + y * ya)
            dphi[:, a, 0] = -x * (1. + y * ya)
            dphi[:, a, 1] = 0.5 * (1. - x * x) * ya
        else:
            phi[:, a] = 0.5 * (1. + x * xa) * (1. - y * y)
            dphi[:, a, 0] = 0.5 * xa * (1. - y * y)
            dphi[:, a, 1] = -y * (1. + x * xa)
    return phi, dphi


def tables(kind, order, gauss="seventh"):
    """(phi[ng,n], dxi[ng,n], deta[ng,n], w[ng]) of the face element; kind = "tri" | "quad"."""
    if kind == "tri":
        w, xi = gauss_tri(gauss)
        phi, dphi = shape_tri(order, xi)
    elif order == "quadratic":
        w, xi = fe_quad.gauss_quad(gauss)
        phi, dphi = shape_quad8(xi)
    else:
        return fe_quad.tables2(order, gauss)
    return phi, dphi[:, :, 0].copy(), dphi[:, :, 1].copy(), w


def face_kind(nvert):
    return "tri" if nvert == 3 else "quad"


def ndofs(kind, order):
    return (NDOFS_TRI if kind == "tri" else NDOFS_QUAD)[order]


def jacobian_sur(X, ig, tabs):
    """elem_type_2D::JacobianSur (ElemType.hpp:1330-1379) for any face element: X[3][>=n] face-node coordinates.
    Returns (weight, phi[n], normal[3]) with the reference's operation order."""
    phi, dxi, deta, w = tabs
    n = phi.shape[1]
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


def neumann_face(X, value, tabs):
    """F[i] = sum_g (phi_i(g) * value) * weight_g over one face (main.cpp:524-548); X[3][n] face nodes."""
    n = tabs[0].shape[1]
    F = np.zeros(n)
    for ig in range(tabs[3].shape[0]):
        wgt, phi, _ = jacobian_sur(X, ig, tabs)
        for i in range(n):
            F[i] += phi[i] * value * wgt
    return F
