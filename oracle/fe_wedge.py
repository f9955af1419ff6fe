"""Oracle (TEST INFRASTRUCTURE ONLY): wedge (triangular prism) Lagrange elements of FEMuS (6 / 15 / 21 dofs),
the "seventh" Gauss rule (52 points) and the element prolongators, restated in numpy from the mathematics:
every basis function is a sum of products of AFFINE factors (triangle barycentrics l0 = 1-x-y, l1 = x,
l2 = y and factors in z), evaluated with the product rule -- not the reference's operation order, so values
agree with the compiled reference to a few ulp rather than bit for bit.

Pinned against the compiled reference (oracle/_ref) by tests/test_oracle_fe.py and the committed fixture
tests/golden/fe_wedge_ref.npz (tests/golden/make_fe_golden.py).

Restates (paths relative to /root/reference/src/02_reference_geom_elements):
  01_fe/3d/Wedge.cpp:27-134          node table Xc / IND, children (fine2CoarseVertexMapping), face dofs
  01_fe/3d/Wedge.cpp:179-330         WedgeLinear / WedgeBiquadratic (triangle x line) / WedgeQuadratic (15-node)
  01_fe/2d/Triangle.hpp:60-170       3 / 6 / 7-node triangle functions
  02_quadrature/3d/quadrature_Wedge.cpp   "seventh" rule: 13-point triangle rule x 4 Gauss-Legendre points
  03_fe_evaluations_at_quadrature/ElemType.cpp:439-532, 637-740

Local nodes (reference element: triangle (0,0) (1,0) (0,1) times z in [-1,1]):
  0-2 bottom vertices, 3-5 top vertices, 6-8 bottom edge midpoints (0,1) (1,2) (2,0), 9-11 top ones,
  12-14 midpoints of the vertical edges, 15-17 centres of the quadrilateral faces (over edges 01, 12, 20),
  18 / 19 centres of the bottom / top triangle, 20 centroid.
"""
import numpy as np

NDOFS = {"linear": 6, "quadratic": 15, "biquadratic": 21}
TRI_EDGES = [(0, 1), (1, 2), (2, 0)]
# node -> (triangle entity, z level): entity 0-2 vertex, 3-5 edge, 6 centre; level 0: z=-1, 1: z=0, 2: z=+1
NODE = ([(v, 0) for v in range(3)] + [(v, 2) for v in range(3)] + [(3 + e, 0) for e in range(3)] + [(3 + e, 2) for e in range(3)] +
        [(v, 1) for v in range(3)] + [(3 + e, 1) for e in range(3)] + [(6, 0), (6, 2), (6, 1)])
_TRI_XY = np.array([[0., 0.], [1., 0.], [0., 1.], [.5, 0.], [.5, .5], [0., .5], [1. / 3., 1. / 3.]])
XC = np.array([[_TRI_XY[t, 0], _TRI_XY[t, 1], float(k - 1)] for t, k in NODE])
EDGES = [(0, 1), (1, 2), (2, 0), (3, 4), (4, 5), (5, 3), (0, 3), (1, 4), (2, 5)]           # mid-edge nodes 6..14
# element faces (Elem.hpp `ig`, Wedge.cpp faceDofs): three quadrilaterals (9 nodes), two triangles (7 nodes)
FACE_NODES = [[0, 1, 4, 3, 6, 13, 9, 12, 15], [1, 2, 5, 4, 7, 14, 10, 13, 16], [2, 0, 3, 5, 8, 12, 11, 14, 17],
              [0, 2, 1, 8, 7, 6, 18], [3, 4, 5, 9, 10, 11, 19]]
FACE_NVERT = [4, 4, 4, 3, 3]
FACE_NDOFS = {"linear": [4, 4, 4, 3, 3], "quadratic": [8, 8, 8, 6, 6], "biquadratic": [9, 9, 9, 7, 7]}
# child j: its 6 vertices as parent local nodes (Wedge.cpp:118-127): 4 children in the lower half, 4 in the upper
CHILD_VERTICES = np.array([[0, 6, 8, 12, 15, 17], [6, 1, 7, 15, 13, 16], [8, 7, 2, 17, 16, 14], [7, 8, 6, 16, 17, 15],
                           [12, 15, 17, 3, 9, 11], [15, 13, 16, 9, 4, 10], [17, 16, 14, 11, 10, 5], [16, 17, 15, 10, 11, 9]])

# affine factors a0 + ax x + ay y + az z
_L = [np.array([1., -1., -1., 0.]), np.array([0., 1., 0., 0.]), np.array([0., 0., 1., 0.])]
_Z, _ONE_M_Z, _ONE_P_Z = np.array([0., 0., 0., 1.]), np.array([1., 0., 0., -1.]), np.array([1., 0., 0., 1.])


def _tri_terms(kind, t):
    """Triangle function of entity t as [(coef, [affine factors])]."""
    b = [_L[0], _L[1], _L[2]]
    if kind == "linear":
        return [(1., [_L[t]])] if t < 3 else []
    out = []
    if t < 3:
        out += [(2., [_L[t], _L[t]]), (-1., [_L[t]])]
    elif t < 6:
        a, c = TRI_EDGES[t - 3]
        out += [(4., [_L[a], _L[c]])]
    if kind == "quadratic":
        return out if t < 6 else []
    if t < 3:
        out += [(3., b)]
    elif t < 6:
        out += [(-12., b)]
    else:
        out += [(27., b)]
    return out


def _z_terms(kind, k):
    if kind == "linear":
        return {0: [(0.5, [_ONE_M_Z])], 2: [(0.5, [_ONE_P_Z])]}.get(k, [])
    return {0: [(-0.5, [_Z, _ONE_M_Z])], 1: [(1., [_ONE_M_Z, _ONE_P_Z])], 2: [(0.5, [_Z, _ONE_P_Z])]}[k]


def _terms(order, a):
    t, k = NODE[a]
    if order in ("linear", "biquadratic"):
        return [(c1 * c2, f1 + f2) for c1, f1 in _tri_terms(order, t) for c2, f2 in _z_terms(order, k)]
    # 15-node serendipity wedge (WedgeQuadratic)
    if t < 3 and k != 1:          # vertex: l (2 l -+ z - 2) (1 -+ z) / 2
        s = -1. if k == 0 else 1.
        return [(0.5, [_L[t], 2. * _L[t] + s * _Z + np.array([-2., 0., 0., 0.]), _ONE_M_Z if k == 0 else _ONE_P_Z])]
    if t >= 3 and k != 1:         # triangle-edge node: 2 la lb (1 -+ z)
        a_, c_ = TRI_EDGES[t - 3]
        return [(2., [_L[a_], _L[c_], _ONE_M_Z if k == 0 else _ONE_P_Z])]
    return [(1., [_L[t], _ONE_M_Z, _ONE_P_Z])]        # vertical mid-edge node: l (1 - z^2)


def shape(order, pts):
    """phi[npts, ndofs], dphi[npts, ndofs, 3] at reference points pts[npts,3]."""
    pts = np.asarray(pts, dtype=np.float64)
    n = NDOFS[order]
    phi = np.zeros((pts.shape[0], n))
    dphi = np.zeros((pts.shape[0], n, 3))
    for a in range(n):
        for coef, fac in _terms(order, a):
            vals = [f[0] + pts @ f[1:] for f in fac]
            v = np.ones(pts.shape[0])
            for x in vals:
                v = v * x
            phi[:, a] += coef * v
            for i, f in enumerate(fac):
                rest = np.ones(pts.shape[0])
                for m, x in enumerate(vals):
                    if m != i:
                        rest = rest * x
                dphi[:, a, :] += coef * rest[:, None] * f[None, 1:]
    return phi, dphi


# "seventh" rule: 13-point triangle rule (centroid, two 3-point orbits, one 6-point orbit) times the 4-point
# Gauss-Legendre rule, z fastest; the reference stores the 52 products truncated separately (outer / inner z weight)
_GZ = (-0.86113631159405, -0.33998104358486, 0.33998104358486, 0.86113631159405)
_TRI13 = ([((0.33333333333333, 0.33333333333333), (-0.026014332327752, -0.048770689906083))] +
          [(p, (0.030544309089101, 0.057263319627501)) for p in ((0.47930806784192, 0.26034596607904), (0.26034596607904, 0.47930806784192), (0.26034596607904, 0.26034596607904))] +
          [(p, (0.009278547190612, 0.017395070613808)) for p in ((0.86973979419557, 0.065130102902216), (0.065130102902216, 0.86973979419557), (0.065130102902216, 0.065130102902216))] +
          [(p, (0.013412197676224, 0.025144682768905)) for p in ((0.63844418856981, 0.048690315425316), (0.63844418856981, 0.31286549600488), (0.048690315425316, 0.63844418856981),
                                                                 (0.048690315425316, 0.31286549600488), (0.31286549600488, 0.63844418856981), (0.31286549600488, 0.048690315425316))])


def gauss_wedge(name="seventh"):
    if name != "seventh":
        raise NotImplementedError(name)
    w, xi = [], []
    for (x, y), (wo, wi) in _TRI13:
        for k, z in enumerate(_GZ):
            xi.append((x, y, z))
            w.append(wo if k in (0, 3) else wi)
    return np.array(w), np.array(xi)


def tables(order, gauss="seventh"):
    w, xi = gauss_wedge(gauss)
    phi, dphi = shape(order, xi)
    return phi, dphi[:, :, 0].copy(), dphi[:, :, 1].copy(), dphi[:, :, 2].copy(), w


def child_points(order):
    """pts[8, n, 3]: parent reference coordinates of local node a of child j (children are affine images:
    triangle-affine in (x, y), linear in z)."""
    n = NDOFS[order]
    lin, _ = shape("linear", XC[:n])              # the 6 vertex functions at the child's own nodes
    out = np.zeros((8, n, 3))
    for j in range(8):
        out[j] = lin @ XC[CHILD_VERTICES[j]]
    return out


def local_prolongator(order):
    """P[8, n, n]: coarse function c at node a of child j; |phi| < 1e-14 dropped (ElemType.cpp:439-532)."""
    n = NDOFS[order]
    pts = child_points(order)
    P = np.zeros((8, n, n))
    for j in range(8):
        phi, _ = shape(order, pts[j])
        phi[np.abs(phi) < 1.0e-14] = 0.0
        P[j] = phi
    return P
