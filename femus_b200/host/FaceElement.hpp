// Host-side face elements of the 3-D elements: what elem_type_2D(geom, order, "seventh") holds for the Neumann
// boundary integrals of applications/001_Poisson/main.cpp:495-594 (the reference picks
// _finiteElement[GetElementFaceType][order_ind] per boundary face and calls JacobianSur, ElemType.hpp:1330-1379):
//   quadrilaterals with 4 / 8 / 9 dofs   (01_fe/2d/Quadrilateral.cpp:22-31, 49-130), 16-point rule
//   triangles      with 3 / 6 / 7 dofs   (01_fe/2d/Triangle.hpp:60-170),             13-point rule
// The 4 / 9-node quadrilateral tables are HexElement's (bit-exact with the reference); the 8-node
// quadrilateral and the triangles are evaluated from products of affine factors (the triangle terms are the
// ones the wedge is built from, WedgeElement::tri_terms) and agree with the reference's expanded polynomials to
// a few ulp (tests/test_host_mesh.py, against a restatement pinned to the compiled reference).
#pragma once
#include <vector>
#include "HexElement.hpp"
#include "WedgeElement.hpp"

namespace femus_b200 {

struct FaceElement {
  enum { QUAD = 0, TRI = 1 };
  static int kind_of_nvert(int nvert) { return nvert == 3 ? TRI : QUAD; }
  static int ngauss(int kind) { return kind == TRI ? 13 : 16; }
  static int ndofs(int kind, int family) {
    return kind == TRI ? (family == LINEAR ? 3 : (family == SERENDIPITY ? 6 : 7)) : HexElement::face_ndofs(family);
  }
  struct Tables {
    int nvf = 0, ng = 0;
    std::vector<double> phi, dxi, deta, w;      // [ng][nvf] row-major, weights [ng]
  };

  // 8-node quadrilateral: vertices 1/4 (1+x xa)(1+y ya)(x xa + y ya - 1), mid-edge nodes 1/2 (1-x^2)(1+y ya)
  // resp. 1/2 (1+x xa)(1-y^2)
  static void shape_quad8(int a, const double p[2], double& phi, double g[2]) {
    static const int xc[8][2] = {{-1, -1}, {1, -1}, {1, 1}, {-1, 1}, {0, -1}, {1, 0}, {0, 1}, {-1, 0}};
    const double x = p[0], y = p[1], xa = xc[a][0], ya = xc[a][1];
    if (xc[a][0] != 0 && xc[a][1] != 0) {
      const double c = x * xa + y * ya - 1.;
      phi = 0.25 * (1. + x * xa) * (1. + y * ya) * c;
      g[0] = 0.25 * (1. + y * ya) * (xa * c + (1. + x * xa) * xa);
      g[1] = 0.25 * (1. + x * xa) * (ya * c + (1. + y * ya) * ya);
    } else if (xc[a][0] == 0) {
      phi = 0.5 * (1. - x * x) * (1. + y * ya);
      g[0] = -x * (1. + y * ya);
      g[1] = 0.5 * (1. - x * x) * ya;
    } else {
      phi = 0.5 * (1. + x * xa) * (1. - y * y);
      g[0] = 0.5 * xa * (1. - y * y);
      g[1] = -y * (1. + x * xa);
    }
  }
  // triangle function of node a (0-2 vertices, 3-5 midpoints of (0,1) (1,2) (2,0), 6 centre)
  static void shape_tri(int family, int a, const double p[2], double& phi, double g[2]) {
    phi = 0.;
    g[0] = g[1] = 0.;
    for (const WedgeElement::Term& m : WedgeElement::tri_terms(family, a)) {
      const int nf = (int)m.f.size();
      double v[3];
      for (int i = 0; i < nf; i++) v[i] = m.f[i][0] + m.f[i][1] * p[0] + m.f[i][2] * p[1];
      double val = 1.;
      for (int i = 0; i < nf; i++) val *= v[i];
      phi += m.c * val;
      for (int i = 0; i < nf; i++) {
        double rest = 1.;
        for (int q = 0; q < nf; q++) if (q != i) rest *= v[q];
        for (int d = 0; d < 2; d++) g[d] += m.c * rest * m.f[i][1 + d];
      }
    }
  }
  // "seventh" triangle rule (quadrature_Triangle.cpp): centroid, two 3-point orbits, one 6-point orbit
  static void gauss_tri(double w[13], double xi[13][2]) {
    const double a1 = 0.47930806784192, b1 = 0.26034596607904, a2 = 0.86973979419557, b2 = 0.065130102902216;
    const double a3 = 0.63844418856981, b3 = 0.048690315425316, c3 = 0.31286549600488;
    const double w0 = -0.074785022233835, w1 = 0.087807628716602, w2 = 0.026673617804419, w3 = 0.038556880445128;
    const double t[13][3] = {{0.33333333333333, 0.33333333333333, w0}, {a1, b1, w1}, {b1, a1, w1}, {b1, b1, w1},
                             {a2, b2, w2}, {b2, a2, w2}, {b2, b2, w2},
                             {a3, b3, w3}, {a3, c3, w3}, {b3, a3, w3}, {b3, c3, w3}, {c3, a3, w3}, {c3, b3, w3}};
    for (int i = 0; i < 13; i++) { xi[i][0] = t[i][0]; xi[i][1] = t[i][1]; w[i] = t[i][2]; }
  }

  static Tables tables(int kind, int family) {
    Tables t;
    t.nvf = ndofs(kind, family);
    t.ng = ngauss(kind);
    if (kind == QUAD && family != SERENDIPITY) {
      HexElement::FaceTables h = HexElement::face_tables(family);
      t.phi = h.phi; t.dxi = h.dxi; t.deta = h.deta; t.w = h.w;
      return t;
    }
    t.phi.resize(t.ng * t.nvf); t.dxi.resize(t.ng * t.nvf); t.deta.resize(t.ng * t.nvf); t.w.resize(t.ng);
    double xi[16][2];
    if (kind == TRI) {
      gauss_tri(t.w.data(), xi);
    } else {          // the quadrilateral's 4 x 4 Gauss-Legendre rule (points and weights of HexElement::face_tables)
      HexElement::FaceTables h = HexElement::face_tables(LINEAR);
      static const double p[4] = {-0.86113631159405, -0.33998104358486, 0.33998104358486, 0.86113631159405};
      t.w = h.w;
      for (int a = 0, g = 0; a < 4; a++)
        for (int b = 0; b < 4; b++, g++) { xi[g][0] = p[a]; xi[g][1] = p[b]; }
    }
    for (int g = 0; g < t.ng; g++)
      for (int i = 0; i < t.nvf; i++) {
        double ph, gr[2];
        if (kind == TRI) shape_tri(family, i, xi[g], ph, gr);
        else shape_quad8(i, xi[g], ph, gr);
        t.phi[g * t.nvf + i] = ph;
        t.dxi[g * t.nvf + i] = gr[0];
        t.deta[g * t.nvf + i] = gr[1];
      }
    return t;
  }
};

}  // namespace femus_b200
