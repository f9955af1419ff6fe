// Host-side hexahedral Lagrange element: node table, shape functions, Gauss rule, tables at the
// quadrature points and the element prolongator.  This is what the backend needs from FEMuS's
// layer L2 when it runs standalone; in a real drop-in the same tables come from the application's
// own elem_type_3D (reference src/02_reference_geom_elements/03_fe_evaluations_at_quadrature/
// ElemType.cpp:637-740 and :439-532, bases in 01_fe/3d/Hexahedron.cpp:32-163, 1-D polynomials in
// 01_fe/1d/Edge.hpp:72-104, rule in 02_quadrature/3d/quadrature_Hexahedron.cpp).
#pragma once
#include <array>
#include <cmath>
#include <cstdlib>
#include <map>
#include <vector>

namespace femus_b200 {

enum FEFamily { LINEAR = 0, SERENDIPITY = 1, BIQUADRATIC = 2 };

struct HexElement {
  // Number of element dofs per family: 8 vertices, +12 edge midpoints, +6 face centres + centre.
  static int nve(int family) { return family == LINEAR ? 8 : (family == SERENDIPITY ? 20 : 27); }
  static int face_ndofs(int family) { return family == LINEAR ? 4 : (family == SERENDIPITY ? 8 : 9); }

  // Local node n at (xc[n][0..2]) in {-1,0,1}^3.
  static const int (&xc())[27][3] {
    static const int t[27][3] = {
        {-1, -1, -1}, {1, -1, -1}, {1, 1, -1}, {-1, 1, -1}, {-1, -1, 1}, {1, -1, 1}, {1, 1, 1}, {-1, 1, 1},
        {0, -1, -1}, {1, 0, -1}, {0, 1, -1}, {-1, 0, -1}, {0, -1, 1}, {1, 0, 1}, {0, 1, 1}, {-1, 0, 1},
        {-1, -1, 0}, {1, -1, 0}, {1, 1, 0}, {-1, 1, 0},
        {0, -1, 0}, {1, 0, 0}, {0, 1, 0}, {-1, 0, 0}, {0, 0, -1}, {0, 0, 1}, {0, 0, 0}};
    return t;
  }
  // local node at lattice position (i,j,k) in {0,1,2}^3
  static int node_at(int i, int j, int k) {
    static int lut[27];
    static bool init = false;
    if (!init) {
      for (int n = 0; n < 27; n++) lut[(xc()[n][0] + 1) + 3 * ((xc()[n][1] + 1) + 3 * (xc()[n][2] + 1))] = n;
      init = true;
    }
    return lut[i + 3 * (j + 3 * k)];
  }
  // face f -> its 9 local nodes: 4 vertices, 4 edge midpoints, centre
  static const int (&face_nodes())[6][9] {
    static const int t[6][9] = {{0, 1, 5, 4, 8, 17, 12, 16, 20},  {1, 2, 6, 5, 9, 18, 13, 17, 21},
                                {2, 3, 7, 6, 10, 19, 14, 18, 22}, {3, 0, 4, 7, 11, 16, 15, 19, 23},
                                {0, 3, 2, 1, 11, 10, 9, 8, 24},   {4, 5, 6, 7, 12, 13, 14, 15, 25}};
    return t;
  }

  // 1-D Lagrange polynomial i in {0,1,2} on [-1,1] and its derivative
  static void lag1d(int family, double x, int i, double& v, double& d) {
    if (family == LINEAR) {
      if (i == 0) { v = 0.5 * (1. - x); d = -0.5; }
      else if (i == 2) { v = 0.5 * (1. + x); d = 0.5; }
      else { v = 0.; d = 0.; }
    } else if (family == SERENDIPITY) {  // 1-D factor of the 20-node element: linear at the ends, bubble in the middle
      if (i == 0) { v = 0.5 * (1. - x); d = -0.5; }
      else if (i == 1) { v = (1. - x) * (1. + x); d = -2. * x; }
      else { v = 0.5 * (1. + x); d = 0.5; }
    } else {  // biquadratic
      if (i == 0) { v = 0.5 * x * (x - 1.); d = x - 0.5; }
      else if (i == 1) { v = (1. - x) * (1. + x); d = -2. * x; }
      else { v = 0.5 * x * (1. + x); d = x + 0.5; }
    }
  }
  // phi_a and its reference gradient at point p
  static void shape(int family, int a, const double p[3], double& phi, double g[3]) {
    double v[3], d[3];
    for (int k = 0; k < 3; k++) lag1d(family, p[k], xc()[a][k] + 1, v[k], d[k]);
    if (family == SERENDIPITY && xc()[a][0] * xc()[a][1] * xc()[a][2] != 0) {
      // vertex function of the 20-node element: trilinear x (x.xa + y.ya + z.za - 2) (HexQuadratic, Hexahedron.cpp:167-197)
      const double ix = xc()[a][0], jx = xc()[a][1], kx = xc()[a][2];
      const double c = -2. + ix * p[0] + jx * p[1] + kx * p[2];
      phi = c * v[0] * v[1] * v[2];
      g[0] = v[1] * v[2] * (ix * v[0] + c * d[0]);
      g[1] = v[0] * v[2] * (jx * v[1] + c * d[1]);
      g[2] = v[0] * v[1] * (kx * v[2] + c * d[2]);
      return;
    }
    phi = v[0] * v[1] * v[2];
    g[0] = d[0] * v[1] * v[2];
    g[1] = v[0] * d[1] * v[2];
    g[2] = v[0] * v[1] * d[2];
  }

  // "seventh" Gauss rule: 4x4x4 tensor grid, first coordinate slowest; the reference stores
  // 14-digit constants and truncates the 64 products separately (four distinct weights).
  static constexpr int NG = 64;
  static void gauss_seventh(double w[NG], double xi[NG][3]) {
    static const double p[4] = {-0.86113631159405, -0.33998104358486, 0.33998104358486, 0.86113631159405};
    static const double w3[4] = {0.042091477490532, 0.078911515795071, 0.14794033605678, 0.27735296695391};
    int g = 0;
    for (int a = 0; a < 4; a++)
      for (int b = 0; b < 4; b++)
        for (int c = 0; c < 4; c++, g++) {
          xi[g][0] = p[a]; xi[g][1] = p[b]; xi[g][2] = p[c];
          const int inner = (a == 1 || a == 2) + (b == 1 || b == 2) + (c == 1 || c == 2);
          w[g] = w3[inner];
        }
  }

  // ---- face element (quadrilateral with 4 / 9 nodes: vertices, edge midpoints, centre;
  //      01_fe/2d/Quadrilateral.cpp:22-31) and its "seventh" rule (quadrature_Quadrangle.cpp:30-33):
  //      what elem_type_2D::JacobianSur needs for the Neumann integrals of main.cpp:495-548
  static constexpr int NG2 = 16;
  static void shape2(int family, int a, const double p[2], double& phi, double g[2]) {
    static const int ind[9][2] = {{0, 0}, {2, 0}, {2, 2}, {0, 2}, {1, 0}, {2, 1}, {1, 2}, {0, 1}, {1, 1}};
    if (family == SERENDIPITY) { std::abort(); }
    double v[2], d[2];
    for (int k = 0; k < 2; k++) lag1d(family, p[k], ind[a][k], v[k], d[k]);
    phi = v[0] * v[1];
    g[0] = d[0] * v[1];
    g[1] = v[0] * d[1];
  }
  struct FaceTables {
    int nvf = 0;
    std::vector<double> phi, dxi, deta, w;      // [NG2][nvf] row-major, weights [NG2]
  };
  static FaceTables face_tables(int family) {
    static const double p[4] = {-0.86113631159405, -0.33998104358486, 0.33998104358486, 0.86113631159405};
    static const double w2[3] = {0.1210029932856, 0.22685185185185, 0.42529330301069};
    FaceTables t;
    t.nvf = face_ndofs(family);
    t.phi.resize(NG2 * t.nvf); t.dxi.resize(NG2 * t.nvf); t.deta.resize(NG2 * t.nvf); t.w.resize(NG2);
    int g = 0;
    for (int a = 0; a < 4; a++)
      for (int b = 0; b < 4; b++, g++) {
        const double xi[2] = {p[a], p[b]};
        t.w[g] = w2[(a == 1 || a == 2) + (b == 1 || b == 2)];
        for (int i = 0; i < t.nvf; i++) {
          double ph, gr[2];
          shape2(family, i, xi, ph, gr);
          t.phi[g * t.nvf + i] = ph;
          t.dxi[g * t.nvf + i] = gr[0];
          t.deta[g * t.nvf + i] = gr[1];
        }
      }
    return t;
  }

  // tables [NG][nve] row-major + weights
  struct Tables {
    int nve = 0;
    std::vector<double> phi, dxi, deta, dzeta, w;
  };
  static Tables tables(int family) {
    Tables t;
    t.nve = nve(family);
    t.phi.resize(NG * t.nve); t.dxi.resize(NG * t.nve); t.deta.resize(NG * t.nve); t.dzeta.resize(NG * t.nve);
    t.w.resize(NG);
    double xi[NG][3];
    gauss_seventh(t.w.data(), xi);
    for (int g = 0; g < NG; g++)
      for (int a = 0; a < t.nve; a++) {
        double ph, gr[3];
        shape(family, a, xi[g], ph, gr);
        t.phi[g * t.nve + a] = ph;
        t.dxi[g * t.nve + a] = gr[0];
        t.deta[g * t.nve + a] = gr[1];
        t.dzeta[g * t.nve + a] = gr[2];
      }
    return t;
  }

  // Element prolongator: for the fine point at position (a,b,c) of the parent's 5x5x5 lattice
  // (reference coordinate = a/2 - 1), the coarse shape functions with |phi| >= 1e-14.
  static int prolongator_row(int family, int a, int b, int c, int idx[27], double val[27]) {
    const double p[3] = {a * 0.5 - 1., b * 0.5 - 1., c * 0.5 - 1.};
    int n = 0;
    for (int j = 0; j < nve(family); j++) {
      double ph, g[3];
      shape(family, j, p, ph, g);
      if (std::fabs(ph) >= 1.0e-14) { idx[n] = j; val[n] = ph; n++; }
    }
    return n;
  }
};

}  // namespace femus_b200
