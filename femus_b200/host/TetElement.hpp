// Host-side tetrahedral Lagrange elements of FEMuS (4 / 10 / 15 dofs), the "seventh" Gauss rule, tables at
// the quadrature points and the element prolongators of the 1 -> 8 refinement.  Like HexElement.hpp this is
// what the backend needs from FEMuS's layer L2 when it runs standalone; in a drop-in the same tables come
// from the application's elem_type_3D("tet", ...) (reference src/02_reference_geom_elements/
// 03_fe_evaluations_at_quadrature/ElemType.cpp:637-740 and :439-532; bases in 01_fe/3d/Tetrahedron.cpp:147-668,
// node / child tables :25-98, rule in 02_quadrature/3d/quadrature_Tetrahedron.cpp).
//
// Every basis function is a short polynomial in the barycentric coordinates l0 = 1-x-y-z, l1 = x, l2 = y,
// l3 = z (vertex v <-> l_v); values and gradients come from the product rule, so the tables agree with the
// reference's hand-expanded expressions to a few ulp (tests/test_host_mesh.py), not bit for bit.
// Local nodes: vertices 0-3; edge midpoints 4-9 on (0,1) (1,2) (2,0) (0,3) (1,3) (2,3); face centres 10-13
// on (0,1,2) (0,1,3) (1,2,3) (0,2,3); centroid 14.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>
#include "HexElement.hpp"

namespace femus_b200 {

struct TetElement {
  static int nve(int family) { return family == LINEAR ? 4 : (family == SERENDIPITY ? 10 : 15); }
  static int face_ndofs(int family) { return family == LINEAR ? 3 : (family == SERENDIPITY ? 6 : 7); }
  static const int (&edges())[6][2] {
    static const int t[6][2] = {{0, 1}, {1, 2}, {2, 0}, {0, 3}, {1, 3}, {2, 3}};
    return t;
  }
  static const int (&faces())[4][3] {
    static const int t[4][3] = {{0, 1, 2}, {0, 1, 3}, {1, 2, 3}, {0, 2, 3}};
    return t;
  }
  // element face f -> 3 vertices, 3 edge nodes, face node (Elem.hpp `ig`, Tetrahedron.cpp faceDofs)
  static const int (&face_nodes())[4][7] {
    static const int t[4][7] = {{0, 2, 1, 6, 5, 4, 10}, {0, 1, 3, 4, 8, 7, 11}, {1, 2, 3, 5, 9, 8, 12}, {2, 0, 3, 6, 7, 9, 13}};
    return t;
  }
  // child j of the 1 -> 8 refinement: its vertices as parent local nodes (4 corner children, then the 4
  // children of the inner octahedron cut along the diagonal 5-7; Tetrahedron.cpp:86-95)
  static const int (&child_vertices())[8][4] {
    static const int t[8][4] = {{0, 4, 6, 7}, {4, 1, 5, 8}, {6, 5, 2, 9}, {7, 8, 9, 3}, {5, 6, 4, 7}, {8, 7, 5, 4}, {7, 9, 8, 5}, {9, 5, 7, 6}};
    return t;
  }
  // barycentric coordinates of local node n
  static void node_bary(int n, double l[4]) {
    for (int v = 0; v < 4; v++) l[v] = 0.;
    if (n < 4) l[n] = 1.;
    else if (n < 10) { l[edges()[n - 4][0]] = 0.5; l[edges()[n - 4][1]] = 0.5; }
    else if (n < 14) { for (int k = 0; k < 3; k++) l[faces()[n - 10][k]] = 1. / 3.; }
    else { for (int v = 0; v < 4; v++) l[v] = 0.25; }
  }
  static void node_xyz(int n, double p[3]) {
    double l[4];
    node_bary(n, l);
    p[0] = l[1]; p[1] = l[2]; p[2] = l[3];
  }

  // ---- basis functions as sums of barycentric monomials
  struct Term { double c; int nv; int v[4]; };
  static std::vector<Term> terms(int family, int a) {
    std::vector<Term> t;
    auto add = [&](double c, std::initializer_list<int> vs) {
      Term m{c, (int)vs.size(), {0, 0, 0, 0}};
      std::copy(vs.begin(), vs.end(), m.v);
      t.push_back(m);
    };
    auto add_face = [&](double c, int f) { add(c, {faces()[f][0], faces()[f][1], faces()[f][2]}); };
    if (family == LINEAR) { add(1., {a}); return t; }
    // P2 part
    if (a < 4) { add(2., {a, a}); add(-1., {a}); }
    else if (a < 10) add(4., {edges()[a - 4][0], edges()[a - 4][1]});
    if (family == SERENDIPITY) return t;
    // 15-node element: P2 enriched with the four face bubbles and the interior bubble (TetBiquadratic)
    auto in_face = [&](int f, int v) { return faces()[f][0] == v || faces()[f][1] == v || faces()[f][2] == v; };
    if (a < 4) {
      for (int f = 0; f < 4; f++) if (in_face(f, a)) add_face(3., f);
      add(-4., {0, 1, 2, 3});
    } else if (a < 10) {
      for (int f = 0; f < 4; f++) if (in_face(f, edges()[a - 4][0]) && in_face(f, edges()[a - 4][1])) add_face(-12., f);
      add(32., {0, 1, 2, 3});
    } else if (a < 14) {
      add_face(27., a - 10);
      add(-108., {0, 1, 2, 3});
    } else {
      add(256., {0, 1, 2, 3});
    }
    return t;
  }
  // phi_a and its reference gradient at p
  static void shape(int family, int a, const double p[3], double& phi, double g[3]) {
    static const double dl[4][3] = {{-1., -1., -1.}, {1., 0., 0.}, {0., 1., 0.}, {0., 0., 1.}};
    const double l[4] = {1. - (p[0] + p[1] + p[2]), p[0], p[1], p[2]};
    phi = 0.;
    g[0] = g[1] = g[2] = 0.;
    for (const Term& m : terms(family, a)) {
      double val = 1.;
      for (int k = 0; k < m.nv; k++) val *= l[m.v[k]];
      phi += m.c * val;
      for (int k = 0; k < m.nv; k++) {
        double rest = 1.;
        for (int q = 0; q < m.nv; q++) if (q != k) rest *= l[m.v[q]];
        for (int d = 0; d < 3; d++) g[d] += m.c * rest * dl[m.v[k]][d];
      }
    }
  }

  // "seventh" rule: Keast's 31 points as the reference stores them (7 significant digits): centroid, three
  // vertex orbits (a,b,b,b) with a on x, y, z, then on l0, the 6 edge midpoints in ascending (x,y,z) order
  // and the 12-point orbit (0.6, 0.2, 0.1, 0.1) in descending (x,y,z) order.
  static constexpr int NG = 31;
  static void gauss_seventh(double w[NG], double xi[NG][3]) {
    int g = 0;
    auto put = [&](double wt, double x, double y, double z) { w[g] = wt; xi[g][0] = x; xi[g][1] = y; xi[g][2] = z; g++; };
    put(0.01826422, 0.25, 0.25, 0.25);
    static const double orb[3][3] = {{0.01059994, 0.7653604, 0.07821319}, {-0.06251774, 0.6344704, 0.1218432}, {0.004891425, 0.002382507, 0.3325392}};
    for (const auto& o : orb)
      for (int pos = 0; pos < 4; pos++) put(o[0], pos == 0 ? o[1] : o[2], pos == 1 ? o[1] : o[2], pos == 2 ? o[1] : o[2]);
    // distinct (x,y,z) of the permutations of a 4-tuple, sorted
    auto orbit = [&](double wt, std::vector<double> q, bool descending) {
      std::sort(q.begin(), q.end());
      std::vector<std::array<double, 3>> pts;
      do {
        std::array<double, 3> p = {q[0], q[1], q[2]};
        if (std::find(pts.begin(), pts.end(), p) == pts.end()) pts.push_back(p);
      } while (std::next_permutation(q.begin(), q.end()));
      std::sort(pts.begin(), pts.end());
      if (descending) std::reverse(pts.begin(), pts.end());
      for (const auto& p : pts) put(wt, p[0], p[1], p[2]);
    };
    orbit(0.0009700176, {0., 0., 0.5, 0.5}, false);
    orbit(0.02755732, {0.6, 0.2, 0.1, 0.1}, true);
    if (g != NG) std::abort();
  }

  static HexElement::Tables tables(int family) {
    HexElement::Tables t;
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

  // parent reference coordinates of local node a of child j (children are affine images of the parent)
  static void child_point(int j, int a, double p[3]) {
    double l[4];
    node_bary(a, l);
    p[0] = p[1] = p[2] = 0.;
    for (int v = 0; v < 4; v++) {
      double q[3];
      node_xyz(child_vertices()[j][v], q);
      for (int d = 0; d < 3; d++) p[d] += l[v] * q[d];
    }
  }
  // Element prolongator row of (child j, local node a): coarse functions with |phi| >= 1e-14
  // (ElemType.cpp:439-532)
  static int prolongator_row(int family, int j, int a, int idx[15], double val[15]) {
    double p[3];
    child_point(j, a, p);
    int n = 0;
    for (int c = 0; c < nve(family); c++) {
      double ph, g[3];
      shape(family, c, p, ph, g);
      if (std::fabs(ph) >= 1.0e-14) { idx[n] = c; val[n] = ph; n++; }
    }
    return n;
  }
};

}  // namespace femus_b200
