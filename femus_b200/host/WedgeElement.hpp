// Host-side wedge (triangular prism) Lagrange elements of FEMuS (6 / 15 / 21 dofs), the "seventh" Gauss rule
// (52 points), tables at the quadrature points and the element prolongators of the 1 -> 8 refinement; the
// counterpart of HexElement.hpp / TetElement.hpp for elem_type_3D("wedge", ...) (reference
// src/02_reference_geom_elements/01_fe/3d/Wedge.cpp:27-330, 01_fe/2d/Triangle.hpp:60-170,
// 02_quadrature/3d/quadrature_Wedge.cpp, 03_fe_evaluations_at_quadrature/ElemType.cpp:439-532, 637-740).
//
// Every basis function is a sum of products of AFFINE factors a0 + a.(x,y,z) -- the triangle barycentrics
// l0 = 1-x-y, l1 = x, l2 = y and factors in z -- evaluated with the product rule, so the tables agree with the
// reference's hand-expanded expressions to a few ulp (tests/test_host_mesh.py), not bit for bit.
// Local nodes (triangle (0,0) (1,0) (0,1) times z in [-1,1]): 0-2 bottom vertices, 3-5 top vertices, 6-8 bottom
// edge midpoints (0,1) (1,2) (2,0), 9-11 top ones, 12-14 midpoints of the vertical edges, 15-17 centres of the
// quadrilateral faces, 18 / 19 centres of the bottom / top triangle, 20 centroid.
#pragma once
#include <array>
#include <cmath>
#include <cstdlib>
#include <vector>
#include "HexElement.hpp"

namespace femus_b200 {

struct WedgeElement {
  static int nve(int family) { return family == LINEAR ? 6 : (family == SERENDIPITY ? 15 : 21); }
  static int nfaces() { return 5; }
  static int face_nvert(int f) { return f < 3 ? 4 : 3; }
  static int face_ndofs(int f, int family) {
    return f < 3 ? (family == LINEAR ? 4 : (family == SERENDIPITY ? 8 : 9)) : (family == LINEAR ? 3 : (family == SERENDIPITY ? 6 : 7));
  }
  // element face f: vertices, edge nodes, face node (Elem.hpp `ig`, Wedge.cpp faceDofs); -1 pads the triangles
  static const int (&face_nodes())[5][9] {
    static const int t[5][9] = {{0, 1, 4, 3, 6, 13, 9, 12, 15}, {1, 2, 5, 4, 7, 14, 10, 13, 16}, {2, 0, 3, 5, 8, 12, 11, 14, 17},
                                {0, 2, 1, 8, 7, 6, 18, -1, -1}, {3, 4, 5, 9, 10, 11, 19, -1, -1}};
    return t;
  }
  static const int (&edges())[9][2] {      // mid-edge nodes 6..14 (MeshRefinement.hpp edge2VerticesMapping)
    static const int t[9][2] = {{0, 1}, {1, 2}, {2, 0}, {3, 4}, {4, 5}, {5, 3}, {0, 3}, {1, 4}, {2, 5}};
    return t;
  }
  // child j: its 6 vertices as parent local nodes (Wedge.cpp:118-127)
  static const int (&child_vertices())[8][6] {
    static const int t[8][6] = {{0, 6, 8, 12, 15, 17}, {6, 1, 7, 15, 13, 16}, {8, 7, 2, 17, 16, 14}, {7, 8, 6, 16, 17, 15},
                                {12, 15, 17, 3, 9, 11}, {15, 13, 16, 9, 4, 10}, {17, 16, 14, 11, 10, 5}, {16, 17, 15, 10, 11, 9}};
    return t;
  }
  // node -> (triangle entity: 0-2 vertex, 3-5 edge (0,1) (1,2) (2,0), 6 centre; z level 0: -1, 1: 0, 2: +1)
  static void node_entity(int n, int& tri, int& lev) {
    static const int t[21][2] = {{0, 0}, {1, 0}, {2, 0}, {0, 2}, {1, 2}, {2, 2}, {3, 0}, {4, 0}, {5, 0}, {3, 2}, {4, 2}, {5, 2},
                                 {0, 1}, {1, 1}, {2, 1}, {3, 1}, {4, 1}, {5, 1}, {6, 0}, {6, 2}, {6, 1}};
    tri = t[n][0];
    lev = t[n][1];
  }
  static void node_xyz(int n, double p[3]) {
    static const double xy[7][2] = {{0., 0.}, {1., 0.}, {0., 1.}, {.5, 0.}, {.5, .5}, {0., .5}, {1. / 3., 1. / 3.}};
    int t, k;
    node_entity(n, t, k);
    p[0] = xy[t][0]; p[1] = xy[t][1]; p[2] = (double)(k - 1);
  }

  // ---- basis functions: sum of coef * product of affine factors
  using Aff = std::array<double, 4>;            // a0 + a1 x + a2 y + a3 z
  struct Term { double c; std::vector<Aff> f; };
  static Aff L(int v) { return v == 0 ? Aff{1., -1., -1., 0.} : (v == 1 ? Aff{0., 1., 0., 0.} : Aff{0., 0., 1., 0.}); }
  static std::vector<Term> tri_terms(int family, int t) {
    static const int te[3][2] = {{0, 1}, {1, 2}, {2, 0}};
    std::vector<Term> out;
    if (family == LINEAR) { if (t < 3) out.push_back({1., {L(t)}}); return out; }
    if (t < 3) { out.push_back({2., {L(t), L(t)}}); out.push_back({-1., {L(t)}}); }
    else if (t < 6) out.push_back({4., {L(te[t - 3][0]), L(te[t - 3][1])}});
    if (family == SERENDIPITY) return t < 6 ? out : std::vector<Term>();
    out.push_back({t < 3 ? 3. : (t < 6 ? -12. : 27.), {L(0), L(1), L(2)}});      // 7-node triangle: bubble enrichment
    return out;
  }
  static std::vector<Term> z_terms(int family, int k) {
    const Aff z{0., 0., 0., 1.}, m{1., 0., 0., -1.}, p{1., 0., 0., 1.};
    if (family == LINEAR) {
      if (k == 0) return {{0.5, {m}}};
      if (k == 2) return {{0.5, {p}}};
      return {};
    }
    if (k == 0) return {{-0.5, {z, m}}};
    if (k == 1) return {{1., {m, p}}};
    return {{0.5, {z, p}}};
  }
  static std::vector<Term> terms(int family, int a) {
    static const int te[3][2] = {{0, 1}, {1, 2}, {2, 0}};
    int t, k;
    node_entity(a, t, k);
    std::vector<Term> out;
    if (family != SERENDIPITY) {
      for (const Term& u : tri_terms(family, t))
        for (const Term& v : z_terms(family, k)) {
          Term w{u.c * v.c, u.f};
          w.f.insert(w.f.end(), v.f.begin(), v.f.end());
          out.push_back(w);
        }
      return out;
    }
    // 15-node element (WedgeQuadratic, Wedge.cpp:239-300)
    const Aff m{1., 0., 0., -1.}, p{1., 0., 0., 1.};
    if (t < 3 && k != 1) {          // vertex: l (2 l -+ z - 2) (1 -+ z) / 2
      const double s = k == 0 ? -1. : 1.;
      Aff q = L(t);
      for (double& x : q) x *= 2.;
      q[0] -= 2.;
      q[3] += s;
      out.push_back({0.5, {L(t), q, k == 0 ? m : p}});
    } else if (t >= 3 && k != 1) {  // triangle-edge node: 2 la lb (1 -+ z)
      out.push_back({2., {L(te[t - 3][0]), L(te[t - 3][1]), k == 0 ? m : p}});
    } else {                        // vertical mid-edge node: l (1 - z^2)
      out.push_back({1., {L(t), m, p}});
    }
    return out;
  }
  static void shape(int family, int a, const double pt[3], double& phi, double g[3]) {
    phi = 0.;
    g[0] = g[1] = g[2] = 0.;
    for (const Term& m : terms(family, a)) {
      const int nf = (int)m.f.size();
      double v[6];
      for (int i = 0; i < nf; i++) v[i] = m.f[i][0] + m.f[i][1] * pt[0] + m.f[i][2] * pt[1] + m.f[i][3] * pt[2];
      double val = 1.;
      for (int i = 0; i < nf; i++) val *= v[i];
      phi += m.c * val;
      for (int i = 0; i < nf; i++) {
        double rest = 1.;
        for (int q = 0; q < nf; q++) if (q != i) rest *= v[q];
        for (int d = 0; d < 3; d++) g[d] += m.c * rest * m.f[i][1 + d];
      }
    }
  }

  // "seventh" rule: 13-point triangle rule (centroid, two 3-point orbits, one 6-point orbit) times the 4-point
  // Gauss-Legendre rule, z fastest; the reference stores the 52 products truncated separately
  static constexpr int NG = 52;
  static void gauss_seventh(double w[NG], double xi[NG][3]) {
    static const double gz[4] = {-0.86113631159405, -0.33998104358486, 0.33998104358486, 0.86113631159405};
    const double a1 = 0.47930806784192, b1 = 0.26034596607904, a2 = 0.86973979419557, b2 = 0.065130102902216;
    const double a3 = 0.63844418856981, b3 = 0.048690315425316, c3 = 0.31286549600488;
    const double t[13][4] = {{0.33333333333333, 0.33333333333333, -0.026014332327752, -0.048770689906083},
                             {a1, b1, 0.030544309089101, 0.057263319627501}, {b1, a1, 0.030544309089101, 0.057263319627501},
                             {b1, b1, 0.030544309089101, 0.057263319627501},
                             {a2, b2, 0.009278547190612, 0.017395070613808}, {b2, a2, 0.009278547190612, 0.017395070613808},
                             {b2, b2, 0.009278547190612, 0.017395070613808},
                             {a3, b3, 0.013412197676224, 0.025144682768905}, {a3, c3, 0.013412197676224, 0.025144682768905},
                             {b3, a3, 0.013412197676224, 0.025144682768905}, {b3, c3, 0.013412197676224, 0.025144682768905},
                             {c3, a3, 0.013412197676224, 0.025144682768905}, {c3, b3, 0.013412197676224, 0.025144682768905}};
    int g = 0;
    for (int i = 0; i < 13; i++)
      for (int k = 0; k < 4; k++, g++) {
        xi[g][0] = t[i][0]; xi[g][1] = t[i][1]; xi[g][2] = gz[k];
        w[g] = (k == 0 || k == 3) ? t[i][2] : t[i][3];
      }
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

  // parent reference coordinates of local node a of child j: the children are images of the reference wedge
  // under the 6-vertex (linear) map with vertices child_vertices()[j]
  static void child_point(int j, int a, double p[3]) {
    double q[3];
    node_xyz(a, q);
    p[0] = p[1] = p[2] = 0.;
    for (int v = 0; v < 6; v++) {
      double ph, g[3], x[3];
      shape(LINEAR, v, q, ph, g);
      node_xyz(child_vertices()[j][v], x);
      for (int d = 0; d < 3; d++) p[d] += ph * x[d];
    }
  }
  static int prolongator_row(int family, int j, int a, int idx[21], double val[21]) {
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
