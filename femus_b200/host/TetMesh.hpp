// Uniform refinement and prolongators of tetrahedral meshes (15-node geometry; unknowns with 4 / 10 / 15 dofs).
// Restates, for element type TET (paths relative to the reference's src/):
//   06_mesh/00_single_level/03_refinement/MeshRefinement.cpp:188-507   children 8*iel+j in coarse element
//     order, child vertices through fine2CoarseVertexMapping (Tetrahedron.cpp:86-95), boundary faces through
//     coarse2FineFaceMapping (MeshRefinement.hpp:88-93), mid-edge nodes shared through the two end vertices
//     (:365-417), face and centre nodes (AddFaceDofAndElementDof, :513-621), children inherit the parent's
//     rank, renumbering as on level 0, fine coordinates = P_biquadratic x coarse coordinates (:470-472)
//   08_equations/00_stationary/LinearImplicitSystem.cpp:761-909        BuildProlongatorMatrix
// As for hexahedra the refinement is topological: a new node is named by the sorted vertex tuple of the
// child edge / child face it is the centre of (centres: one per child), and the reference's renumbering by
// first visit makes the temporary numbering irrelevant.
#pragma once
#include <array>
#include <map>
#include "BoxMesh.hpp"

namespace femus_b200 {

namespace detail {
// (child, child face) pairs lying on parent face f: all three child-face vertices are nodes of the parent face
struct TetChildFaces {
  int parent_face[8][4];      // parent face the child face lies on, or -1
  TetChildFaces() {
    for (int j = 0; j < 8; j++)
      for (int cf = 0; cf < 4; cf++) {
        parent_face[j][cf] = -1;
        for (int f = 0; f < 4; f++) {
          bool all = true;
          for (int k = 0; k < 3 && all; k++) {
            const int pn = TetElement::child_vertices()[j][TetElement::face_nodes()[cf][k]];
            bool on = false;
            for (int i = 0; i < 6; i++) on = on || TetElement::face_nodes()[f][i] == pn;
            all = on;
          }
          if (all) parent_face[j][cf] = f;
        }
      }
  }
};
inline const TetChildFaces& tet_child_faces() {
  static const TetChildFaces t;
  return t;
}
}  // namespace detail

// P of `family` from tetrahedral level C to its refinement F: row of fine dof (child j, node a) = the coarse
// functions at that point (TetElement::prolongator_row); rows are INSERTED, identical from every coarse
// element that sees the dof, so the first visit defines the row.
inline HostCsr BuildTetProlongator(const MeshLevel& C, const MeshLevel& F, int family) {
  const int nve = TetElement::nve(family);
  HostCsr P;
  P.nrows = F.ndofs(family);
  P.ncols = C.ndofs(family);
  int lidx[8][15][15], lcnt[8][15];
  double lval[8][15][15];
  for (int j = 0; j < 8; j++)
    for (int a = 0; a < nve; a++) lcnt[j][a] = TetElement::prolongator_row(family, j, a, lidx[j][a], lval[j][a]);
  std::vector<int32_t> len((size_t)P.nrows, -1);
  for (int64_t E = 0; E < C.nel; E++)
    for (int j = 0; j < 8; j++)
      for (int a = 0; a < nve; a++) {
        const int32_t r = F.GetSolutionDof(a, C.child_el[E * 8 + j], family);
        if (len[r] < 0) len[r] = lcnt[j][a];
      }
  P.rowptr.assign(P.nrows + 1, 0);
  for (int64_t r = 0; r < P.nrows; r++) P.rowptr[r + 1] = P.rowptr[r] + (len[r] > 0 ? len[r] : 0);
  P.col.resize(P.rowptr[P.nrows]);
  P.val.resize(P.rowptr[P.nrows]);
  std::vector<char> done((size_t)P.nrows, 0);
  std::vector<std::pair<int32_t, double>> tmp(15);
  for (int64_t E = 0; E < C.nel; E++) {
    int32_t cd[15];
    for (int c = 0; c < nve; c++) cd[c] = C.GetSolutionDof(c, E, family);
    for (int j = 0; j < 8; j++)
      for (int a = 0; a < nve; a++) {
        const int32_t r = F.GetSolutionDof(a, C.child_el[E * 8 + j], family);
        if (done[r]) continue;
        done[r] = 1;
        const int n = lcnt[j][a];
        for (int k = 0; k < n; k++) tmp[k] = {cd[lidx[j][a][k]], lval[j][a][k]};
        std::sort(tmp.begin(), tmp.begin() + n);
        for (int k = 0; k < n; k++) { P.col[P.rowptr[r] + k] = tmp[k].first; P.val[P.rowptr[r] + k] = tmp[k].second; }
      }
  }
  return P;
}

inline MeshLevel RefineTetMesh(MeshLevel& C) {
  const detail::TetChildFaces& CF = detail::tet_child_faces();
  MeshLevel F;
  F.level = C.level + 1;
  F.nel = C.nel * 8;
  F.conn.assign((size_t)F.nel * 27, -1);
  F.face.assign((size_t)F.nel * 6, -1);
  F.etype.assign((size_t)F.nel, (uint8_t)TET);
  std::vector<int32_t> part(F.nel);
  int32_t next = (int32_t)C.nnode;      // coarse nodes keep their ids in the temporary numbering
  std::map<std::array<int32_t, 2>, int32_t> edge_node;
  std::map<std::array<int32_t, 3>, int32_t> face_node;
  for (int64_t E = 0; E < C.nel; E++) {
    const int32_t* cn = &C.conn[E * 27];
    for (int j = 0; j < 8; j++) {
      const int64_t fe = E * 8 + j;
      int32_t* fn = &F.conn[fe * 27];
      for (int v = 0; v < 4; v++) fn[v] = cn[TetElement::child_vertices()[j][v]];
      for (int e = 0; e < 6; e++) {
        std::array<int32_t, 2> key = {fn[TetElement::edges()[e][0]], fn[TetElement::edges()[e][1]]};
        if (key[0] > key[1]) std::swap(key[0], key[1]);
        auto it = edge_node.find(key);
        if (it == edge_node.end()) it = edge_node.emplace(key, next++).first;
        fn[4 + e] = it->second;
      }
      for (int f = 0; f < 4; f++) {
        std::array<int32_t, 3> key = {fn[TetElement::faces()[f][0]], fn[TetElement::faces()[f][1]], fn[TetElement::faces()[f][2]]};
        std::sort(key.begin(), key.end());
        auto it = face_node.find(key);
        if (it == face_node.end()) it = face_node.emplace(key, next++).first;
        fn[10 + f] = it->second;
      }
      fn[14] = next++;
      part[fe] = C.part[E];
      for (int cf = 0; cf < 4; cf++) {
        const int pf = CF.parent_face[j][cf];
        if (pf >= 0 && C.face[E * 6 + pf] < -1) F.face[fe * 6 + cf] = C.face[E * 6 + pf];
      }
    }
  }
  F.nnode = next;
  F.FillISvectorDofMapAllFEFamilies(part, C.nprocs, /*drop_unreferenced=*/true);
  C.child_el.resize(C.nel * 8);
  for (int64_t pos = 0; pos < F.nel; pos++) C.child_el[F.elem_order[pos]] = (int32_t)pos;
  // coordinates: x_f = P_biquadratic x_c, each row summed in ascending coarse-node order
  HostCsr P = BuildTetProlongator(C, F, BIQUADRATIC);
  F.xyz.assign(3 * F.nnode, 0.0);
  for (int d = 0; d < 3; d++)
    for (int64_t r = 0; r < P.nrows; r++) {
      double s = 0.0;
      for (int64_t k = P.rowptr[r]; k < P.rowptr[r + 1]; k++) s += P.val[k] * C.xyz[d * C.nnode + P.col[k]];
      F.xyz[d * F.nnode + r] = s;
    }
  return F;
}

// dispatch on the (single) element type of the level
inline MeshLevel RefineAnyMesh(MeshLevel& C) { return C.uniform_type() == TET ? RefineTetMesh(C) : RefineMesh(C); }
inline HostCsr BuildAnyProlongator(const MeshLevel& C, const MeshLevel& F, int family) {
  return C.uniform_type() == TET ? BuildTetProlongator(C, F, family) : BuildProlongator(C, F, family);
}

}  // namespace femus_b200
