// Uniform refinement, prolongators and sparsity for meshes of ANY supported element type -- tetrahedra, wedges,
// hexahedra, and mixtures of them (15 / 21 / 27-node geometry; unknowns of the three Lagrange families).
// Pure-hexahedral meshes keep the specialised code of BoxMesh.hpp (lattice names for the sharded run); both give
// the same result on them (tests/test_host_mesh.py).  Restates (paths relative to the reference's src/):
//   06_mesh/00_single_level/03_refinement/MeshRefinement.cpp:188-507   children 8*iel+j in coarse element
//     order, child vertices through the basis' fine2CoarseVertexMapping, boundary faces through
//     coarse2FineFaceMapping (MeshRefinement.hpp:79-100), mid-edge nodes shared through the two end vertices
//     (:365-417), face and centre nodes (AddFaceDofAndElementDof, :513-621), children inherit the parent's
//     rank, renumbering as on level 0, fine coordinates = P_biquadratic x coarse coordinates (:470-472)
//   08_equations/00_stationary/LinearImplicitSystem.cpp:761-909        BuildProlongatorMatrix
//   08_algebra.../LinearEquation.cpp:407-548                           GetSparsityPatternSize
// The refinement is topological: a new node is named by the sorted vertex tuple of the child edge / child
// face it is the centre of (centres: one per child), and the reference's renumbering by first visit makes the
// temporary numbering irrelevant.
#pragma once
#include <array>
#include <map>
#include "BoxMesh.hpp"

namespace femus_b200 {

namespace detail {
// parent face a child face lies on (all child-face vertices are nodes of the parent face), or -1:
// the geometric content of coarse2FineFaceMapping
struct ChildFaces {
  int parent_face[3][8][6];
  ChildFaces() {
    for (int t = 0; t < 3; t++)
      for (int j = 0; j < 8; j++)
        for (int cf = 0; cf < 6; cf++) {
          parent_face[t][j][cf] = -1;
          if (cf >= ElemTopology::nfaces(t)) continue;
          for (int f = 0; f < ElemTopology::nfaces(t); f++) {
            bool all = true;
            for (int k = 0; k < ElemTopology::face_nvert(t, cf) && all; k++) {
              const int pn = ElemTopology::child_vertex(t, j, ElemTopology::face_node(t, cf, k));
              bool on = false;
              for (int i = 0; i < ElemTopology::face_ndofs(t, f, BIQUADRATIC); i++) on = on || ElemTopology::face_node(t, f, i) == pn;
              all = on;
            }
            if (all) parent_face[t][j][cf] = f;
          }
        }
  }
};
inline const ChildFaces& child_faces() {
  static const ChildFaces t;
  return t;
}
}  // namespace detail

// P of `family` from level C to its refinement F: row of fine dof (child j, node a) = the coarse functions at
// that point; rows are INSERTED, identical from every coarse element that sees the dof, so the first visit
// defines the row.
inline HostCsr BuildGeneralProlongator(const MeshLevel& C, const MeshLevel& F, int family) {
  HostCsr P;
  P.nrows = F.ndofs(family);
  P.ncols = C.ndofs(family);
  struct Local { int cnt[8][27]; int idx[8][27][27]; double val[8][27][27]; };
  std::vector<Local> loc(3);
  bool have[3] = {false, false, false};
  for (int64_t E = 0; E < C.nel; E++) have[C.type_of(E)] = true;
  for (int t = 0; t < 3; t++)
    if (have[t])
      for (int j = 0; j < 8; j++)
        for (int a = 0; a < ElemTopology::nve(t, family); a++)
          loc[t].cnt[j][a] = ElemTopology::prolongator_row(t, family, j, a, loc[t].idx[j][a], loc[t].val[j][a]);
  std::vector<int32_t> len((size_t)P.nrows, -1);
  for (int64_t E = 0; E < C.nel; E++) {
    const int t = C.type_of(E), nve = ElemTopology::nve(t, family);
    for (int j = 0; j < 8; j++)
      for (int a = 0; a < nve; a++) {
        const int32_t r = F.GetSolutionDof(a, C.child_el[E * 8 + j], family);
        if (len[r] < 0) len[r] = loc[t].cnt[j][a];
      }
  }
  P.rowptr.assign(P.nrows + 1, 0);
  for (int64_t r = 0; r < P.nrows; r++) P.rowptr[r + 1] = P.rowptr[r] + (len[r] > 0 ? len[r] : 0);
  P.col.resize(P.rowptr[P.nrows]);
  P.val.resize(P.rowptr[P.nrows]);
  std::vector<char> done((size_t)P.nrows, 0);
  std::vector<std::pair<int32_t, double>> tmp(27);
  for (int64_t E = 0; E < C.nel; E++) {
    const int t = C.type_of(E), nve = ElemTopology::nve(t, family);
    int32_t cd[27];
    for (int c = 0; c < nve; c++) cd[c] = C.GetSolutionDof(c, E, family);
    for (int j = 0; j < 8; j++)
      for (int a = 0; a < nve; a++) {
        const int32_t r = F.GetSolutionDof(a, C.child_el[E * 8 + j], family);
        if (done[r]) continue;
        done[r] = 1;
        const int n = loc[t].cnt[j][a];
        for (int k = 0; k < n; k++) tmp[k] = {cd[loc[t].idx[j][a][k]], loc[t].val[j][a][k]};
        std::sort(tmp.begin(), tmp.begin() + n);
        for (int k = 0; k < n; k++) { P.col[P.rowptr[r] + k] = tmp[k].first; P.val[P.rowptr[r] + k] = tmp[k].second; }
      }
  }
  return P;
}

inline MeshLevel RefineGeneralMesh(MeshLevel& C) {
  const detail::ChildFaces& CF = detail::child_faces();
  MeshLevel F;
  F.level = C.level + 1;
  F.nel = C.nel * 8;
  F.conn.assign((size_t)F.nel * 27, -1);
  F.face.assign((size_t)F.nel * 6, -1);
  F.etype.assign((size_t)F.nel, (uint8_t)HEX);
  if (!C.material.empty()) { F.material.resize(F.nel); F.group.resize(F.nel); }     // children inherit (MeshRefinement.cpp:252-258)
  std::vector<int32_t> part(F.nel);
  int32_t next = (int32_t)C.nnode;      // coarse nodes keep their ids in the temporary numbering
  std::map<std::array<int32_t, 2>, int32_t> edge_node;
  std::map<std::array<int32_t, 4>, int32_t> face_node;      // triangles: fourth entry -1
  // vertices and mid-edge nodes of all children first, then face nodes, then centres: the order of creation is
  // irrelevant after the renumbering, only which entities share a node matters
  for (int64_t E = 0; E < C.nel; E++) {
    const int t = C.type_of(E), nv = ElemTopology::nvert(t), ne = ElemTopology::nedges(t), nf = ElemTopology::nfaces(t);
    const int32_t* cn = &C.conn[E * 27];
    for (int j = 0; j < 8; j++) {
      const int64_t fe = E * 8 + j;
      int32_t* fn = &F.conn[fe * 27];
      F.etype[fe] = (uint8_t)t;
      part[fe] = C.part[E];
      if (!C.material.empty()) { F.material[fe] = C.material[E]; F.group[fe] = C.group[E]; }
      for (int v = 0; v < nv; v++) fn[v] = cn[ElemTopology::child_vertex(t, j, v)];
      for (int e = 0; e < ne; e++) {
        int a, b;
        ElemTopology::edge(t, e, a, b);
        std::array<int32_t, 2> key = {fn[a], fn[b]};
        if (key[0] > key[1]) std::swap(key[0], key[1]);
        auto it = edge_node.find(key);
        if (it == edge_node.end()) it = edge_node.emplace(key, next++).first;
        fn[nv + e] = it->second;
      }
      for (int f = 0; f < nf; f++) {
        std::array<int32_t, 4> key = {-1, -1, -1, -1};
        for (int k = 0; k < ElemTopology::face_nvert(t, f); k++) key[k] = fn[ElemTopology::face_node(t, f, k)];
        std::sort(key.begin(), key.end());
        auto it = face_node.find(key);
        if (it == face_node.end()) it = face_node.emplace(key, next++).first;
        fn[nv + ne + f] = it->second;
      }
      fn[nv + ne + nf] = next++;
      for (int cf = 0; cf < nf; cf++) {
        const int pf = CF.parent_face[t][j][cf];
        if (pf >= 0 && C.face[E * 6 + pf] < -1) F.face[fe * 6 + cf] = C.face[E * 6 + pf];
      }
    }
  }
  F.nnode = next;
  F.FillISvectorDofMapAllFEFamilies(part, C.nprocs, /*drop_unreferenced=*/true);
  C.child_el.resize(C.nel * 8);
  for (int64_t pos = 0; pos < F.nel; pos++) C.child_el[F.elem_order[pos]] = (int32_t)pos;
  // coordinates: x_f = P_biquadratic x_c, each row summed in ascending coarse-node order
  HostCsr P = BuildGeneralProlongator(C, F, BIQUADRATIC);
  F.xyz.assign(3 * F.nnode, 0.0);
  for (int d = 0; d < 3; d++)
    for (int64_t r = 0; r < P.nrows; r++) {
      double s = 0.0;
      for (int64_t k = P.rowptr[r]; k < P.rowptr[r + 1]; k++) s += P.val[k] * C.xyz[d * C.nnode + P.col[k]];
      F.xyz[d * F.nnode + r] = s;
    }
  return F;
}

// LinearEquation::GetSparsityPatternSize + SparseMatrix::init for a single-variable system: every (i, j) pair
// of every element, zeros included, columns sorted.  Host-side like the reference's; the device builder
// (b2_csr_create_from_elements) needs one dof count per element and serves the single-type meshes.
inline HostCsr BuildSparsity(const MeshLevel& L, int family) {
  HostCsr A;
  A.nrows = A.ncols = L.ndofs(family);
  std::vector<std::vector<int32_t>> rows((size_t)A.nrows);
  for (int64_t e = 0; e < L.nel; e++) {
    const int nve = ElemTopology::nve(L.type_of(e), family);
    int32_t d[27];
    for (int i = 0; i < nve; i++) d[i] = L.GetSolutionDof(i, e, family);
    for (int i = 0; i < nve; i++) rows[d[i]].insert(rows[d[i]].end(), d, d + nve);
  }
  A.rowptr.assign(A.nrows + 1, 0);
  for (int64_t r = 0; r < A.nrows; r++) {
    std::sort(rows[r].begin(), rows[r].end());
    rows[r].erase(std::unique(rows[r].begin(), rows[r].end()), rows[r].end());
    A.rowptr[r + 1] = A.rowptr[r] + (int64_t)rows[r].size();
  }
  A.col.resize(A.rowptr[A.nrows]);
  A.val.assign(A.rowptr[A.nrows], 0.0);
  for (int64_t r = 0; r < A.nrows; r++) std::copy(rows[r].begin(), rows[r].end(), A.col.begin() + A.rowptr[r]);
  return A;
}

// pure-hexahedral levels keep the specialised code
inline MeshLevel RefineAnyMesh(MeshLevel& C) { return C.etype.empty() ? RefineMesh(C) : RefineGeneralMesh(C); }
inline HostCsr BuildAnyProlongator(const MeshLevel& C, const MeshLevel& F, int family) {
  return C.etype.empty() ? BuildProlongator(C, F, family) : BuildGeneralProlongator(C, F, family);
}

}  // namespace femus_b200
