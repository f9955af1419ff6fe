// Host-side mesh layer of the backend: structured HEX27 box generation, FEMuS element / node
// numbering, uniform refinement, dof maps, Dirichlet flags and prolongators.  Integer results must
// be bit-identical to what FEMuS produces on the same mesh, because they define the DOF->row map
// and the CSR structure.  Restates (paths relative to the reference's src/):
//   06_mesh/00_single_level/01_input/02_from_implemented_code/MeshGeneration.cpp:790-849, 976-1071
//   06_mesh/00_single_level/00_definition/Mesh.cpp:517-559 (node renumbering), :589-616 (element
//     reorder by rank), :706-853 (dof offsets), :1021-1074 (GetSolutionDof)
//   06_mesh/00_single_level/03_refinement/MeshRefinement.cpp:188-507, :513-621
//   06_mesh/00_single_level/02_partitioning/MeshMetisPartitioning.cpp:143-155 (children inherit rank)
//   06_solution/01_multiple_levels/00_definition/MultiLevelSolution.cpp:725-840 (GenerateBdc)
//   08_equations/00_stationary/LinearImplicitSystem.cpp:761-909 (BuildProlongatorMatrix)
// Refinement is topological (works on any conforming HEX27 mesh, not only boxes): a new node is
// identified by the pair (lowest-numbered corner, opposite corner) of the coarse edge / face
// quadrant / cell octant it is the centre of.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <vector>
#include "HexElement.hpp"
#include "TetElement.hpp"
#include "WedgeElement.hpp"

namespace femus_b200 {

// Geometric element types in the reference's numbering (GeomElTypeEnum: HEX 0, TET 1, WEDGE 2) and the
// per-type tables of Elem.hpp:102-140, 581-615 (NVE, NFC, ig, NFACENODES), MeshRefinement.hpp:118-135
// (edge2VerticesMapping) and the bases' fine2CoarseVertexMapping that the mesh layer needs.  Local node
// classes are the same for every type: [vertices][edge midpoints][face centres][centre], and face f owns the
// node nve(type, SERENDIPITY) + f.
enum GeomType { HEX = 0, TET = 1, WEDGE = 2 };
struct ElemTopology {
  static int nve(int type, int family) {
    return type == TET ? TetElement::nve(family) : (type == WEDGE ? WedgeElement::nve(family) : HexElement::nve(family));
  }
  static int nvert(int type) { return nve(type, LINEAR); }
  static int nedges(int type) { return nve(type, SERENDIPITY) - nve(type, LINEAR); }
  static int nfaces(int type) { return type == TET ? 4 : (type == WEDGE ? 5 : 6); }
  static int face_nvert(int type, int f) { return type == TET ? 3 : (type == WEDGE ? WedgeElement::face_nvert(f) : 4); }
  static int face_ndofs(int type, int f, int family) {
    return type == TET ? TetElement::face_ndofs(family) : (type == WEDGE ? WedgeElement::face_ndofs(f, family) : HexElement::face_ndofs(family));
  }
  static int face_node(int type, int f, int i) {
    return type == TET ? TetElement::face_nodes()[f][i] : (type == WEDGE ? WedgeElement::face_nodes()[f][i] : HexElement::face_nodes()[f][i]);
  }
  static void edge(int type, int e, int& a, int& b) {
    static const int hex_edges[12][2] = {{0, 1}, {1, 2}, {2, 3}, {3, 0}, {4, 5}, {5, 6}, {6, 7}, {7, 4}, {0, 4}, {1, 5}, {2, 6}, {3, 7}};
    const int* p = type == TET ? TetElement::edges()[e] : (type == WEDGE ? WedgeElement::edges()[e] : hex_edges[e]);
    a = p[0];
    b = p[1];
  }
  // vertex v of child j as a parent local node
  static int child_vertex(int type, int j, int v) {
    if (type == TET) return TetElement::child_vertices()[j][v];
    if (type == WEDGE) return WedgeElement::child_vertices()[j][v];
    const int* a = HexElement::xc()[j];          // child j = octant at parent vertex j: node at (xc[j] + xc[v]) / 2
    const int* b = HexElement::xc()[v];
    return HexElement::node_at((a[0] + b[0]) / 2 + 1, (a[1] + b[1]) / 2 + 1, (a[2] + b[2]) / 2 + 1);
  }
  // element prolongator row of (child j, child-local node a): coarse functions with |phi| >= 1e-14
  static int prolongator_row(int type, int family, int j, int a, int* idx, double* val) {
    if (type == TET) return TetElement::prolongator_row(family, j, a, idx, val);
    if (type == WEDGE) return WedgeElement::prolongator_row(family, j, a, idx, val);
    const int* o = HexElement::xc()[j];
    const int* x = HexElement::xc()[a];
    return HexElement::prolongator_row(family, o[0] + x[0] + 2, o[1] + x[1] + 2, o[2] + x[2] + 2, idx, val);
  }
  static int ngauss(int type) { return type == TET ? TetElement::NG : (type == WEDGE ? WedgeElement::NG : HexElement::NG); }
  static HexElement::Tables tables(int type, int family) {
    return type == TET ? TetElement::tables(family) : (type == WEDGE ? WedgeElement::tables(family) : HexElement::tables(family));
  }
};

struct HostCsr {
  int64_t nrows = 0, ncols = 0;
  std::vector<int64_t> rowptr;
  std::vector<int32_t> col;
  std::vector<double> val;
};

class MeshLevel {
 public:
  int level = 0;
  int nprocs = 1;
  int64_t nel = 0, nnode = 0;
  std::vector<int32_t> conn;            // [nel][27] node ids in FEMuS numbering (rows of shorter elements padded with -1)
  std::vector<int32_t> face;            // [nel][6] faceElementIndex: -1 interior/unset, < -1 boundary
  std::vector<uint8_t> etype;           // [nel] GeomType per element; empty = all hexahedra
  std::vector<int16_t> material, group; // [nel] element material / group of a mesh file; empty = one group
  std::vector<int32_t> part;            // [nel] owning rank
  std::vector<int64_t> elem_offset;     // [nprocs+1]
  std::vector<int64_t> dof_offset[3];   // [family][nprocs+1]
  std::vector<double> xyz;              // [3][nnode]
  std::vector<int32_t> child_el;        // [nel][8] fine element of (element, child) once refined
  std::vector<int32_t> ijk;             // [3][nnode] integer lattice coordinates of the nodes on this level's
                                        // (2*nx_l+1)(2*ny_l+1)(2*nz_l+1) lattice: a rank-independent name of a node,
                                        // used to match interface nodes between the sub-meshes of different ranks

  int32_t node(int64_t iel, int i) const { return conn[iel * 27 + i]; }
  int type_of(int64_t iel) const { return etype.empty() ? (int)HEX : (int)etype[iel]; }
  // the one element type of the level, -1 for a mixed mesh
  int uniform_type() const {
    const int t = type_of(0);
    for (int64_t e = 1; e < nel; e++)
      if (type_of(e) != t) return -1;
    return t;
  }

  int owner_of_node(int32_t nd) const {
    const std::vector<int64_t>& o = dof_offset[2];
    return (int)(std::upper_bound(o.begin(), o.end(), (int64_t)nd) - o.begin()) - 1;
  }
  // Mesh::GetSolutionDof (Mesh.cpp:1021-1074), uniform meshes (no "owned ghost" nodes)
  int32_t GetSolutionDof(int i, int64_t iel, int family) const {
    const int32_t nd = node(iel, i);
    if (family == BIQUADRATIC) return nd;
    const int p = owner_of_node(nd);
    return (int32_t)((nd - dof_offset[2][p]) + dof_offset[family][p]);
  }
  int64_t ndofs(int family) const { return dof_offset[family][nprocs]; }

  // [nel][nve] system dofs of a single-variable system (LinearEquation::GetSystemDof,
  // LinearEquation.cpp:76-85: KKoffset[0][p] == dofOffset[family][p])
  std::vector<int32_t> system_dofs(int family) const {
    const int ut = uniform_type();
    if (ut < 0) { std::fprintf(stderr, "femus_b200: system_dofs on a mixed mesh: use system_dofs27\n"); std::abort(); }
    const int nve = ElemTopology::nve(ut, family);
    std::vector<int32_t> d((size_t)nel * nve);
    for (int64_t e = 0; e < nel; e++)
      for (int i = 0; i < nve; i++) d[e * nve + i] = GetSolutionDof(i, e, family);
    return d;
  }
  // the same for meshes of several element types: rows of 27, padded with -1 after nve(type, family) entries
  std::vector<int32_t> system_dofs27(int family) const {
    std::vector<int32_t> d((size_t)nel * 27, -1);
    for (int64_t e = 0; e < nel; e++)
      for (int i = 0; i < ElemTopology::nve(type_of(e), family); i++) d[e * 27 + i] = GetSolutionDof(i, e, family);
    return d;
  }

  // MultiLevelSolution::GenerateBdc: 2 = free, 0 = Dirichlet on every exterior face whose boundary
  // index (1..6 = -(faceElementIndex+1), Elem.cpp:361-364) is flagged in dirichlet_faces[1..6].
  std::vector<double> GenerateBdc(int family, const bool dirichlet_faces[7]) const {
    std::vector<double> bdc((size_t)ndofs(family), 2.0);
    for (int64_t e = 0; e < nel; e++) {
      const int t = type_of(e);
      for (int f = 0; f < ElemTopology::nfaces(t); f++) {
        const int bidx = -(face[e * 6 + f] + 1);
        if (bidx > 0 && bidx <= 6 && dirichlet_faces[bidx])
          for (int iv = 0; iv < ElemTopology::face_ndofs(t, f, family); iv++) bdc[GetSolutionDof(ElemTopology::face_node(t, f, iv), e, family)] = 0.0;
      }
    }
    return bdc;
  }

  // Mesh::FillISvectorDofMapAllFEFamilies: element reorder by rank (stable) and node renumbering
  // by first visit over (rank, family k, element, local node in [nve(k-1), nve(k))).
  // `conn` holds temporary node ids in [0, nnode) on entry.  Returns the node map old -> new.
  std::vector<int32_t> FillISvectorDofMapAllFEFamilies(const std::vector<int32_t>& partition, int nprocs_,
                                                       bool drop_unreferenced = false) {
    nprocs = nprocs_;
    // --- elements by rank (Mesh.cpp:589-616), then inside each rank by (material, group, index) -- the
    // bubble sort of Mesh.cpp:621-702; one stable sort on the three keys gives the same order
    std::vector<int64_t> order(nel);
    std::iota(order.begin(), order.end(), (int64_t)0);
    const bool groups = !material.empty();
    std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) {
      if (partition[a] != partition[b]) return partition[a] < partition[b];
      if (!groups) return false;
      if (material[a] != material[b]) return material[a] < material[b];
      return group[a] < group[b];
    });
    bool identity = true;
    for (int64_t e = 0; e < nel && identity; e++) identity = (order[e] == e);
    part.resize(nel);
    if (identity) {
      part = partition;
    } else {
      std::vector<int32_t> c2(conn.size()), f2(face.size());
      std::vector<uint8_t> t2(etype.size());
      std::vector<int16_t> m2(material.size()), g2(group.size());
      for (int64_t e = 0; e < nel; e++) {
        std::copy(conn.begin() + order[e] * 27, conn.begin() + order[e] * 27 + 27, c2.begin() + e * 27);
        std::copy(face.begin() + order[e] * 6, face.begin() + order[e] * 6 + 6, f2.begin() + e * 6);
        if (!etype.empty()) t2[e] = etype[order[e]];
        if (groups) { m2[e] = material[order[e]]; g2[e] = group[order[e]]; }
        part[e] = partition[order[e]];
      }
      conn.swap(c2);
      face.swap(f2);
      etype.swap(t2);
      material.swap(m2);
      group.swap(g2);
    }
    elem_order.assign(order.begin(), order.end());
    elem_offset.assign(nprocs + 1, 0);
    for (int64_t e = 0; e < nel; e++) elem_offset[part[e] + 1]++;
    for (int p = 0; p < nprocs; p++) elem_offset[p + 1] += elem_offset[p];
    // --- node renumbering (Mesh.cpp:517-559)
    std::vector<int32_t> map(nnode, -1);
    std::vector<std::vector<int64_t>> own(3, std::vector<int64_t>(nprocs, 0));
    int32_t counter = 0;
    for (int p = 0; p < nprocs; p++)
      for (int k = 0; k < 3; k++) {
        for (int64_t e = elem_offset[p]; e < elem_offset[p + 1]; e++) {
          const int t = type_of(e), lo = k == 0 ? 0 : ElemTopology::nve(t, k - 1), hi = ElemTopology::nve(t, k);
          for (int i = lo; i < hi; i++) {
            const int32_t ii = conn[e * 27 + i];
            if (map[ii] < 0) {
              map[ii] = counter++;
              for (int j = k; j < 3; j++) own[j][p]++;
            }
          }
        }
      }
    // temporary ids no element refers to (the face and centre nodes of refined tetrahedra are not nodes of
    // the children) disappear: the reference sets the node count to _dofOffset[2][nprocs] (Mesh.cpp:886-891)
    if (counter != nnode && !drop_unreferenced) { std::fprintf(stderr, "femus_b200: mesh has %lld unreferenced nodes\n", (long long)(nnode - counter)); std::abort(); }
    nnode = counter;
    for (auto& c : conn) if (c >= 0) c = map[c];
    for (int k = 0; k < 3; k++) {
      dof_offset[k].assign(nprocs + 1, 0);
      for (int p = 0; p < nprocs; p++) dof_offset[k][p + 1] = dof_offset[k][p] + own[k][p];
    }
    return map;
  }
  std::vector<int64_t> elem_order;      // position -> element index before the rank reorder
};

// z-slab partition of the level-0 box (elements in k, j, i order)
inline std::vector<int32_t> SlabPartition(int nx, int ny, int nz, int nprocs) {
  std::vector<int32_t> p((size_t)nx * ny * nz);
  for (int k = 0; k < nz; k++)
    for (int64_t t = 0; t < (int64_t)nx * ny; t++) p[(int64_t)k * nx * ny + t] = (int32_t)(((int64_t)k * nprocs) / nz);
  return p;
}

// MeshTools::Generation::BuildBox, HEX27 branch
inline MeshLevel GenerateCoarseBoxMesh(int nx, int ny, int nz, double xmin, double xmax, double ymin, double ymax,
                                       double zmin, double zmax, const std::vector<int32_t>* partition = nullptr,
                                       int nprocs = 1) {
  MeshLevel L;
  L.level = 0;
  L.nel = (int64_t)nx * ny * nz;
  const int64_t sx = 2 * nx + 1, sy = 2 * ny + 1, sz = 2 * nz + 1;
  L.nnode = sx * sy * sz;
  L.conn.resize(L.nel * 27);
  L.face.assign(L.nel * 6, -1);
  int64_t iel = 0;
  for (int k = 0; k < 2 * nz; k += 2)
    for (int j = 0; j < 2 * ny; j += 2)
      for (int i = 0; i < 2 * nx; i += 2, iel++) {
        for (int n = 0; n < 27; n++) {
          const int* o = HexElement::xc()[n];
          L.conn[iel * 27 + n] = (int32_t)((i + o[0] + 1) + sx * ((j + o[1] + 1) + sy * (k + o[2] + 1)));
        }
        if (k == 0) L.face[iel * 6 + 4] = -2;               // bottom
        if (k == 2 * (nz - 1)) L.face[iel * 6 + 5] = -7;    // top
        if (j == 0) L.face[iel * 6 + 0] = -3;               // front
        if (j == 2 * (ny - 1)) L.face[iel * 6 + 2] = -5;    // behind
        if (i == 0) L.face[iel * 6 + 3] = -6;               // left
        if (i == 2 * (nx - 1)) L.face[iel * 6 + 1] = -4;    // right
      }
  std::vector<int32_t> part = partition ? *partition : std::vector<int32_t>((size_t)L.nel, 0);
  std::vector<int32_t> map = L.FillISvectorDofMapAllFEFamilies(part, nprocs);
  L.xyz.resize(3 * L.nnode);
  L.ijk.resize(3 * L.nnode);
  for (int64_t k = 0; k < sz; k++)
    for (int64_t j = 0; j < sy; j++)
      for (int64_t i = 0; i < sx; i++) {
        const int64_t nd = map[i + sx * (j + sy * k)];
        L.ijk[nd] = (int32_t)i;
        L.ijk[L.nnode + nd] = (int32_t)j;
        L.ijk[2 * L.nnode + nd] = (int32_t)k;
        L.xyz[nd] = (static_cast<double>(i) / static_cast<double>(2 * nx)) * (xmax - xmin) + xmin;
        L.xyz[L.nnode + nd] = (static_cast<double>(j) / static_cast<double>(2 * ny)) * (ymax - ymin) + ymin;
        L.xyz[2 * L.nnode + nd] = (static_cast<double>(k) / static_cast<double>(2 * nz)) * (zmax - zmin) + zmin;
      }
  return L;
}

namespace detail {

// open-addressing map uint64 -> int32
class PairMap {
 public:
  explicit PairMap(size_t expected) {
    size_t cap = 64;
    while (cap < expected * 2) cap <<= 1;
    keys_.assign(cap, ~0ull);
    vals_.assign(cap, -1);
    mask_ = cap - 1;
  }
  // returns the stored value, inserting `fresh` if the key is new (inserted = true)
  int32_t get_or_insert(uint64_t key, int32_t fresh, bool& inserted) {
    size_t h = (size_t)((key * 0x9E3779B97F4A7C15ull) >> 17) & mask_;
    while (true) {
      if (keys_[h] == key) { inserted = false; return vals_[h]; }
      if (keys_[h] == ~0ull) { keys_[h] = key; vals_[h] = fresh; inserted = true; return fresh; }
      h = (h + 1) & mask_;
    }
  }
 private:
  std::vector<uint64_t> keys_;
  std::vector<int32_t> vals_;
  size_t mask_;
};

// The 125 points of a parent's 5x5x5 fine lattice: for each, one (child, local node) that sits
// there, and the corner structure used to name new nodes.
struct FinePointTable {
  int child[125], node[125];
  int ncorner[125];
  int corner[125][8];      // parent local nodes at the corners of the sub-entity
  int opposite[125][8];    // index (into corner[]) of the diagonally opposite corner
  int pos_of_child_node[8][27];
  FinePointTable() {
    for (int q = 0; q < 125; q++) child[q] = -1;
    for (int j = 0; j < 8; j++) {
      // child j is the octant at parent vertex j: origin = parent lattice position of that octant
      const int* v = HexElement::xc()[j];
      const int o[3] = {v[0] + 1, v[1] + 1, v[2] + 1};     // 0 or 2 on the 5-lattice
      for (int n = 0; n < 27; n++) {
        const int* x = HexElement::xc()[n];
        const int a = o[0] + x[0] + 1, b = o[1] + x[1] + 1, c = o[2] + x[2] + 1;
        const int q = a + 5 * (b + 5 * c);
        pos_of_child_node[j][n] = q;
        if (child[q] < 0 || n < node[q]) { child[q] = j; node[q] = n; }
      }
    }
    for (int c = 0; c < 5; c++)
      for (int b = 0; b < 5; b++)
        for (int a = 0; a < 5; a++) {
          const int q = a + 5 * (b + 5 * c);
          const int p[3] = {a, b, c};
          int odd[3], nodd = 0;
          for (int d = 0; d < 3; d++) if (p[d] & 1) odd[nodd++] = d;
          ncorner[q] = 1 << nodd;
          for (int m = 0; m < (1 << nodd); m++) {
            int cp[3] = {a >> 1, b >> 1, c >> 1};
            for (int t = 0; t < nodd; t++) if (m & (1 << t)) cp[odd[t]] += 1;
            corner[q][m] = HexElement::node_at(cp[0], cp[1], cp[2]);
            opposite[q][m] = ((1 << nodd) - 1) ^ m;
          }
        }
  }
};
inline const FinePointTable& fine_points() {
  static const FinePointTable t;
  return t;
}

}  // namespace detail

// Element prolongator scattered to global ids: P of `family` from level C to its refinement F
// (LinearImplicitSystem::BuildProlongatorMatrix; rows are INSERTED, identical from every
// neighbouring coarse element, so the first visit defines the row).
inline HostCsr BuildProlongator(const MeshLevel& C, const MeshLevel& F, int family) {
  const detail::FinePointTable& T = detail::fine_points();
  const int nve = HexElement::nve(family);
  HostCsr P;
  P.nrows = F.ndofs(family);
  P.ncols = C.ndofs(family);
  // local rows
  int lidx[125][27], lcnt[125];
  double lval[125][27];
  bool used[125];
  for (int q = 0; q < 125; q++) {
    used[q] = T.node[q] < nve;
    lcnt[q] = used[q] ? HexElement::prolongator_row(family, q % 5, (q / 5) % 5, q / 25, lidx[q], lval[q]) : 0;
  }
  std::vector<int32_t> len((size_t)P.nrows, -1);
  for (int64_t E = 0; E < C.nel; E++)
    for (int q = 0; q < 125; q++)
      if (used[q]) {
        const int32_t r = F.GetSolutionDof(T.node[q], C.child_el[E * 8 + T.child[q]], family);
        if (len[r] < 0) len[r] = lcnt[q];
      }
  P.rowptr.assign(P.nrows + 1, 0);
  for (int64_t r = 0; r < P.nrows; r++) P.rowptr[r + 1] = P.rowptr[r] + (len[r] > 0 ? len[r] : 0);
  P.col.resize(P.rowptr[P.nrows]);
  P.val.resize(P.rowptr[P.nrows]);
  std::vector<char> done((size_t)P.nrows, 0);
  std::vector<std::pair<int32_t, double>> tmp(27);
  for (int64_t E = 0; E < C.nel; E++) {
    int32_t cd[27];
    for (int j = 0; j < nve; j++) cd[j] = C.GetSolutionDof(j, E, family);
    for (int q = 0; q < 125; q++)
      if (used[q]) {
        const int32_t r = F.GetSolutionDof(T.node[q], C.child_el[E * 8 + T.child[q]], family);
        if (done[r]) continue;
        done[r] = 1;
        for (int k = 0; k < lcnt[q]; k++) tmp[k] = {cd[lidx[q][k]], lval[q][k]};
        std::sort(tmp.begin(), tmp.begin() + lcnt[q]);
        for (int k = 0; k < lcnt[q]; k++) { P.col[P.rowptr[r] + k] = tmp[k].first; P.val[P.rowptr[r] + k] = tmp[k].second; }
      }
  }
  return P;
}

// MeshRefinement::RefineMesh, uniform: children 8*iel+j in coarse element order, child j = octant
// at parent vertex j, boundary faces inherited, partition inherited, then the same renumbering
// as on level 0; coordinates by biquadratic prolongation of the coarse ones (:470-472).
inline MeshLevel RefineMesh(MeshLevel& C) {
  const detail::FinePointTable& T = detail::fine_points();
  MeshLevel F;
  F.level = C.level + 1;
  F.nel = C.nel * 8;
  F.conn.resize(F.nel * 27);
  F.face.assign(F.nel * 6, -1);
  if (!C.material.empty()) { F.material.resize(F.nel); F.group.resize(F.nel); }     // children inherit (MeshRefinement.cpp:252-258)
  std::vector<int32_t> part(F.nel);
  detail::PairMap map((size_t)C.nnode * 8 + 1024);
  int32_t next = (int32_t)C.nnode;
  // lattice coordinates in temporary numbering: a coarse node doubles its coordinates, a new node
  // sits at the sum of the coordinates of the two opposite corners it is the centre of
  const bool lattice = !C.ijk.empty();
  std::vector<int32_t> tijk[3];
  if (lattice)
    for (int d = 0; d < 3; d++) {
      tijk[d].reserve((size_t)C.nnode * 8);
      tijk[d].resize(C.nnode);
      for (int64_t n = 0; n < C.nnode; n++) tijk[d][n] = 2 * C.ijk[d * C.nnode + n];
    }
  for (int64_t E = 0; E < C.nel; E++) {
    const int32_t* cn = &C.conn[E * 27];
    int32_t fid[125];
    for (int q = 0; q < 125; q++) {
      const int nc = T.ncorner[q];
      if (nc == 1) { fid[q] = cn[T.corner[q][0]]; continue; }
      int best = 0;
      for (int m = 1; m < nc; m++) if (cn[T.corner[q][m]] < cn[T.corner[q][best]]) best = m;
      const int32_t na = cn[T.corner[q][best]], nb = cn[T.corner[q][T.opposite[q][best]]];
      const uint64_t key = ((uint64_t)(uint32_t)na << 32) | (uint32_t)nb;
      bool ins;
      fid[q] = map.get_or_insert(key, next, ins);
      if (ins) {
        next++;
        if (lattice)
          for (int d = 0; d < 3; d++) tijk[d].push_back(C.ijk[d * C.nnode + na] + C.ijk[d * C.nnode + nb]);
      }
    }
    for (int j = 0; j < 8; j++) {
      const int64_t fe = E * 8 + j;
      for (int n = 0; n < 27; n++) F.conn[fe * 27 + n] = fid[T.pos_of_child_node[j][n]];
      part[fe] = C.part[E];
      if (!C.material.empty()) { F.material[fe] = C.material[E]; F.group[fe] = C.group[E]; }
      for (int f = 0; f < 6; f++) {
        const int* fn = HexElement::face_nodes()[f];
        if ((fn[0] == j || fn[1] == j || fn[2] == j || fn[3] == j) && C.face[E * 6 + f] < -1) F.face[fe * 6 + f] = C.face[E * 6 + f];
      }
    }
  }
  F.nnode = next;
  const std::vector<int32_t> nmap = F.FillISvectorDofMapAllFEFamilies(part, C.nprocs);
  if (lattice) {
    F.ijk.resize(3 * F.nnode);
    for (int d = 0; d < 3; d++)
      for (int64_t n = 0; n < F.nnode; n++) F.ijk[d * F.nnode + nmap[n]] = tijk[d][n];
  }
  // SetChildElement: (coarse element, child) -> fine element after the reorder
  C.child_el.resize(C.nel * 8);
  for (int64_t pos = 0; pos < F.nel; pos++) C.child_el[F.elem_order[pos]] = (int32_t)pos;
  // coordinates: x_f = P_biquadratic x_c, each row summed in ascending coarse-node order
  HostCsr P = BuildProlongator(C, F, BIQUADRATIC);
  F.xyz.assign(3 * F.nnode, 0.0);
  for (int d = 0; d < 3; d++)
    for (int64_t r = 0; r < P.nrows; r++) {
      double s = 0.0;
      for (int64_t k = P.rowptr[r]; k < P.rowptr[r + 1]; k++) s += P.val[k] * C.xyz[d * C.nnode + P.col[k]];
      F.xyz[d * F.nnode + r] = s;
    }
  return F;
}

// Sub-mesh of one rank: the elements [elem_offset[rank], elem_offset[rank+1]) of G with their nodes
// renumbered locally by the same first-visit rule (one "rank").  This is what one GPU holds: it is
// refined locally with RefineMesh (children inherit the parent's rank in the reference too,
// MeshMetisPartitioning.cpp:143-155), so no rank ever builds the global fine mesh.  Nodes shared
// with other ranks are found again through their lattice coordinates `ijk`.
inline MeshLevel ExtractRankSubmesh(const MeshLevel& G, int rank) {
  MeshLevel L;
  L.level = G.level;
  const int64_t e0 = G.elem_offset[rank], e1 = G.elem_offset[rank + 1];
  L.nel = e1 - e0;
  L.conn.resize(L.nel * 27);
  L.face.assign(G.face.begin() + e0 * 6, G.face.begin() + e1 * 6);
  std::vector<int32_t> g2t((size_t)G.nnode, -1), t2g;
  for (int64_t e = e0; e < e1; e++)
    for (int n = 0; n < 27; n++) {
      const int32_t g = G.conn[e * 27 + n];
      if (g2t[g] < 0) { g2t[g] = (int32_t)t2g.size(); t2g.push_back(g); }
      L.conn[(e - e0) * 27 + n] = g2t[g];
    }
  L.nnode = (int64_t)t2g.size();
  const std::vector<int32_t> part((size_t)L.nel, 0);
  const std::vector<int32_t> nmap = L.FillISvectorDofMapAllFEFamilies(part, 1);
  L.xyz.resize(3 * L.nnode);
  L.ijk.resize(3 * L.nnode);
  for (int64_t t = 0; t < L.nnode; t++)
    for (int d = 0; d < 3; d++) {
      L.xyz[d * L.nnode + nmap[t]] = G.xyz[d * G.nnode + t2g[t]];
      L.ijk[d * L.nnode + nmap[t]] = G.ijk[d * G.nnode + t2g[t]];
    }
  return L;
}

// Nodes of L on faces that are neither on the domain boundary nor shared by two elements of L:
// the interface with the sub-meshes of other ranks (sorted).  Empty for a complete mesh.
inline std::vector<int32_t> InterfaceNodes(const MeshLevel& L) {
  if (!L.etype.empty()) return {};      // hexahedral slabs only: the sharded run partitions generated boxes
  std::vector<uint8_t> cnt((size_t)L.nnode, 0), mark((size_t)L.nnode, 0);
  for (int64_t e = 0; e < L.nel; e++)
    for (int f = 0; f < 6; f++) cnt[L.conn[e * 27 + 20 + f]]++;
  for (int64_t e = 0; e < L.nel; e++)
    for (int f = 0; f < 6; f++)
      if (L.face[e * 6 + f] == -1 && cnt[L.conn[e * 27 + 20 + f]] == 1)
        for (int iv = 0; iv < 9; iv++) mark[L.conn[e * 27 + HexElement::face_nodes()[f][iv]]] = 1;
  std::vector<int32_t> out;
  for (int64_t n = 0; n < L.nnode; n++) if (mark[n]) out.push_back((int32_t)n);
  return out;
}

// Per-coarse-element maps of the element-gather Galerkin product (device kernel b2_galerkin.cu):
// the fine dofs of every coarse element in lattice order, the entity code of every fine lattice
// point, the valence (number of elements in [e0,e1)) of the 27 sub-entities of every element, and
// the dense element prolongator P_loc[nf][nc] (ElemType.cpp:439-532).  Fine lattice: 5 points per
// direction for the biquadratic family (125 fine dofs), the 3 even ones for the linear family (27).
struct GalerkinElement {
  int nf = 0, nc = 0;
  std::vector<int> q;                 // lattice point (0..124) of local fine dof a
  std::vector<uint8_t> entity;        // [nf] 3-trit code: per direction 0 = low face, 1 = interior, 2 = high face
  std::vector<double> ploc;           // [nf][nc]
};
inline GalerkinElement BuildGalerkinElement(int family) {
  const detail::FinePointTable& T = detail::fine_points();
  GalerkinElement g;
  g.nc = HexElement::nve(family);
  for (int q = 0; q < 125; q++)
    if (T.node[q] < g.nc) g.q.push_back(q);
  g.nf = (int)g.q.size();
  g.entity.resize(g.nf);
  g.ploc.assign((size_t)g.nf * g.nc, 0.0);
  for (int a = 0; a < g.nf; a++) {
    const int q = g.q[a], p[3] = {q % 5, (q / 5) % 5, q / 25};
    int code = 0, mul = 1;
    for (int d = 0; d < 3; d++) { code += (p[d] == 0 ? 0 : (p[d] == 4 ? 2 : 1)) * mul; mul *= 3; }
    g.entity[a] = (uint8_t)code;
    int idx[27];
    double val[27];
    const int n = HexElement::prolongator_row(family, p[0], p[1], p[2], idx, val);
    for (int k = 0; k < n; k++) g.ploc[(size_t)a * g.nc + idx[k]] = val[k];
  }
  return g;
}
// fine_dofs[(e1-e0)][nf], valence[(e1-e0)][27] for the coarse elements [e0, e1) of C (F = its refinement)
inline void BuildGalerkinMaps(const MeshLevel& C, const MeshLevel& F, int family, int64_t e0, int64_t e1,
                              int32_t* fine_dofs, uint8_t* valence) {
  const detail::FinePointTable& T = detail::fine_points();
  const GalerkinElement g = BuildGalerkinElement(family);
  std::vector<uint8_t> count((size_t)C.nnode, 0);
  for (int64_t E = e0; E < e1; E++)
    for (int n = 0; n < 27; n++) count[C.conn[E * 27 + n]]++;
  for (int64_t E = e0; E < e1; E++) {
    for (int a = 0; a < g.nf; a++) {
      const int q = g.q[a];
      fine_dofs[(E - e0) * g.nf + a] = F.GetSolutionDof(T.node[q], C.child_el[E * 8 + T.child[q]], family);
    }
    for (int code = 0; code < 27; code++)
      valence[(E - e0) * 27 + code] = count[C.conn[E * 27 + HexElement::node_at(code % 3, (code / 3) % 3, code / 9)]];
  }
}

}  // namespace femus_b200
