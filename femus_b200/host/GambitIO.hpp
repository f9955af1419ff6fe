// Gambit neutral-file (.neu) reader for meshes of 27-node hexahedra, 10-node tetrahedra and 18-node wedges,
// also mixed (cube_all_shapes_Six_boundary_groups.neu): the coarse-mesh
// input of the reference's shipped 3-D Poisson cases (applications/001_Poisson/input/cube_Hex.neu,
// cube_Tet.neu, input3D_*.json).  Follows GambitIO::read (src/06_mesh/00_single_level/01_input/
// 01_from_external_file/GambitIO.cpp:92-352): section order CONTROL INFO / NODAL COORDINATES /
// ELEMENTS/CELLS / ELEMENT GROUP / BOUNDARY CONDITIONS, local nodes permuted by GambitToFemusVertexIndex
// (:56-67), faces by GambitToFemusFaceIndex (:83-85), boundary flag = -(set name) - 1 (:330), coordinates
// divided by Lref (:262-264); then Mesh::AddBiquadraticNodesNotInMeshFile (Mesh.cpp:1207-1333) creates the
// face and centre nodes a 10-node tetrahedron lacks, and the reference renumbers nodes by first visit exactly
// as for a generated box (Mesh.cpp:517-559).
// Element groups set the material / group of their elements, by which the reference orders the elements of
// a rank (Mesh.cpp:621-702).  Anything the reference would reject aborts.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <map>
#include <string>
#include "BoxMesh.hpp"

namespace femus_b200 {

// Mesh::AddBiquadraticNodesNotInMeshFile for tetrahedra and wedges: one node per TRIANGULAR face (shared by the
// two elements, of either type, that have the same three vertices), one per element; coordinates from the
// element's file nodes with the barycentric weights of Mesh.cpp:105-125 (-1/9 on the three vertices, 4/9 on the
// three edge nodes of the triangle; tetrahedron centre -1/8, 1/4; wedge centre from its mid-height triangle
// 12-14 / 15-17), every element overwriting what an earlier one wrote, in element order, as the reference.
inline void AddBiquadraticNodesNotInMeshFile(MeshLevel& L, std::vector<double>& xyz_file) {
  if (L.etype.empty()) return;
  struct Key { uint32_t a, b, c; bool operator<(const Key& o) const { return a != o.a ? a < o.a : (b != o.b ? b < o.b : c < o.c); } };
  std::map<Key, int32_t> face_node;
  int64_t nn = L.nnode;
  for (int64_t e = 0; e < L.nel; e++) {
    const int t = L.etype[e];
    if (t == HEX) continue;
    for (int f = 0; f < ElemTopology::nfaces(t); f++) {
      if (ElemTopology::face_nvert(t, f) != 3) continue;
      uint32_t v[3];
      for (int k = 0; k < 3; k++) v[k] = (uint32_t)L.conn[e * 27 + ElemTopology::face_node(t, f, k)];
      std::sort(v, v + 3);
      auto it = face_node.find(Key{v[0], v[1], v[2]});
      if (it == face_node.end()) it = face_node.emplace(Key{v[0], v[1], v[2]}, (int32_t)nn++).first;
      L.conn[e * 27 + ElemTopology::nve(t, SERENDIPITY) + f] = it->second;
    }
  }
  for (int64_t e = 0; e < L.nel; e++)
    if (L.etype[e] != HEX) L.conn[e * 27 + ElemTopology::nve(L.etype[e], BIQUADRATIC) - 1] = (int32_t)nn++;
  const int64_t n0 = L.nnode;
  std::vector<double> x2((size_t)3 * nn);
  for (int d = 0; d < 3; d++) std::copy(xyz_file.begin() + d * n0, xyz_file.begin() + (d + 1) * n0, x2.begin() + d * nn);
  for (int64_t e = 0; e < L.nel; e++) {
    const int t = L.etype[e];
    if (t == HEX) continue;
    const int ndof = ElemTopology::nve(t, BIQUADRATIC), jstart = t == TET ? 10 : 18;
    for (int j = jstart; j < ndof; j++) {
      double w[18];
      for (int i = 0; i < jstart; i++) w[i] = 0.;
      if (j < ndof - 1) {           // centre of the triangular face that owns node j
        const int f = j - ElemTopology::nve(t, SERENDIPITY);
        for (int k = 0; k < 3; k++) { w[ElemTopology::face_node(t, f, k)] = -1. / 9.; w[ElemTopology::face_node(t, f, 3 + k)] = 4. / 9.; }
      } else if (t == TET) {
        for (int i = 0; i < 4; i++) w[i] = -1. / 8.;
        for (int i = 4; i < 10; i++) w[i] = 1. / 4.;
      } else {
        for (int i = 12; i < 15; i++) w[i] = -1. / 9.;
        for (int i = 15; i < 18; i++) w[i] = 4. / 9.;
      }
      const int32_t jn = L.conn[e * 27 + j];
      for (int d = 0; d < 3; d++) {
        double s = 0.;
        for (int i = 0; i < jstart; i++) s += x2[(size_t)d * nn + L.conn[e * 27 + i]] * w[i];
        x2[(size_t)d * nn + jn] = s;
      }
    }
  }
  L.nnode = nn;
  xyz_file.swap(x2);
}

inline MeshLevel ReadGambit(const char* path, double Lref = 1.0) {
  static const int vertex_map_hex[27] = {4, 16, 0, 15, 23, 11, 7, 19, 3, 12, 20, 8, 25, 26, 24, 14, 22, 10, 5, 17, 1, 13, 21, 9, 6, 18, 2};
  static const int vertex_map_tet[10] = {0, 4, 1, 6, 5, 2, 7, 8, 9, 3};
  static const int vertex_map_wedge[18] = {3, 11, 5, 9, 10, 4, 12, 17, 14, 15, 16, 13, 0, 8, 2, 6, 7, 1};
  static const int face_map_hex[6] = {0, 4, 2, 5, 3, 1};
  static const int face_map_tet[4] = {0, 1, 2, 3};
  static const int face_map_wedge[5] = {2, 1, 0, 4, 3};
  auto fail = [&](const char* what) {
    std::fprintf(stderr, "femus_b200: Gambit file %s: %s\n", path, what);
    std::abort();
  };
  std::ifstream in(path);
  if (!in) fail("cannot open");
  std::string tok;
  auto seek = [&](const char* word) {
    while (in >> tok)
      if (tok == word) return;
    fail("unexpected end of file");
  };
  long nvt = 0, nel = 0, ngroup = 0, nbcd = 0, dim = 0, dimNodes = 0;
  seek("NDFVL");
  in >> nvt >> nel >> ngroup >> nbcd >> dim >> dimNodes;
  in >> tok;
  if (tok != "ENDOFSECTION" || dim != 3 || dimNodes != 3 || nvt <= 0 || nel <= 0) fail("bad control section (3-D meshes only)");
  // the reference re-opens the file for every section; the sections come in this order in Gambit files
  std::vector<double> xyz_file((size_t)3 * nvt);
  seek("COORDINATES");
  in >> tok;                                    // version
  for (long j = 0; j < nvt; j++) {
    double x, y, z;
    in >> tok >> x >> y >> z;
    xyz_file[j] = x / Lref;
    xyz_file[nvt + j] = y / Lref;
    xyz_file[2 * nvt + j] = z / Lref;
  }
  in >> tok;
  if (tok != "ENDOFSECTION") fail("bad node section");
  MeshLevel L;
  L.level = 0;
  L.nel = nel;
  L.nnode = nvt;
  L.conn.assign((size_t)nel * 27, -1);
  L.face.assign((size_t)nel * 6, -1);
  std::vector<uint8_t> etype((size_t)nel, (uint8_t)HEX);
  seek("ELEMENTS/CELLS");
  in >> tok;                                    // version
  for (long iel = 0; iel < nel; iel++) {
    long id, type, nve;
    in >> id >> type >> nve;
    if (nve != 27 && nve != 10 && nve != 18) fail("element is not a 27-node hexahedron, 10-node tetrahedron or 18-node wedge (use a second-order mesh)");
    etype[iel] = nve == 27 ? (uint8_t)HEX : (nve == 10 ? (uint8_t)TET : (uint8_t)WEDGE);
    const int* vmap = nve == 27 ? vertex_map_hex : (nve == 10 ? vertex_map_tet : vertex_map_wedge);
    for (int i = 0; i < nve; i++) {
      long v;
      in >> v;
      if (v < 1 || v > nvt) fail("node id out of range");
      L.conn[(size_t)iel * 27 + vmap[i]] = (int32_t)(v - 1);
    }
  }
  bool all_hex = true;
  for (long iel = 0; iel < nel; iel++) all_hex = all_hex && etype[iel] == HEX;
  if (!all_hex) L.etype = etype;
  in >> tok;
  if (tok != "ENDOFSECTION") fail("bad element section");
  // ELEMENT GROUP sections (GambitIO.cpp:290-313): "GROUP: id ELEMENTS: n MATERIAL: mat NFLAGS: k", the group
  // NAME (an integer in FEMuS meshes), the flags line, then the n element ids; elements start in group 1
  if (ngroup < 1) fail("no element group");
  std::vector<int16_t> material((size_t)nel, 0), group((size_t)nel, 1);
  for (long k = 0; k < ngroup; k++) {
    seek("GROUP:");
    long ngel = 0, gr_mat = 0, gr_name = 0;
    in >> tok >> tok >> ngel >> tok >> gr_mat >> tok >> tok >> gr_name >> tok;
    if (!in || ngel < 0) fail("bad group header (the group name must be an integer)");
    for (long i = 0; i < ngel; i++) {
      long iel;
      in >> iel;
      if (iel < 1 || iel > nel) fail("group element out of range");
      group[iel - 1] = (int16_t)gr_name;
      material[iel - 1] = (int16_t)gr_mat;
    }
    in >> tok;
    if (tok != "ENDOFSECTION") fail("bad group section");
  }
  L.material = material;
  L.group = group;
  for (long k = 0; k < nbcd; k++) {
    seek("CONDITIONS");
    in >> tok;                                  // version
    long value, itype, nface, d0, d1;
    in >> value >> itype >> nface >> d0 >> d1;
    const int32_t flag = (int32_t)(-value - 1);
    for (long i = 0; i < nface; i++) {
      long iel, file_type, iface;
      in >> iel >> file_type >> iface;
      if (iel < 1 || iel > nel || iface < 1 || iface > ElemTopology::nfaces(etype[iel - 1])) fail("boundary face out of range");
      const int* fmap = etype[iel - 1] == HEX ? face_map_hex : (etype[iel - 1] == TET ? face_map_tet : face_map_wedge);
      L.face[(size_t)(iel - 1) * 6 + fmap[iface - 1]] = flag;
    }
    in >> tok;
    if (tok != "ENDOFSECTION") fail("bad boundary section");
  }
  AddBiquadraticNodesNotInMeshFile(L, xyz_file);
  nvt = (long)L.nnode;
  // node renumbering by first visit (serial: one rank), coordinates follow
  const std::vector<int32_t> part((size_t)nel, 0);
  const std::vector<int32_t> map = L.FillISvectorDofMapAllFEFamilies(part, 1);
  L.xyz.resize((size_t)3 * nvt);
  for (long j = 0; j < nvt; j++)
    for (int d = 0; d < 3; d++) L.xyz[(size_t)d * nvt + map[j]] = xyz_file[(size_t)d * nvt + j];
  return L;
}

}  // namespace femus_b200
