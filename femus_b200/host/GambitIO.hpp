// Gambit neutral-file (.neu) reader for HEX27 meshes: the coarse-mesh input of the reference's shipped
// 3-D Poisson cases (applications/001_Poisson/input/cube_Hex.neu, input3D_Hex_*.json).  Follows
// GambitIO::read (src/06_mesh/00_single_level/01_input/01_from_external_file/GambitIO.cpp:92-352):
// section order CONTROL INFO / NODAL COORDINATES / ELEMENTS/CELLS / ELEMENT GROUP / BOUNDARY CONDITIONS,
// local nodes permuted by GambitToFemusVertexIndex (:56-61), faces by GambitToFemusFaceIndex (:83),
// boundary flag = -(set name) - 1 (:330), coordinates divided by Lref (:262-264); then the
// reference renumbers nodes by first visit exactly as for a generated box (Mesh.cpp:517-559).
// Only 27-node hexahedra and a single element group are accepted (tets, wedges and the
// material/group reordering of Mesh.cpp:621-702 are the next step); anything else aborts.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <string>
#include "BoxMesh.hpp"

namespace femus_b200 {

inline MeshLevel ReadGambitHex27(const char* path, double Lref = 1.0) {
  static const int vertex_map[27] = {4, 16, 0, 15, 23, 11, 7, 19, 3, 12, 20, 8, 25, 26, 24, 14, 22, 10, 5, 17, 1, 13, 21, 9, 6, 18, 2};
  static const int face_map[6] = {0, 4, 2, 5, 3, 1};
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
  L.conn.resize((size_t)nel * 27);
  L.face.assign((size_t)nel * 6, -1);
  seek("ELEMENTS/CELLS");
  in >> tok;                                    // version
  for (long iel = 0; iel < nel; iel++) {
    long id, type, nve;
    in >> id >> type >> nve;
    if (nve != 27) fail("only 27-node hexahedra are supported by the B200 backend so far");
    for (int i = 0; i < 27; i++) {
      long v;
      in >> v;
      if (v < 1 || v > nvt) fail("node id out of range");
      L.conn[(size_t)iel * 27 + vertex_map[i]] = (int32_t)(v - 1);
    }
  }
  in >> tok;
  if (tok != "ENDOFSECTION") fail("bad element section");
  if (ngroup != 1) fail("more than one element group: the material/group element reordering is not implemented");
  seek("GROUP:");
  seek("ENDOFSECTION");
  for (long k = 0; k < nbcd; k++) {
    seek("CONDITIONS");
    in >> tok;                                  // version
    long value, itype, nface, d0, d1;
    in >> value >> itype >> nface >> d0 >> d1;
    const int32_t flag = (int32_t)(-value - 1);
    for (long i = 0; i < nface; i++) {
      long iel, etype, iface;
      in >> iel >> etype >> iface;
      if (iel < 1 || iel > nel || iface < 1 || iface > 6) fail("boundary face out of range");
      L.face[(size_t)(iel - 1) * 6 + face_map[iface - 1]] = flag;
    }
    in >> tok;
    if (tok != "ENDOFSECTION") fail("bad boundary section");
  }
  // node renumbering by first visit (serial: one rank), coordinates follow
  const std::vector<int32_t> part((size_t)nel, 0);
  const std::vector<int32_t> map = L.FillISvectorDofMapAllFEFamilies(part, 1);
  L.xyz.resize((size_t)3 * nvt);
  for (long j = 0; j < nvt; j++)
    for (int d = 0; d < 3; d++) L.xyz[(size_t)d * nvt + map[j]] = xyz_file[(size_t)d * nvt + j];
  return L;
}

}  // namespace femus_b200
