// C entry points of the HOST layer (mesh hierarchy, dof maps, Dirichlet flags, prolongators, FE
// tables) so that the Python harness can drive and inspect it.  Pure CPU code: no CUDA calls.
#include <cstring>
#include <memory>
#include <string>
#include "../host/AsmPartition.hpp"
#include "../host/SystemLayout.hpp"
#include "../host/BoxMesh.hpp"
#include "../host/FaceElement.hpp"
#include "../host/GambitIO.hpp"
#include "../host/GeneralMesh.hpp"
#include "../../include/femus_b200_host.h"

using namespace femus_b200;

struct b2h_hier {
  std::vector<MeshLevel> levels;
};
struct b2h_csr {
  HostCsr m;
};
struct b2h_asm {
  AsmIndex ix;
};
static thread_local std::string g_b2h_error;

extern "C" {

b2h_hier* b2h_hier_create(int nx, int ny, int nz, int nlevels, const double* bounds6, int nprocs) {
  if (nx < 1 || ny < 1 || nz < 1 || nlevels < 1 || nprocs < 1) return nullptr;
  b2h_hier* h = new b2h_hier();
  const double unit[6] = {0., 1., 0., 1., 0., 1.};
  const double* b = bounds6 ? bounds6 : unit;
  std::vector<int32_t> part;
  if (nprocs > 1) part = SlabPartition(nx, ny, nz, nprocs);
  h->levels.reserve(nlevels);
  h->levels.push_back(GenerateCoarseBoxMesh(nx, ny, nz, b[0], b[1], b[2], b[3], b[4], b[5], nprocs > 1 ? &part : nullptr, nprocs));
  for (int l = 1; l < nlevels; l++) h->levels.push_back(RefineMesh(h->levels.back()));
  return h;
}
b2h_hier* b2h_hier_create_local(int nx, int ny, int nz, int nlevels, const double* bounds6, int nprocs, int rank) {
  if (nx < 1 || ny < 1 || nz < 1 || nlevels < 1 || nprocs < 1 || rank < 0 || rank >= nprocs) return nullptr;
  const double unit[6] = {0., 1., 0., 1., 0., 1.};
  const double* b = bounds6 ? bounds6 : unit;
  std::vector<int32_t> part = SlabPartition(nx, ny, nz, nprocs);
  MeshLevel G = GenerateCoarseBoxMesh(nx, ny, nz, b[0], b[1], b[2], b[3], b[4], b[5], &part, nprocs);
  if (G.elem_offset[rank + 1] == G.elem_offset[rank]) return nullptr;      // more ranks than z-layers
  b2h_hier* h = new b2h_hier();
  h->levels.reserve(nlevels);
  h->levels.push_back(ExtractRankSubmesh(G, rank));
  for (int l = 1; l < nlevels; l++) h->levels.push_back(RefineMesh(h->levels.back()));
  return h;
}
b2h_hier* b2h_hier_create_from_neu(const char* path, int nlevels, double Lref) {
  if (!path || nlevels < 1 || Lref <= 0.) return nullptr;
  b2h_hier* h = new b2h_hier();
  h->levels.reserve(nlevels);
  h->levels.push_back(ReadGambit(path, Lref));
  for (int l = 1; l < nlevels; l++) h->levels.push_back(RefineAnyMesh(h->levels.back()));
  return h;
}
void b2h_hier_destroy(b2h_hier* h) { delete h; }
const int32_t* b2h_level_ijk(const b2h_hier* h, int l) { return h->levels[l].ijk.empty() ? nullptr : h->levels[l].ijk.data(); }
int64_t b2h_level_interface_nodes(const b2h_hier* h, int l, int32_t* out) {
  const std::vector<int32_t> v = InterfaceNodes(h->levels[l]);
  if (out) std::copy(v.begin(), v.end(), out);
  return (int64_t)v.size();
}
int b2h_hier_nlevels(const b2h_hier* h) { return (int)h->levels.size(); }
int b2h_hier_nprocs(const b2h_hier* h) { return h->levels[0].nprocs; }
int64_t b2h_level_nel(const b2h_hier* h, int l) { return h->levels[l].nel; }
int64_t b2h_level_nnode(const b2h_hier* h, int l) { return h->levels[l].nnode; }
const int32_t* b2h_level_conn(const b2h_hier* h, int l) { return h->levels[l].conn.data(); }
const int32_t* b2h_level_face(const b2h_hier* h, int l) { return h->levels[l].face.data(); }
const int32_t* b2h_level_part(const b2h_hier* h, int l) { return h->levels[l].part.data(); }
const double* b2h_level_xyz(const b2h_hier* h, int l) { return h->levels[l].xyz.data(); }
const int32_t* b2h_level_child_el(const b2h_hier* h, int l) {
  return h->levels[l].child_el.empty() ? nullptr : h->levels[l].child_el.data();
}
void b2h_level_offsets(const b2h_hier* h, int l, int64_t* elem_offset, int64_t* dof_offset3) {
  const MeshLevel& L = h->levels[l];
  const int np1 = L.nprocs + 1;
  std::copy(L.elem_offset.begin(), L.elem_offset.end(), elem_offset);
  for (int k = 0; k < 3; k++) std::copy(L.dof_offset[k].begin(), L.dof_offset[k].end(), dof_offset3 + k * np1);
}
int64_t b2h_level_ndofs(const b2h_hier* h, int l, int family) { return h->levels[l].ndofs(family); }
void b2h_level_system_dofs(const b2h_hier* h, int l, int family, int32_t* out) {
  std::vector<int32_t> d = h->levels[l].system_dofs(family);
  std::copy(d.begin(), d.end(), out);
}
void b2h_level_bdc(const b2h_hier* h, int l, int family, const int* dirichlet_faces7, double* out) {
  bool f[7];
  for (int i = 0; i < 7; i++) f[i] = dirichlet_faces7[i] != 0;
  std::vector<double> b = h->levels[l].GenerateBdc(family, f);
  std::copy(b.begin(), b.end(), out);
}

b2h_csr* b2h_prolongator_create(const b2h_hier* h, int lfine, int family) {
  if (lfine < 1 || lfine >= (int)h->levels.size()) return nullptr;
  b2h_csr* p = new b2h_csr();
  p->m = BuildAnyProlongator(h->levels[lfine - 1], h->levels[lfine], family);
  return p;
}
void b2h_csr_destroy(b2h_csr* p) { delete p; }
int64_t b2h_csr_nrows(const b2h_csr* p) { return p->m.nrows; }
int64_t b2h_csr_ncols(const b2h_csr* p) { return p->m.ncols; }
int64_t b2h_csr_nnz(const b2h_csr* p) { return (int64_t)p->m.col.size(); }
const int64_t* b2h_csr_rowptr(const b2h_csr* p) { return p->m.rowptr.data(); }
const int32_t* b2h_csr_col(const b2h_csr* p) { return p->m.col.data(); }
const double* b2h_csr_val(const b2h_csr* p) { return p->m.val.data(); }

int b2h_galerkin_nf(int family) { return BuildGalerkinElement(family).nf; }
void b2h_galerkin_element(int family, double* ploc, uint8_t* fine_entity) {
  const GalerkinElement g = BuildGalerkinElement(family);
  std::copy(g.ploc.begin(), g.ploc.end(), ploc);
  std::copy(g.entity.begin(), g.entity.end(), fine_entity);
}
int b2h_galerkin_maps(const b2h_hier* h, int lcoarse, int family, int64_t e0, int64_t e1, int32_t* fine_dofs,
                      uint8_t* valence) {
  if (lcoarse < 0 || lcoarse + 1 >= (int)h->levels.size()) return 1;
  const MeshLevel& C = h->levels[lcoarse];
  if (e0 < 0 || e1 > C.nel || e0 > e1) return 1;
  if (C.uniform_type() != HEX) return 1;       // the element-gather Galerkin product is for hexahedra
  BuildGalerkinMaps(C, h->levels[lcoarse + 1], family, e0, e1, fine_dofs, valence);
  return 0;
}

int b2h_level_elem_type(const b2h_hier* h, int l) { return h->levels[l].uniform_type(); }
int b2h_elem_nve(int type, int family) { return ElemTopology::nve(type, family); }
int b2h_elem_ngauss(int type) { return ElemTopology::ngauss(type); }
void b2h_elem_tables(int type, int family, double* phi, double* dxi, double* deta, double* dzeta, double* w) {
  HexElement::Tables t = ElemTopology::tables(type, family);
  std::copy(t.phi.begin(), t.phi.end(), phi);
  std::copy(t.dxi.begin(), t.dxi.end(), dxi);
  std::copy(t.deta.begin(), t.deta.end(), deta);
  std::copy(t.dzeta.begin(), t.dzeta.end(), dzeta);
  std::copy(t.w.begin(), t.w.end(), w);
}
int b2h_elem_prolongator_row(int type, int family, int child, int node, int* idx, double* val) {
  return ElemTopology::prolongator_row(type, family, child, node, idx, val);
}
int b2h_elem_child_face(int type, int child, int child_face) { return detail::child_faces().parent_face[type][child][child_face]; }
void b2h_level_elem_types(const b2h_hier* h, int l, uint8_t* out) {
  const MeshLevel& L = h->levels[l];
  for (int64_t e = 0; e < L.nel; e++) out[e] = (uint8_t)L.type_of(e);
}
void b2h_level_system_dofs27(const b2h_hier* h, int l, int family, int32_t* out) {
  std::vector<int32_t> d = h->levels[l].system_dofs27(family);
  std::copy(d.begin(), d.end(), out);
}
b2h_csr* b2h_sparsity_create(const b2h_hier* h, int l, int family) {
  if (l < 0 || l >= (int)h->levels.size()) return nullptr;
  b2h_csr* p = new b2h_csr();
  p->m = BuildSparsity(h->levels[l], family);
  return p;
}
/* test hook: the general (any element type) refinement and prolongator on a hexahedral box hierarchy */
b2h_hier* b2h_hier_create_general(int nx, int ny, int nz, int nlevels) {
  if (nx < 1 || ny < 1 || nz < 1 || nlevels < 1) return nullptr;
  b2h_hier* h = new b2h_hier();
  h->levels.reserve(nlevels);
  h->levels.push_back(GenerateCoarseBoxMesh(nx, ny, nz, 0., 1., 0., 1., 0., 1., nullptr, 1));
  h->levels.back().etype.assign((size_t)h->levels.back().nel, (uint8_t)HEX);
  for (int l = 1; l < nlevels; l++) h->levels.push_back(RefineGeneralMesh(h->levels.back()));
  return h;
}

int b2h_hex_nve(int family) { return HexElement::nve(family); }
void b2h_hex_tables(int family, double* phi, double* dxi, double* deta, double* dzeta, double* w) {
  HexElement::Tables t = HexElement::tables(family);
  std::copy(t.phi.begin(), t.phi.end(), phi);
  std::copy(t.dxi.begin(), t.dxi.end(), dxi);
  std::copy(t.deta.begin(), t.deta.end(), deta);
  std::copy(t.dzeta.begin(), t.dzeta.end(), dzeta);
  std::copy(t.w.begin(), t.w.end(), w);
}
int b2h_face_nvf(int family) { return HexElement::face_ndofs(family); }
void b2h_face_tables(int family, double* phi, double* dxi, double* deta, double* w) {
  HexElement::FaceTables t = HexElement::face_tables(family);
  std::copy(t.phi.begin(), t.phi.end(), phi);
  std::copy(t.dxi.begin(), t.dxi.end(), dxi);
  std::copy(t.deta.begin(), t.deta.end(), deta);
  std::copy(t.w.begin(), t.w.end(), w);
}
void b2h_hex_face_nodes(int32_t* out) {
  for (int f = 0; f < 6; f++)
    for (int i = 0; i < 9; i++) out[f * 9 + i] = HexElement::face_nodes()[f][i];
}
int64_t b2h_level_boundary_faces(const b2h_hier* h, int l, int32_t* elem, int32_t* face, int32_t* bidx) {
  const MeshLevel& L = h->levels[l];
  int64_t n = 0;
  for (int64_t e = 0; e < L.nel; e++)
    for (int f = 0; f < 6; f++)
      if (L.face[e * 6 + f] < -1) {
        if (elem) { elem[n] = (int32_t)e; face[n] = f; bidx[n] = -(L.face[e * 6 + f] + 1); }
        n++;
      }
  return n;
}
int b2h_face_kind_ngauss(int kind) { return FaceElement::ngauss(kind); }
int b2h_face_kind_ndofs(int kind, int family) { return FaceElement::ndofs(kind, family); }
void b2h_face_kind_tables(int kind, int family, double* phi, double* dxi, double* deta, double* w) {
  FaceElement::Tables t = FaceElement::tables(kind, family);
  std::copy(t.phi.begin(), t.phi.end(), phi);
  std::copy(t.dxi.begin(), t.dxi.end(), dxi);
  std::copy(t.deta.begin(), t.deta.end(), deta);
  std::copy(t.w.begin(), t.w.end(), w);
}
void b2h_elem_face_nodes(int type, int32_t* out) {
  for (int f = 0; f < 6; f++)
    for (int i = 0; i < 9; i++)
      out[f * 9 + i] = (f < ElemTopology::nfaces(type) && i < ElemTopology::face_ndofs(type, f, BIQUADRATIC)) ? ElemTopology::face_node(type, f, i) : -1;
}
int b2h_elem_face_kind(int type, int f) {
  return f < ElemTopology::nfaces(type) ? FaceElement::kind_of_nvert(ElemTopology::face_nvert(type, f)) : -1;
}
b2h_asm* b2h_asm_create(const b2h_hier* h, int l, int family, int block_elems, int iproc) {
  try {
    if (!h || l < 0 || l >= (int)h->levels.size() || family < 0 || family > 2 || block_elems < 1)
      throw std::invalid_argument("b2h_asm_create: bad level, family or block size");
    std::unique_ptr<b2h_asm> a(new b2h_asm());
    a->ix = BuildAsmIndex(h->levels[l], family, (unsigned)block_elems, iproc);
    return a.release();
  } catch (const std::exception& e) {
    g_b2h_error = e.what();
    return nullptr;
  }
}
b2h_asm* b2h_asm_create_system(const b2h_hier* h, int l, int nvars, const int* families, int nschur, int block_elems, int iproc) {
  try {
    if (!h || l < 0 || l >= (int)h->levels.size() || nvars < 1 || !families || block_elems < 1)
      throw std::invalid_argument("b2h_asm_create_system: bad level, variables or block size");
    std::unique_ptr<b2h_asm> a(new b2h_asm());
    const SystemLayout sys(h->levels[l], std::vector<int>(families, families + nvars));
    a->ix = BuildAsmIndexSystem(h->levels[l], sys, nschur, (unsigned)block_elems, iproc);
    return a.release();
  } catch (const std::exception& e) {
    g_b2h_error = e.what();
    return nullptr;
  }
}
int b2h_system_offsets(const b2h_hier* h, int l, int nvars, const int* families, int64_t* out) {
  try {
    if (!h || l < 0 || l >= (int)h->levels.size() || nvars < 1 || !families || !out) throw std::invalid_argument("b2h_system_offsets: bad arguments");
    const SystemLayout sys(h->levels[l], std::vector<int>(families, families + nvars));
    const int np = h->levels[l].nprocs;
    for (int k = 0; k <= nvars; k++) std::copy(sys.KKoffset[k].begin(), sys.KKoffset[k].end(), out + (size_t)k * np);
    return 0;
  } catch (const std::exception& e) {
    g_b2h_error = e.what();
    return 1;
  }
}
void b2h_system_elem_dofs(const b2h_hier* h, int l, int nvars, const int* families, int32_t* out) {
  const SystemLayout sys(h->levels[l], std::vector<int>(families, families + nvars));
  const std::vector<int32_t> d = SystemElementDofs(h->levels[l], sys);
  std::copy(d.begin(), d.end(), out);
}
b2h_csr* b2h_system_sparsity_create(const b2h_hier* h, int l, int nvars, const int* families, const uint8_t* pattern) {
  try {
    if (!h || l < 0 || l >= (int)h->levels.size() || nvars < 1 || !families) throw std::invalid_argument("b2h_system_sparsity_create: bad arguments");
    const SystemLayout sys(h->levels[l], std::vector<int>(families, families + nvars));
    std::unique_ptr<b2h_csr> p(new b2h_csr());
    p->m = BuildSystemSparsity(h->levels[l], sys, pattern);
    return p.release();
  } catch (const std::exception& e) {
    g_b2h_error = e.what();
    return nullptr;
  }
}
b2h_csr* b2h_system_prolongator_create(const b2h_hier* h, int lfine, int nvars, const int* families) {
  try {
    if (!h || lfine < 1 || lfine >= (int)h->levels.size() || nvars < 1 || !families) throw std::invalid_argument("b2h_system_prolongator_create: bad arguments");
    std::unique_ptr<b2h_csr> p(new b2h_csr());
    p->m = BuildSystemProlongator(h->levels[lfine - 1], h->levels[lfine], std::vector<int>(families, families + nvars));
    return p.release();
  } catch (const std::exception& e) {
    g_b2h_error = e.what();
    return nullptr;
  }
}
int b2h_system_bdc(const b2h_hier* h, int l, int nvars, const int* families, const uint8_t* dirichlet, double* out) {
  try {
    if (!h || l < 0 || l >= (int)h->levels.size() || nvars < 1 || !families || !dirichlet || !out) throw std::invalid_argument("b2h_system_bdc: bad arguments");
    const SystemLayout sys(h->levels[l], std::vector<int>(families, families + nvars));
    const std::vector<double> b = SystemBdc(h->levels[l], sys, dirichlet);
    std::copy(b.begin(), b.end(), out);
    return 0;
  } catch (const std::exception& e) {
    g_b2h_error = e.what();
    return 1;
  }
}
void b2h_asm_destroy(b2h_asm* a) { delete a; }
int64_t b2h_asm_nblocks(const b2h_asm* a) { return a->ix.nblocks(); }
void b2h_asm_block_type_range(const b2h_asm* a, int64_t* out3) { std::copy(a->ix.block_type_range, a->ix.block_type_range + 3, out3); }
const int64_t* b2h_asm_elem_ptr(const b2h_asm* a) { return a->ix.elem_ptr.data(); }
const int32_t* b2h_asm_elems(const b2h_asm* a) { return a->ix.elems.data(); }
const int64_t* b2h_asm_local_ptr(const b2h_asm* a) { return a->ix.local_ptr.data(); }
const int32_t* b2h_asm_local(const b2h_asm* a) { return a->ix.local.data(); }
const int64_t* b2h_asm_overlap_ptr(const b2h_asm* a) { return a->ix.overlap_ptr.data(); }
const int32_t* b2h_asm_overlap(const b2h_asm* a) { return a->ix.overlap.data(); }
int64_t b2h_asm_schedule(int64_t n, const int64_t* rowptr, const int32_t* col, int64_t nblocks, const int64_t* blk_ptr,
                         const int32_t* blk_dofs, int mode, int32_t* group_of_block) {
  try {
    if (n < 1 || nblocks < 1 || !rowptr || !col || !blk_ptr || !blk_dofs || !group_of_block)
      throw std::invalid_argument("b2h_asm_schedule: null or empty argument");
    return AsmSchedule(n, rowptr, col, nblocks, blk_ptr, blk_dofs, mode, group_of_block);
  } catch (const std::exception& e) {
    g_b2h_error = e.what();
    return -1;
  }
}
const char* b2h_last_error(void) { return g_b2h_error.c_str(); }
int b2h_hex_prolongator_row(int family, int a, int b, int c, int* idx, double* val) {
  return HexElement::prolongator_row(family, a, b, c, idx, val);
}

}  // extern "C"
