// Host side of the element-block (ASM / Vanka) smoother for a single Lagrange variable: the index sets the
// reference hands to PETSc's PCASM, and the schedules the device sweep runs them in.  Reference:
//   src/06_mesh/00_single_level/02_partitioning/MeshASMPartitioning.cpp:89-148          DoPartition
//   src/08_algebra.../03_solvers_with_preconditioner/petsc_asm/LinearEquationSolverPetscAsm.cpp:91-262   BuildASMIndex
// For applications/001_Poisson ("smoother": "asm", main.cpp:234-250) there is one variable and no Schur variable, so
// every block is the set of dofs of its own elements (GetElementNearElementSize(iel, 0) == 1); systems of several
// Lagrange variables whose last ones are Schur variables (velocity-pressure Vanka blocks) take the velocities of one
// layer of near elements and the pressures of the block's own elements.
//
// DoPartition: the owned elements of one material class, block_size at a time in element order; classes in the
// order 4 (solid), 3 (porous), 2 (fluid).  (The reference counts every element that is neither 4 nor 3 into the
// third class but only places material 2 there; levels with other materials are refused here.)
// BuildASMIndex: per block the "overlapping" set = dofs of the block's elements this rank owns, then the dofs other
// ranks own (ghosts), sorted; the "local" set = the owned ones no earlier block has claimed, sorted.
//
// Schedules (what makes the multiplicative sweep parallel without changing its result): two blocks DEPEND on each
// other when one writes (its dofs) what the other reads (the columns of its rows) or writes.
//   levels : group(j) = 1 + max group of the earlier blocks j depends on -- sweeping group by group IS the
//            reference's sequential sweep, whatever happens inside a group;
//   colours: greedy colouring of that dependency graph in block order; sweeping colour by colour equals the
//            reference's sweep with the block list stably sorted by colour (a permutation of the index-set list
//            handed to PCASMSetLocalSubdomains; few large groups).
#pragma once
#include <algorithm>
#include <cstdint>
#include <map>
#include <stdexcept>
#include <vector>
#include "BoxMesh.hpp"
#include "SystemLayout.hpp"

namespace femus_b200 {

struct AsmIndex {
  std::vector<int64_t> elem_ptr, local_ptr, overlap_ptr;     // [nblocks+1]
  std::vector<int32_t> elems, local, overlap;
  int64_t block_type_range[3] = {0, 0, 0};
  int64_t nblocks() const { return (int64_t)elem_ptr.size() - 1; }
};

// The element blocks of rank iproc (what MeshASMPartitioning::DoPartition produces): the rank's elements are split
// by material class -- solid (4) first, then porous (3), then fluid (2) -- keeping element order inside a class (a
// stable partition), and every class is cut into consecutive chunks of its block size, the last chunk ragged.
// block_type_range[k] = number of blocks up to and including class k.
inline void DoPartition(const MeshLevel& L, int iproc, const unsigned block_size[3], std::vector<std::vector<unsigned>>& block_elements,
                        int64_t block_type_range[3]) {
  static const int kClassOfMaterial[5] = {-1, -1, 2, 1, 0};      // material 4 -> class 0, 3 -> 1, 2 -> 2
  std::vector<unsigned> members[3];
  for (int64_t iel = L.elem_offset[iproc]; iel < L.elem_offset[iproc + 1]; iel++) {
    const int m = L.material.empty() ? 2 : (int)L.material[iel];
    if (m < 2 || m > 4) throw std::invalid_argument("DoPartition: element material must be 2, 3 or 4");
    members[kClassOfMaterial[m]].push_back((unsigned)iel);
  }
  block_elements.clear();
  for (int k = 0; k < 3; k++) {
    const std::vector<unsigned>& e = members[k];
    if (!e.empty() && block_size[k] == 0) throw std::invalid_argument("DoPartition: block size 0");
    for (size_t first = 0; first < e.size(); first += block_size[k])
      block_elements.emplace_back(e.begin() + (std::ptrdiff_t)first, e.begin() + (std::ptrdiff_t)std::min(e.size(), first + block_size[k]));
    block_type_range[k] = (int64_t)block_elements.size();
  }
}

// elem::BuildElementNearElement (Elem.cpp:493-526): the element itself, then every other element sharing a vertex
// with it, ascending; CSR-like (ptr, list) over all elements of the level
inline void BuildElementNearElement(const MeshLevel& L, std::vector<int64_t>& ptr, std::vector<int32_t>& list) {
  std::vector<int64_t> vptr((size_t)L.nnode + 1, 0);
  for (int64_t e = 0; e < L.nel; e++)
    for (int i = 0; i < ElemTopology::nvert(L.type_of(e)); i++) vptr[L.node(e, i) + 1]++;
  for (int64_t n = 0; n < L.nnode; n++) vptr[n + 1] += vptr[n];
  std::vector<int32_t> vlist((size_t)vptr[L.nnode]);
  std::vector<int64_t> fill(vptr.begin(), vptr.end() - 1);
  for (int64_t e = 0; e < L.nel; e++)
    for (int i = 0; i < ElemTopology::nvert(L.type_of(e)); i++) vlist[fill[L.node(e, i)]++] = (int32_t)e;
  ptr.assign(1, 0);
  list.clear();
  std::vector<int32_t> others;
  for (int64_t e = 0; e < L.nel; e++) {
    others.clear();
    for (int i = 0; i < ElemTopology::nvert(L.type_of(e)); i++) {
      const int32_t nd = L.node(e, i);
      for (int64_t q = vptr[nd]; q < vptr[nd + 1]; q++)
        if (vlist[q] != e) others.push_back(vlist[q]);
    }
    std::sort(others.begin(), others.end());
    others.erase(std::unique(others.begin(), others.end()), others.end());
    list.push_back((int32_t)e);
    list.insert(list.end(), others.begin(), others.end());
    ptr.push_back((int64_t)list.size());
  }
}

// BuildASMIndex for the system `sys` on rank iproc, the LAST nschur variables being Schur (pressure-like) variables:
// a block takes the non-Schur dofs of every owned element near its elements (FastVankaBlock == false for Lagrange
// Schur variables => one layer of near elements; no Schur variable => the elements themselves) and the Schur dofs of
// its own elements.  block_elems elements per block in every material class, capped by the level's element count
// (LinearImplicitSystem::SetElementBlockNumber, :1191-1201).
inline AsmIndex BuildAsmIndexSystem(const MeshLevel& L, const SystemLayout& sys, int nschur, unsigned block_elems, int iproc) {
  if (iproc < 0 || iproc >= L.nprocs) throw std::invalid_argument("BuildAsmIndex: rank out of range");
  if (block_elems == 0) throw std::invalid_argument("BuildAsmIndex: block size 0");
  const int nv = sys.nvars();
  if (nschur < 0 || nschur > nv) throw std::invalid_argument("BuildAsmIndex: bad number of Schur variables");
  const unsigned nb = (unsigned)std::min<int64_t>(block_elems, L.nel);
  const unsigned bs[3] = {nb, nb, nb};
  std::vector<std::vector<unsigned>> be;
  AsmIndex out;
  DoPartition(L, iproc, bs, be, out.block_type_range);
  const bool fast = nschur == 0;
  std::vector<int64_t> near_ptr;
  std::vector<int32_t> near;
  if (!fast) BuildElementNearElement(L, near_ptr, near);
  const int64_t e0 = L.elem_offset[iproc], e1 = L.elem_offset[iproc + 1];
  const int64_t d0 = sys.KKoffset[0][iproc], size = sys.KKoffset[nv][iproc] - d0;
  std::vector<int64_t> indexa((size_t)size, size), indexb((size_t)size, size);
  std::vector<char> owned((size_t)size, 0);
  std::map<int, bool> ghosts;
  out.elem_ptr.push_back(0);
  out.local_ptr.push_back(0);
  out.overlap_ptr.push_back(0);
  std::vector<int32_t> loc, ovl;
  std::vector<char> in_block((size_t)L.nel, 0);
  auto add = [&](int k, int64_t jel) {
    const int f = sys.family[k];
    const int nve = ElemTopology::nve(L.type_of(jel), f);
    for (int jj = 0; jj < nve; jj++) {
      const int64_t jdof = L.GetSolutionDof(jj, jel, f);
      const int64_t kk = sys.system_dof(L, k, jj, jel);
      if (jdof >= L.dof_offset[f][iproc] && jdof < L.dof_offset[f][iproc + 1]) {
        if (indexa[kk - d0] == size && !owned[kk - d0]) {
          owned[kk - d0] = 1;
          indexa[kk - d0] = (int64_t)loc.size();
          loc.push_back((int32_t)kk);
        }
        if (indexb[kk - d0] == size) {
          indexb[kk - d0] = (int64_t)ovl.size();
          ovl.push_back((int32_t)kk);
        }
      } else {
        ghosts[(int)kk] = true;
      }
    }
  };
  for (const std::vector<unsigned>& elems : be) {
    loc.clear();
    ovl.clear();
    std::vector<int64_t> added;
    for (unsigned iel : elems) {
      const int64_t n0 = fast ? 0 : near_ptr[iel], n1 = fast ? 1 : near_ptr[iel + 1];
      for (int64_t q = n0; q < n1; q++) {
        const int64_t jel = fast ? (int64_t)iel : (int64_t)near[q];
        if (jel < e0 || jel >= e1 || in_block[jel]) continue;
        in_block[jel] = 1;
        added.push_back(jel);
        for (int k = 0; k < nv - nschur; k++) add(k, jel);
      }
      for (int k = nv - nschur; k < nv; k++) add(k, iel);
    }
    for (int32_t kk : loc) indexa[kk - d0] = size;
    for (int32_t kk : ovl) indexb[kk - d0] = size;
    for (int64_t jel : added) in_block[jel] = 0;
    for (const auto& g : ghosts) ovl.push_back((int32_t)g.first);
    ghosts.clear();
    std::sort(loc.begin(), loc.end());
    std::sort(ovl.begin(), ovl.end());
    out.elems.insert(out.elems.end(), elems.begin(), elems.end());
    out.local.insert(out.local.end(), loc.begin(), loc.end());
    out.overlap.insert(out.overlap.end(), ovl.begin(), ovl.end());
    out.elem_ptr.push_back((int64_t)out.elems.size());
    out.local_ptr.push_back((int64_t)out.local.size());
    out.overlap_ptr.push_back((int64_t)out.overlap.size());
  }
  return out;
}

// one Lagrange variable without Schur variables (001_Poisson's "asm" setting)
inline AsmIndex BuildAsmIndex(const MeshLevel& L, int family, unsigned block_elems, int iproc) {
  return BuildAsmIndexSystem(L, SystemLayout(L, std::vector<int>(1, family)), 0, block_elems, iproc);
}

// Schedule of the multiplicative sweep over blocks (sorted dof lists blk_dofs[blk_ptr[b] .. blk_ptr[b+1])) of the
// operator with pattern (rowptr, col), n rows.  mode 0: dependency levels of the given block order; mode 1: greedy
// colours.  group[b] = group of block b; returns the number of groups.
inline int64_t AsmSchedule(int64_t n, const int64_t* rowptr, const int32_t* col, int64_t nblocks, const int64_t* blk_ptr,
                           const int32_t* blk_dofs, int mode, int32_t* group) {
  if (mode != 0 && mode != 1) throw std::invalid_argument("AsmSchedule: mode must be 0 (levels) or 1 (colours)");
  std::vector<int32_t> stamp((size_t)n, -1);
  std::vector<int32_t> cols;
  auto columns_of = [&](int64_t b) {          // distinct columns of the block's rows, and its own dofs
    cols.clear();
    for (int64_t k = blk_ptr[b]; k < blk_ptr[b + 1]; k++) {
      const int32_t r = blk_dofs[k];
      if (r < 0 || r >= n) throw std::invalid_argument("AsmSchedule: dof outside the operator");
      if (stamp[r] != (int32_t)b) { stamp[r] = (int32_t)b; cols.push_back(r); }
      for (int64_t q = rowptr[r]; q < rowptr[r + 1]; q++) {
        const int32_t c = col[q];
        if (c < 0 || c >= n) throw std::invalid_argument("AsmSchedule: column outside the operator");
        if (stamp[c] != (int32_t)b) { stamp[c] = (int32_t)b; cols.push_back(c); }
      }
    }
  };
  int64_t ngroups = 0;
  if (mode == 0) {
    std::vector<int32_t> wlev((size_t)n, -1), rlev((size_t)n, -1);   // highest group of an earlier writer / reader
    for (int64_t b = 0; b < nblocks; b++) {
      columns_of(b);
      int32_t lv = -1;
      for (int32_t c : cols) lv = std::max(lv, wlev[c]);
      for (int64_t k = blk_ptr[b]; k < blk_ptr[b + 1]; k++) lv = std::max(lv, rlev[blk_dofs[k]]);
      lv += 1;
      group[b] = lv;
      for (int64_t k = blk_ptr[b]; k < blk_ptr[b + 1]; k++) wlev[blk_dofs[k]] = std::max(wlev[blk_dofs[k]], lv);
      for (int32_t c : cols) rlev[c] = std::max(rlev[c], lv);
      ngroups = std::max<int64_t>(ngroups, lv + 1);
    }
    return ngroups;
  }
  // colours: blocks writing dof d / reading dof d, as lists; a block conflicts with every block that writes what it
  // reads or writes, or reads what it writes
  std::vector<int64_t> wptr((size_t)n + 1, 0), rptr((size_t)n + 1, 0);
  for (int64_t b = 0; b < nblocks; b++) {
    columns_of(b);
    for (int32_t c : cols) rptr[c + 1]++;
    for (int64_t k = blk_ptr[b]; k < blk_ptr[b + 1]; k++) wptr[blk_dofs[k] + 1]++;
  }
  for (int64_t i = 0; i < n; i++) { wptr[i + 1] += wptr[i]; rptr[i + 1] += rptr[i]; }
  std::vector<int32_t> wlist((size_t)wptr[n]), rlist((size_t)rptr[n]);
  {
    std::vector<int64_t> wp(wptr.begin(), wptr.end() - 1), rp(rptr.begin(), rptr.end() - 1);
    std::fill(stamp.begin(), stamp.end(), -1);
    for (int64_t b = 0; b < nblocks; b++) {
      columns_of(b);
      for (int32_t c : cols) rlist[rp[c]++] = (int32_t)b;
      for (int64_t k = blk_ptr[b]; k < blk_ptr[b + 1]; k++) wlist[wp[blk_dofs[k]]++] = (int32_t)b;
    }
  }
  std::fill(group, group + nblocks, -1);
  std::vector<int64_t> used;          // used[colour] == b + 1: a neighbour of block b has the colour
  std::fill(stamp.begin(), stamp.end(), -1);
  for (int64_t b = 0; b < nblocks; b++) {
    columns_of(b);
    auto mark = [&](int32_t other) {
      if (other == b || group[other] < 0) return;
      if ((size_t)group[other] >= used.size()) used.resize((size_t)group[other] + 1, 0);
      used[group[other]] = b + 1;
    };
    for (int32_t c : cols)
      for (int64_t q = wptr[c]; q < wptr[c + 1]; q++) mark(wlist[q]);
    for (int64_t k = blk_ptr[b]; k < blk_ptr[b + 1]; k++)
      for (int64_t q = rptr[blk_dofs[k]]; q < rptr[blk_dofs[k] + 1]; q++) mark(rlist[q]);
    int32_t colour = 0;
    while ((size_t)colour < used.size() && used[colour] == b + 1) colour++;
    group[b] = colour;
    ngroups = std::max<int64_t>(ngroups, colour + 1);
  }
  return ngroups;
}

}  // namespace femus_b200
