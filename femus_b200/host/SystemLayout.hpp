// Systems of several Lagrange variables on a mesh level (host side): the row numbering, sparsity
// pattern, prolongator and Dirichlet flags LinearEquation / LinearImplicitSystem build for them.  Reference:
//   LinearEquation::InitPde, GetSystemDof          LinearEquation.cpp:76-85, 211-237     rows [rank][variable][dof]
//   LinearEquation::GetSparsityPatternSize          LinearEquation.cpp:407-548            every element couples the
//       dofs of variable i with those of variable j wherever _SparsityPattern[i * nvars + j] is set (default: all)
//   LinearImplicitSystem::BuildProlongatorMatrix     LinearImplicitSystem.cpp:826-909      variable by variable: the
//       scalar prolongator of the variable's family between the variable's rows of the two levels
//   ZeroInterpolatorDirichletNodes / BuildBdcIndex   LinearImplicitSystem.cpp:1032-1120, LinearEquationSolverPetsc.cpp:53-90
#pragma once
#include <algorithm>
#include <cstdint>
#include <stdexcept>
#include <vector>
#include "GeneralMesh.hpp"

namespace femus_b200 {

// LinearEquation::InitPde (LinearEquation.cpp:211-237): rows of a system of several variables are numbered
// [rank][variable][dof]; KKoffset[k][p] = first row of variable k on rank p (k = nvars: end of the rank's rows).
struct SystemLayout {
  std::vector<int> family;                        // FE family of every variable (0 linear, 1 serendipity, 2 biquadratic)
  std::vector<std::vector<int64_t>> KKoffset;     // [nvars+1][nprocs]
  SystemLayout(const MeshLevel& L, const std::vector<int>& fam) : family(fam) {
    if (fam.empty()) throw std::invalid_argument("SystemLayout: no variable");
    for (int f : fam)
      if (f < 0 || f > 2) throw std::invalid_argument("SystemLayout: Lagrange families 0, 1, 2 only");
    const int nv = (int)fam.size();
    KKoffset.assign((size_t)nv + 1, std::vector<int64_t>((size_t)L.nprocs, 0));
    for (int j = 1; j <= nv; j++) KKoffset[j][0] = KKoffset[j - 1][0] + (L.dof_offset[fam[j - 1]][1] - L.dof_offset[fam[j - 1]][0]);
    for (int i = 1; i < L.nprocs; i++) {
      KKoffset[0][i] = KKoffset[nv][i - 1];
      for (int j = 1; j <= nv; j++) KKoffset[j][i] = KKoffset[j - 1][i] + (L.dof_offset[fam[j - 1]][i + 1] - L.dof_offset[fam[j - 1]][i]);
    }
  }
  int nvars() const { return (int)family.size(); }
  int64_t size() const { return KKoffset.back().back(); }
  // LinearEquation::GetSystemDof (LinearEquation.cpp:76-85)
  int64_t system_dof(const MeshLevel& L, int k, int i, int64_t iel) const {
    const int f = family[k];
    const int64_t idof = L.GetSolutionDof(i, iel, f);
    const std::vector<int64_t>& o = L.dof_offset[f];
    const int isub = (int)(std::upper_bound(o.begin(), o.end(), idof) - o.begin()) - 1;
    return KKoffset[k][isub] + idof - o[isub];
  }
};


// [nel][27 * nvars]: system dofs of every element, variable k in columns [27 k, 27 k + nve_k), -1 elsewhere
inline std::vector<int32_t> SystemElementDofs(const MeshLevel& L, const SystemLayout& sys) {
  const int nv = sys.nvars();
  std::vector<int32_t> d((size_t)L.nel * 27 * nv, -1);
  for (int64_t e = 0; e < L.nel; e++)
    for (int k = 0; k < nv; k++)
      for (int i = 0; i < ElemTopology::nve(L.type_of(e), sys.family[k]); i++) d[(e * nv + k) * 27 + i] = (int32_t)sys.system_dof(L, k, i, e);
  return d;
}

// GetSparsityPatternSize + SparseMatrix::init: union over the elements of (dofs of variable i) x (dofs of variable
// j) for every coupled pair, zeros included, columns sorted.  pattern: nvars x nvars flags, NULL = all coupled.
inline HostCsr BuildSystemSparsity(const MeshLevel& L, const SystemLayout& sys, const uint8_t* pattern) {
  const int nv = sys.nvars();
  HostCsr A;
  A.nrows = A.ncols = sys.size();
  std::vector<std::vector<int32_t>> rows((size_t)A.nrows);
  const std::vector<int32_t> d = SystemElementDofs(L, sys);
  for (int64_t e = 0; e < L.nel; e++)
    for (int i = 0; i < nv; i++)
      for (int j = 0; j < nv; j++) {
        if (pattern && !pattern[i * nv + j]) continue;
        const int32_t* di = &d[(e * nv + i) * 27];
        const int32_t* dj = &d[(e * nv + j) * 27];
        const int ni = ElemTopology::nve(L.type_of(e), sys.family[i]), nj = ElemTopology::nve(L.type_of(e), sys.family[j]);
        for (int a = 0; a < ni; a++) rows[di[a]].insert(rows[di[a]].end(), dj, dj + nj);
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

// solution dof of family f -> system row of variable k
inline int64_t SystemRowOfSolutionDof(const MeshLevel& L, const SystemLayout& sys, int k, int64_t idof) {
  const std::vector<int64_t>& o = L.dof_offset[sys.family[k]];
  const int isub = (int)(std::upper_bound(o.begin(), o.end(), idof) - o.begin()) - 1;
  return sys.KKoffset[k][isub] + idof - o[isub];
}

// BuildProlongatorMatrix of the system: block diagonal by variable through the two levels' row numberings
inline HostCsr BuildSystemProlongator(const MeshLevel& C, const MeshLevel& F, const std::vector<int>& family) {
  const SystemLayout sc(C, family), sf(F, family);
  HostCsr P;
  P.nrows = sf.size();
  P.ncols = sc.size();
  std::vector<std::vector<std::pair<int32_t, double>>> rows((size_t)P.nrows);
  HostCsr scalar[3];
  bool have[3] = {false, false, false};
  for (int k = 0; k < sf.nvars(); k++) {
    const int f = family[k];
    if (!have[f]) { scalar[f] = BuildAnyProlongator(C, F, f); have[f] = true; }
    const HostCsr& S = scalar[f];
    for (int64_t r = 0; r < S.nrows; r++) {
      std::vector<std::pair<int32_t, double>>& row = rows[SystemRowOfSolutionDof(F, sf, k, r)];
      for (int64_t q = S.rowptr[r]; q < S.rowptr[r + 1]; q++) row.emplace_back((int32_t)SystemRowOfSolutionDof(C, sc, k, S.col[q]), S.val[q]);
    }
  }
  P.rowptr.assign(P.nrows + 1, 0);
  for (int64_t r = 0; r < P.nrows; r++) {
    std::sort(rows[r].begin(), rows[r].end());
    P.rowptr[r + 1] = P.rowptr[r] + (int64_t)rows[r].size();
  }
  P.col.resize(P.rowptr[P.nrows]);
  P.val.resize(P.rowptr[P.nrows]);
  for (int64_t r = 0; r < P.nrows; r++)
    for (size_t q = 0; q < rows[r].size(); q++) {
      P.col[P.rowptr[r] + q] = rows[r][q].first;
      P.val[P.rowptr[r] + q] = rows[r][q].second;
    }
  return P;
}

// Dirichlet flags of the system rows: variable k is Dirichlet on the boundary sets flagged in dirichlet[k * 7 + 1..6]
// (MultiLevelSolution::GenerateBdc per variable), 2 = free / 0 = Dirichlet, laid out in system numbering
inline std::vector<double> SystemBdc(const MeshLevel& L, const SystemLayout& sys, const uint8_t* dirichlet) {
  std::vector<double> out((size_t)sys.size(), 2.0);
  for (int k = 0; k < sys.nvars(); k++) {
    bool f[7];
    for (int i = 0; i < 7; i++) f[i] = dirichlet[k * 7 + i] != 0;
    const std::vector<double> b = L.GenerateBdc(sys.family[k], f);
    for (int64_t i = 0; i < (int64_t)b.size(); i++) out[SystemRowOfSolutionDof(L, sys, k, i)] = b[i];
  }
  return out;
}

}  // namespace femus_b200
