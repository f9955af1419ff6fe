// The REPLACEMENT assembly callback inside a FEMuS build: same signature as the application's own
// callback (`AssembleFunctionType`, System.hpp:113; registered with SetAssembleFunction, called at
// LinearImplicitSystem.cpp:326), same result as the element loop of applications/001_Poisson/main.cpp:283-609 for
// constant diffusivity nu and constant source f -- but the loop runs on the device: b2_asm_poisson on the level's mesh,
// b2_asm_neumann_faces on the boundary faces whose condition is a flux.  Everything it needs is read out of the
// reference's own objects: coordinates (Mesh::GetTopology()->_Sol), connectivity (GetSolutionDof), matrix rows
// (LinearEquation::GetSystemDof), the tables of the reference's own finite element objects
// (_finiteElement[type][order]: GetPhi, GetDPhiDXi ..., Gauss weights), the boundary function of MultiLevelSolution.
// One plan per (level, element type), built at the first call of a level and rebuilt when the level's matrix object was
// replaced.  Selectively refined (AMR) levels need nothing special: a level lists coarse and fine elements alike, the
// hanging-node constraints act on the assembled matrix (KK <- Pamr^T KKamr Pamr, LinearImplicitSystem.cpp:329-342).
#pragma once
#include <map>
#include <memory>
#include "RefBackend.hpp"
#include "LinearImplicitSystem.hpp"
#include "MultiLevelProblem.hpp"
#include "MultiLevelSolution.hpp"

namespace femus {

struct B200AssembleOptions {
  std::string system = "Poisson", solution = "Sol";
  double nu = 1.0, source = 1.0;
};
inline B200AssembleOptions& B200AssembleSettings() {
  static B200AssembleOptions o;
  return o;
}

namespace b200_detail {

struct TypePlan {
  b2_mesh* mesh = nullptr;
  b2_asm* plan = nullptr;
  // boundary faces with a flux, per face kind (number of face dofs): element (local to this plan), local face, value
  struct Faces { int nvf, ngf; std::vector<int32_t> elem, local; std::vector<double> value, phi, dxi, deta, w; };
  std::vector<Faces> faces;
  std::vector<int32_t> face_nodes;      // [6][9] local nodes of the element type's faces
  ~TypePlan() {
    if (plan) b2_asm_destroy(plan);
    if (mesh) b2_mesh_destroy(mesh);
  }
};
struct LevelPlans {
  uint64_t generation = 0;
  std::vector<std::unique_ptr<TypePlan>> types;
};

inline void build_level(MultiLevelProblem& ml_prob, LinearImplicitSystem& sys, const unsigned level, B200Matrix& KK, LevelPlans& out) {
  const B200AssembleOptions& opt = B200AssembleSettings();
  Mesh* msh = ml_prob._ml_msh->GetLevel(level);
  elem* el = msh->GetMeshElements();
  MultiLevelSolution* ml_sol = ml_prob._ml_sol;
  LinearEquationSolver* pde = sys._LinSolver[level];
  const unsigned solIndex = ml_sol->GetIndex(opt.solution.c_str()), solPde = sys.GetSolPdeIndex(opt.solution.c_str());
  const unsigned order = ml_sol->GetSolutionType(solIndex);
  const unsigned iproc = msh->processor_id();
  const unsigned dim = msh->GetDimension();
  if (dim != 3) { std::fprintf(stderr, "femus_b200: AssemblePoissonB200 assembles 3-D meshes\n"); std::abort(); }
  const int64_t nnode = msh->GetNumberOfNodes();
  std::vector<double> xyz((size_t)3 * nnode);
  for (unsigned k = 0; k < 3; k++) {
    std::vector<double> c;
    msh->GetTopology()->_Sol[k]->localize(c);
    std::copy(c.begin(), c.begin() + nnode, xyz.begin() + (size_t)k * nnode);
  }
  const int e0 = msh->GetElementOffset(iproc), e1 = msh->GetElementOffset(iproc + 1);
  std::map<short unsigned, std::vector<int>> by_type;
  for (int iel = e0; iel < e1; iel++) by_type[msh->GetElementType(iel)].push_back(iel);
  // the matrix with the element-coupling pattern on the device (LinearEquation::InitPde sized it by counts: the first
  // close() of a count-initialised matrix would freeze whatever was inserted; here the pattern comes from the dof lists)
  if (!KK.frozen()) {
    if (by_type.size() != 1) { std::fprintf(stderr, "femus_b200: AssemblePoissonB200: a level with several element types keeps the host-built pattern: close() the matrix once before\n"); std::abort(); }
    const short unsigned t = by_type.begin()->first;
    const std::vector<int>& els = by_type.begin()->second;
    const int nve = (int)msh->GetElementDofNumber(els[0], order);
    std::vector<int32_t> dof((size_t)els.size() * nve);
    for (size_t e = 0; e < els.size(); e++)
      for (int i = 0; i < nve; i++) dof[e * nve + i] = (int32_t)pde->GetSystemDof(solIndex, solPde, i, els[e]);
    (void)t;
    KK.init_from_elements(KK.m(), (int64_t)els.size(), nve, dof.data());
  }
  out.types.clear();
  for (auto& kv : by_type) {
    const short unsigned t = kv.first;
    const std::vector<int>& els = kv.second;
    const elem_type* fe = msh->_finiteElement[t][order];
    const int nve = fe->GetNDofs(), ng = (int)fe->GetGaussPointNumber();
    const int nve2 = (int)msh->GetElementDofNumber(els[0], 2);
    std::vector<int32_t> conn((size_t)els.size() * 27, 0), dof((size_t)els.size() * nve);
    for (size_t e = 0; e < els.size(); e++) {
      for (int i = 0; i < nve2; i++) conn[e * 27 + i] = (int32_t)msh->GetSolutionDof(i, els[e], 2);
      for (int i = 0; i < nve; i++) dof[e * nve + i] = (int32_t)pde->GetSystemDof(solIndex, solPde, i, els[e]);
    }
    std::vector<double> phi((size_t)ng * nve), dxi(phi.size()), deta(phi.size()), dzeta(phi.size()), w(ng);
    for (int g = 0; g < ng; g++) {
      w[g] = fe->GetGaussWeight(g);
      for (int i = 0; i < nve; i++) {
        phi[(size_t)g * nve + i] = fe->GetPhi(g)[i];
        dxi[(size_t)g * nve + i] = fe->GetDPhiDXi(g)[i];
        deta[(size_t)g * nve + i] = fe->GetDPhiDEta(g)[i];
        dzeta[(size_t)g * nve + i] = fe->GetDPhiDZeta(g)[i];
      }
    }
    std::unique_ptr<TypePlan> P(new TypePlan);
    B2_ABORT_IF(b2_mesh_create(B200Context::get(), nnode, (int64_t)els.size(), xyz.data(), conn.data(), &P->mesh), "b2_mesh_create");
    B2_ABORT_IF(b2_asm_create(P->mesh, KK.handle(), nve, dof.data(), ng, phi.data(), dxi.data(), deta.data(), dzeta.data(), w.data(), &P->plan),
                "b2_asm_create");
    // boundary faces whose condition is a flux (main.cpp:552-585: the boundary function says "not Dirichlet" with tau != 0)
    P->face_nodes.assign(54, -1);
    std::map<int, size_t> kind;
    std::vector<double> xx(3, 0.);
    for (size_t e = 0; e < els.size(); e++) {
      const int iel = els[e];
      for (unsigned jf = 0; jf < msh->GetElementFaceNumber(iel); jf++) {
        const unsigned nvf = msh->GetElementFaceDofNumber(iel, jf, order);
        for (unsigned i = 0; i < nvf && i < 9; i++) P->face_nodes[jf * 9 + i] = (int32_t)msh->GetLocalFaceVertexIndex(iel, jf, i);
        if (el->GetFaceElementIndex(iel, jf) >= 0) continue;
        const unsigned faceIndex = el->GetBoundaryIndex(iel, jf);
        double tau = 0.;
        if (ml_sol->GetBdcFunction()(xx, opt.solution.c_str(), tau, faceIndex, 0.) || tau == 0.) continue;
        const unsigned felt = msh->GetElementFaceType(iel, jf);
        const elem_type* ff = msh->_finiteElement[felt][order];
        auto it = kind.find((int)felt);
        if (it == kind.end()) {
          TypePlan::Faces F;
          F.nvf = ff->GetNDofs();
          F.ngf = (int)ff->GetGaussPointNumber();
          F.phi.resize((size_t)F.ngf * F.nvf); F.dxi.resize(F.phi.size()); F.deta.resize(F.phi.size()); F.w.resize(F.ngf);
          for (int g = 0; g < F.ngf; g++) {
            F.w[g] = ff->GetGaussWeight(g);
            for (int i = 0; i < F.nvf; i++) {
              F.phi[(size_t)g * F.nvf + i] = ff->GetPhi(g)[i];
              F.dxi[(size_t)g * F.nvf + i] = ff->GetDPhiDXi(g)[i];
              F.deta[(size_t)g * F.nvf + i] = ff->GetDPhiDEta(g)[i];
            }
          }
          P->faces.push_back(std::move(F));
          it = kind.emplace((int)felt, P->faces.size() - 1).first;
        }
        TypePlan::Faces& F = P->faces[it->second];
        F.elem.push_back((int32_t)e);
        F.local.push_back((int32_t)jf);
        F.value.push_back(tau);
      }
    }
    out.types.push_back(std::move(P));
  }
  out.generation = KK.generation();
}

}  // namespace b200_detail

// drop-in for the application's assembly callback: system.SetAssembleFunction(femus::AssemblePoissonB200);
inline void AssemblePoissonB200(MultiLevelProblem& ml_prob) {
  const B200AssembleOptions& opt = B200AssembleSettings();
  LinearImplicitSystem& sys = ml_prob.get_system<LinearImplicitSystem>(opt.system.c_str());
  const unsigned level = sys.GetLevelToAssemble();
  LinearEquationSolver* pde = sys._LinSolver[level];
  B200Matrix& KK = static_cast<B200Matrix&>(*pde->_KK);
  B200Vector& RES = static_cast<B200Vector&>(*pde->_RES);
  const unsigned solIndex = ml_prob._ml_sol->GetIndex(opt.solution.c_str());
  B200Vector& Sol = static_cast<B200Vector&>(*ml_prob._ml_sol->GetSolutionLevel(level)->_Sol[solIndex]);
  static std::map<std::pair<const void*, unsigned>, b200_detail::LevelPlans> plans;      // per (system, level)
  b200_detail::LevelPlans& L = plans[std::make_pair((const void*)&sys, level)];
  if (L.types.empty() || L.generation != KK.generation() || !KK.frozen()) b200_detail::build_level(ml_prob, sys, level, KK, L);
  KK.zero();                                   // main.cpp:346
  Sol.close();
  RES.close();
  for (auto& P : L.types) {
    B2_ABORT_IF(b2_asm_poisson(P->plan, Sol.handle(), RES.handle(), opt.nu, opt.source), "b2_asm_poisson");
    for (auto& F : P->faces)
      B2_ABORT_IF(b2_asm_neumann_faces(P->plan, (int64_t)F.elem.size(), F.elem.data(), F.local.data(), F.value.data(), F.nvf, F.ngf, F.phi.data(),
                                       F.dxi.data(), F.deta.data(), F.w.data(), P->face_nodes.data(), RES.handle()),
                  "b2_asm_neumann_faces");
  }
  KK.touched();
  RES.touched();
  RES.close();
  KK.close();
}

}  // namespace femus
