// C++ driver of the steady Stokes application through the FEMuS-shaped adapter classes: the sequence of
// applications/003_NavierStokes/SteadyStokes/main.cpp + LinearImplicitSystem::MGsolve on a box mesh in three
// dimensions (U, V, W triquadratic, P trilinear), written against B200Vector / B200Matrix /
// LinearEquationSolverB200Asm (the LinearEquationSolverPetscAsm surface: Vanka blocks with the pressure as Schur
// variable) and the host mesh layer.  tests/test_zz_stokes_gpu.py runs it on the GPU against the oracle.
//
//   stokes_driver nx ny nz nlevels ncycles [ns]
//
// Velocity Dirichlet on the boundary sets 1, 3, 4, 5, 6 (U = 1 on set 6), natural outflow on set 2, IRe = 1.
// "ns": the steady Navier-Stokes routine (03_navier_stokes.hpp) at nu = 0.1 instead -- every cycle is then one Newton
// iteration of NonLinearImplicitSystem::solve (residual + exact Jacobian at Sol, three V-cycles, update).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include "../../femus_b200/host/LinearEquationSolverB200Asm.hpp"

using namespace femus;
using namespace femus_b200;

int main(int argc, char** argv) {
  if (argc < 6) { std::fprintf(stderr, "usage: %s nx ny nz nlevels ncycles\n", argv[0]); return 2; }
  const int nx = std::atoi(argv[1]), ny = std::atoi(argv[2]), nz = std::atoi(argv[3]), nl = std::atoi(argv[4]), ncycles = std::atoi(argv[5]);
  const bool ns = argc > 6 && std::strcmp(argv[6], "ns") == 0;
  const int vcycles = ns ? 3 : 1;
  const std::vector<int> families = {BIQUADRATIC, BIQUADRATIC, BIQUADRATIC, LINEAR};       // system "Navier-Stokes": U, V, W, P

  std::vector<MeshLevel> msh;
  msh.push_back(GenerateCoarseBoxMesh(nx, ny, nz, 0., 1., 0., 1., 0., 1., nullptr, 1));
  for (int l = 1; l < nl; l++) msh.push_back(RefineMesh(msh.back()));
  uint8_t walls[4 * 7] = {0}, lid[4 * 7] = {0};
  for (int k = 0; k < 3; k++)
    for (int f : {1, 3, 4, 5, 6}) walls[k * 7 + f] = 1;
  lid[0 * 7 + 6] = 1;

  // system.init(): per-level solver with _KK on the multi-variable pattern, _RES, _EPS; system prolongators
  std::vector<std::unique_ptr<LinearEquationSolverB200Asm>> LinSolver;
  std::vector<std::unique_ptr<B200Matrix>> PP(nl);
  for (int l = 0; l < nl; l++) {
    LinSolver.emplace_back(new LinearEquationSolverB200Asm((unsigned)l));
    const SystemLayout sys(msh[l], families);
    LinSolver[l]->InitPdeSystem(msh[l], families, SystemBdc(msh[l], sys, walls));
    LinSolver[l]->SetMesh(&msh[l], families);
    LinSolver[l]->SetNumberOfSchurVariables(1);          // the pressure
    LinSolver[l]->SetElementBlockNumber(1);
    LinSolver[l]->SetRichardsonScaleFactor(1.0);
    if (l == 0) LinSolver[l]->SetCoarseDirect(true);      // PREONLY + LU on the coarsest level
  }
  for (int l = 1; l < nl; l++) {
    const HostCsr P = BuildSystemProlongator(msh[l - 1], msh[l], families);
    PP[l].reset(new B200Matrix);
    PP[l]->init_from_csr((int)P.nrows, (int)P.ncols, P.rowptr.data(), P.col.data(), P.val.data());
    std::vector<int> fine(LinSolver[l]->BdcIndex().begin(), LinSolver[l]->BdcIndex().end());
    std::vector<int> coarse(LinSolver[l - 1]->BdcIndex().begin(), LinSolver[l - 1]->BdcIndex().end());
    PP[l]->mat_zero_rows(fine, 0.);
    PP[l]->mat_zero_cols(coarse);
  }

  // assembly plan of the finest level (the replacement of the AssembleMatrixResNS callback)
  const MeshLevel& top = msh[nl - 1];
  LinearEquationSolverB200Asm& fine = *LinSolver[nl - 1];
  const SystemLayout sys(top, families);
  const std::vector<int32_t> edof = SystemElementDofs(top, sys);
  const HexElement::Tables tv = HexElement::tables(BIQUADRATIC), tp = HexElement::tables(LINEAR);
  b2_mesh* dmesh = nullptr;
  b2_stokes* plan = nullptr;
  B2_ABORT_IF(b2_mesh_create(B200Context::get(), top.nnode, top.nel, top.xyz.data(), top.conn.data(), &dmesh), "b2_mesh_create");
  if (ns)
    B2_ABORT_IF(b2_ns_create(dmesh, fine._KK->handle(), edof.data(), tv.nve, tp.nve, HexElement::NG, tv.phi.data(), tv.dxi.data(), tv.deta.data(),
                             tv.dzeta.data(), tv.w.data(), tp.phi.data(), &plan),
                "b2_ns_create");
  else
    B2_ABORT_IF(b2_stokes_create(dmesh, fine._KK->handle(), edof.data(), tv.nve, tp.nve, HexElement::NG, tv.dxi.data(), tv.deta.data(),
                                 tv.dzeta.data(), tv.w.data(), tp.phi.data(), &plan),
                "b2_stokes_create");
  B200Vector Sol((int)sys.size());
  {
    const std::vector<double> on_lid = SystemBdc(top, sys, lid);
    for (size_t i = 0; i < on_lid.size(); i++)
      if (on_lid[i] < 1.5) Sol.set((int)i, 1.0);
    Sol.close();
  }

  const std::vector<unsigned> vars = {0u, 1u, 2u, 3u};
  std::printf("levels %d rows %d nnz %lld blocks %lld\n", nl, fine._KK->m(), (long long)fine._KK->nnz(), 0LL);
  for (int cycle = 0; cycle <= ncycles; cycle++) {
    fine.SetResZero();
    fine.SetEpsZero();
    fine._KK->zero();
    if (ns) B2_ABORT_IF(b2_ns_assemble(plan, Sol.handle(), fine._RES->handle(), 0.1), "b2_ns_assemble");
    else B2_ABORT_IF(b2_stokes_assemble(plan, Sol.handle(), fine._RES->handle(), 1.0), "b2_stokes_assemble");
    fine._KK->touched();
    fine._RES->touched();
    {
      std::vector<double> r;
      fine._RES->localize(r);
      double s = 0.;
      for (size_t i = 0; i < r.size(); i++) if (fine.Bdc()[i] > 1.1) s += r[i] * r[i];
      std::printf("cycle %d residual %.17e\n", cycle, std::sqrt(s));
    }
    if (cycle == ncycles) break;
    for (int l = nl - 1; l > 0; l--) LinSolver[l - 1]->_KK->matrix_PtAP(*PP[l], *LinSolver[l]->_KK, true);
    fine.MGInit(MULTIPLICATIVE, (unsigned)nl, PREONLY_B200);
    for (int l = 0; l < nl; l++) LinSolver[l]->MGSetLevel(&fine, (unsigned)(nl - 1), vars, l ? PP[l].get() : nullptr, nullptr, 1, 1);
    for (int v = 0; v < vcycles; v++) fine.MGSolve(true);
    fine.MGClear();
    Sol += *fine._EPS;
  }
  std::printf("solution l2 %.17e linf %.17e vanka blocks %lld groups %lld\n", Sol.l2_norm(), Sol.linfty_norm(), (long long)fine.BlockNumber(),
              (long long)fine.GroupNumber());
  b2_stokes_destroy(plan);
  b2_mesh_destroy(dmesh);
  return 0;
}
