// C++ driver of the in-scope application through the FEMuS-shaped adapter classes: the sequence of
// applications/001_Poisson/main.cpp + LinearImplicitSystem::MGsolve (LinearImplicitSystem.cpp:288-411)
// on a box mesh, written against B200Vector / B200Matrix / LinearEquationSolverB200 and the host mesh
// layer.  tests/test_adapters.py runs it on the GPU and checks what it prints against the oracle.
//
//   poisson_driver nx ny nz nlevels family(0 linear | 2 biquadratic) ncycles [compat | asm<N> | asmref<N> | asmsor<N> | asmilu<N>]
//
// "asm<N>" (e.g. asm8) runs the levels through LinearEquationSolverB200Asm: the element-block smoother of
// "smoother": "asm" (main.cpp:234-250) with N elements per block, Richardson scale 1, coloured sweep; "asmref<N>"
// sweeps the blocks in the reference's own order; "asmsor<N>" uses one SSOR iteration as the block solve
// (SetPreconditionerFineGrids(SOR_PRECOND), main.cpp:242) -- asmsor4096 is the application's own "asm" setting;
// "asmilu<N>" uses ILU(0) (ILU_PRECOND); "gmresilu" = GMRES as level solver around ILU(0) of the whole level (one block
// with every element): the reference's DEFAULT level solver (LinearEquationSolverPetsc.hpp:128-151).
// "compat" additionally rebuilds the finest matrix through the slow plugin path (init with counts,
// add_matrix_blocked per element, close) from the rows of the device-assembled one and checks that
// both give the same matrix-vector product.
#include <cstdio>
#include <cstdlib>
#include <memory>
#include "../../femus_b200/host/BoxMesh.hpp"
#include <cstring>
#include "../../femus_b200/host/LinearEquationSolverB200Asm.hpp"

using namespace femus;
using namespace femus_b200;

int main(int argc, char** argv) {
  if (argc < 7) { std::fprintf(stderr, "usage: %s nx ny nz nlevels family ncycles [compat]\n", argv[0]); return 2; }
  const int nx = std::atoi(argv[1]), ny = std::atoi(argv[2]), nz = std::atoi(argv[3]), nl = std::atoi(argv[4]);
  const int family = std::atoi(argv[5]), ncycles = std::atoi(argv[6]);
  const bool gmres_ilu = argc > 7 && std::strcmp(argv[7], "gmresilu") == 0;
  const bool use_asm = argc > 7 && (std::strncmp(argv[7], "asm", 3) == 0 || gmres_ilu);
  const bool asm_ref = use_asm && std::strncmp(argv[7], "asmref", 6) == 0;
  const bool asm_sor = use_asm && std::strncmp(argv[7], "asmsor", 6) == 0;
  const bool asm_ilu = use_asm && (std::strncmp(argv[7], "asmilu", 6) == 0 || gmres_ilu);
  const int asm_blocks = gmres_ilu ? (1 << 30) : (use_asm ? std::atoi(argv[7] + ((asm_ref || asm_sor || asm_ilu) ? 6 : 3)) : 0);
  const bool compat = argc > 7 && !use_asm;
  const int nve = HexElement::nve(family);

  // mlMsh.GenerateCoarseBoxMesh(...); mlMsh.RefineMesh(nl, nl, NULL)            (main.cpp:133-141)
  std::vector<MeshLevel> msh;
  msh.push_back(GenerateCoarseBoxMesh(nx, ny, nz, 0., 1., 0., 1., 0., 1., nullptr, 1));
  for (int l = 1; l < nl; l++) msh.push_back(RefineMesh(msh.back()));
  // mlSol.AddSolution("Sol", LAGRANGE, order); GenerateBdc("All")                 (main.cpp:149-184)
  const bool dirichlet[7] = {false, true, true, true, true, true, true};

  // system.init(): per-level LinearEquationSolver with _KK, _RES, _EPS; prolongators   (LinearImplicitSystem.cpp:138-282)
  std::vector<std::unique_ptr<LinearEquationSolverB200>> LinSolver;
  std::vector<std::unique_ptr<B200Matrix>> PP(nl);
  for (int l = 0; l < nl; l++) {
    if (use_asm) {       // system.SetLinearEquationSolverType(FEMuS_ASM); SetNumberOfSchurVariables(0); SetElementBlockNumber(n)
      LinearEquationSolverB200Asm* s = new LinearEquationSolverB200Asm((unsigned)l);
      s->SetMesh(&msh[l], family);
      s->SetNumberOfSchurVariables(0);
      s->SetElementBlockNumber((unsigned)std::min<int64_t>(asm_blocks, msh[l].nel));      // LinearImplicitSystem.cpp:1198
      s->SetSweepOrder(asm_ref ? 0 : 1);
      s->set_preconditioner_type(asm_sor ? SOR_PRECOND_B200 : (asm_ilu ? ILU_PRECOND_B200 : MLU_PRECOND_B200));
      LinSolver.emplace_back(s);
    } else {
      LinSolver.emplace_back(new LinearEquationSolverB200((unsigned)l));
    }
    const std::vector<int32_t> dof = msh[l].system_dofs(family);
    LinSolver[l]->InitPde((int)msh[l].ndofs(family), msh[l].nel, nve, dof.data(), msh[l].GenerateBdc(family, dirichlet));
    LinSolver[l]->SetTolerances(1.e-10, 1.e-20, 1.e+50, 1, 30);
    LinSolver[l]->SetRichardsonScaleFactor(use_asm ? 1.0 : 0.5);
    if (gmres_ilu) LinSolver[l]->set_solver_type(GMRES_B200);
  }
  for (int l = 1; l < nl; l++) {      // BuildProlongatorMatrix + ZeroInterpolatorDirichletNodes (:826-909, :1032-1120)
    const HostCsr P = BuildProlongator(msh[l - 1], msh[l], family);
    PP[l].reset(new B200Matrix);
    PP[l]->init_from_csr((int)P.nrows, (int)P.ncols, P.rowptr.data(), P.col.data(), P.val.data());
    std::vector<int> fine(LinSolver[l]->BdcIndex().begin(), LinSolver[l]->BdcIndex().end());
    std::vector<int> coarse(LinSolver[l - 1]->BdcIndex().begin(), LinSolver[l - 1]->BdcIndex().end());
    PP[l]->mat_zero_rows(fine, 0.);
    PP[l]->mat_zero_cols(coarse);
  }

  // assembly plan of the finest level (the replacement of the AssemblePoissonProblem callback)
  const MeshLevel& top = msh[nl - 1];
  LinearEquationSolverB200& fine = *LinSolver[nl - 1];
  b2_mesh* dmesh = nullptr;
  b2_asm* plan = nullptr;
  const std::vector<int32_t> topdof = top.system_dofs(family);
  const HexElement::Tables t = HexElement::tables(family);
  B2_ABORT_IF(b2_mesh_create(B200Context::get(), top.nnode, top.nel, top.xyz.data(), top.conn.data(), &dmesh), "b2_mesh_create");
  B2_ABORT_IF(b2_asm_create(dmesh, fine._KK->handle(), nve, topdof.data(), HexElement::NG, t.phi.data(), t.dxi.data(), t.deta.data(),
                            t.dzeta.data(), t.w.data(), &plan),
              "b2_asm_create");
  B200Vector Sol((int)top.ndofs(family));      // _Sol of the finest level, initial guess 0

  const std::vector<unsigned> vars(1, 0u);
  std::printf("levels %d family %d dofs %d nnz %lld\n", nl, family, fine._KK->m(), (long long)fine._KK->nnz());
  for (int cycle = 0; cycle <= ncycles; cycle++) {
    // ---- MGsolve: SetResZero, assemble the finest level                                  (:318-326)
    fine.SetResZero();
    fine.SetEpsZero();
    fine._KK->zero();
    B2_ABORT_IF(b2_asm_poisson(plan, Sol.handle(), fine._RES->handle(), 1.0, 1.0), "b2_asm_poisson");
    fine._KK->touched();
    fine._RES->touched();
    // UpdateRes + HasLinearConverged: ||Res||_2 over the dofs with Bdc > 1.1           (Solution.cpp:595-628)
    {
      std::vector<double> r;
      fine._RES->localize(r);
      double s = 0.;
      for (size_t i = 0; i < r.size(); i++) if (fine.Bdc()[i] > 1.1) s += r[i] * r[i];
      std::printf("cycle %d residual %.17e\n", cycle, std::sqrt(s));
    }
    if (cycle == ncycles) break;
    // ---- Galerkin chain KK[l-1] = PP[l]^T KK[l] PP[l]                                      (:347-370)
    for (int l = nl - 1; l > 0; l--) LinSolver[l - 1]->_KK->matrix_PtAP(*PP[l], *LinSolver[l]->_KK, true);
    // ---- MGInit / MGSetLevel / MGSolve / MGClear                                           (:376-400)
    fine.MGInit(MULTIPLICATIVE, (unsigned)nl, PREONLY_B200);
    for (int l = 0; l < nl; l++) LinSolver[l]->MGSetLevel(&fine, (unsigned)(nl - 1), vars, l ? PP[l].get() : nullptr, nullptr, 1, 1);
    fine.MGSolve(true);
    fine.MGClear();
    // ---- UpdateSol: Sol += EPS                                                             (Solution.cpp:544-590)
    Sol += *fine._EPS;
  }
  std::printf("solution l2 %.17e linf %.17e\n", Sol.l2_norm(), Sol.linfty_norm());

  if (compat) {
    // the unchanged-callback path: counts only, element blocks staged on the host, pattern frozen at close()
    fine._KK->zero();
    B2_ABORT_IF(b2_asm_poisson(plan, nullptr, nullptr, 1.0, 1.0), "b2_asm_poisson");
    fine._KK->touched();
    const int n = fine._KK->m();
    B200Matrix K2;
    std::vector<int> nnz(n), noz(n, 0);
    for (int i = 0; i < n; i++) nnz[i] = fine._KK->MatGetRowM(i);
    K2.init(n, n, n, n, nnz, noz);
    std::vector<int> cols(256), row(1);
    std::vector<double> vals(256), blk;
    for (int pass = 0; pass < 2; pass++) {           // second pass goes through the frozen-pattern staging
      if (pass) K2.zero();
      for (int i = 0; i < n; i++) {
        const int len = fine._KK->MatGetRowM(i, cols.data(), vals.data());
        row[0] = i;
        std::vector<int> c(cols.begin(), cols.begin() + len);
        blk.assign(vals.begin(), vals.begin() + len);
        K2.add_matrix_blocked(blk, row, c);
      }
      K2.close();
      B200Vector x(n), y1(n), y2(n);
      for (int i = 0; i < n; i++) x.set(i, std::sin(0.37 * i));
      x.close();
      y1.matrix_mult(x, *fine._KK);
      y2.matrix_mult(x, K2);
      y2 -= y1;
      std::printf("compat pass %d nnz %lld diff %.3e ref %.3e\n", pass, (long long)K2.nnz(), y2.linfty_norm(), y1.linfty_norm());
    }
  }
  b2_asm_destroy(plan);
  b2_mesh_destroy(dmesh);
  return 0;
}
