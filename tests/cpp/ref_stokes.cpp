// A FEMuS application for the multi-variable row of the path, written against the reference's
// public API and compiled together with the reference's own sources (oracle/ref_build): a 3-D box of HEX27 elements,
// three velocity components of the SECOND Lagrange family and a pressure of the FIRST (Taylor-Hood), the system
// "Stokes" with the variables U, V, W, P -- and the assembly callback of
// applications/003_NavierStokes/SteadyStokes/main.cpp (AssembleMatrixResSteadyStokes, :290-598) COMPILED IN PLACE: the
// reference's file is included unmodified, its main() renamed away.  Everything above the algebra backend is the
// REFERENCE's code: numbering of a system of several variables (LinearEquation::InitPde, GetSystemDof), sparsity
// with all couplings (GetSparsityPatternSize), Dirichlet flags per variable from the application's own
// SetBoundaryCondition (GenerateBdc), prolongators variable by variable (BuildProlongatorMatrix), the element loop.
//
//   ref_stokes <nx> <ny> <nz> <levels> [ns [fix|-] [newton <n>]]
//
// "ns": the system is the NonLinearImplicitSystem "NS" and the callback the reference's library routine
// femus::AssembleNavierStokes_AD (src/08_equations/assemble/03_navier_stokes.hpp:21-413: Galerkin residual with nu = 1,
// the Jacobian recorded by adept, the boundary pressure block on faces whose normal velocity is not Dirichlet); the
// boundary function is SteadyStokes's with a prescribed pressure 0.75 on boundary set 2.
//
// The velocity starts from the smooth fields below (so that the assembled residual F = -B sol is not trivial), the
// pressure from a linear one.  Run with FEMUS_REF_DUMP=<dir> on the host backend it leaves every level's numbering,
// pattern, assembled / Galerkin operator and the assembled residual behind: tests/golden/ref_stokes_*.npz.
// Test infrastructure.
#define main femus_steady_stokes_main_unused
#include "applications/003_NavierStokes/SteadyStokes/main.cpp"
#undef main
#include <fstream>
#include "NonLinearImplicitSystem.hpp"
#include "03_navier_stokes.hpp"

// SteadyStokes's boundary function in the form the library routine asks for, with a boundary pressure on set 2
static bool SetBoundaryConditionNS(const MultiLevelProblem*, const std::vector<double>& x, const char name[], double& value, const int facename,
                                   const double time) {
  const bool dirichlet = SetBoundaryCondition(x, name, value, facename, time);
  if (!strcmp(name, "P") && facename == 2) value = 0.75;
  return dirichlet;
}

// an enclosed flow (lid-driven cavity): every velocity component Dirichlet (U = 1 on boundary set 6), the pressure natural
// everywhere -- defined up to a constant, hence FixSolutionAtOnePoint("P")
static bool SetBoundaryConditionEnclosed(const MultiLevelProblem*, const std::vector<double>&, const char name[], double& value, const int facename,
                                         const double) {
  value = (!strcmp(name, "U") && facename == 6) ? 1. : 0.;
  return strcmp(name, "P") != 0;
}

static double InitU(const std::vector<double>& x) { return 0.3 * x[1] * (1.0 - x[2]) + 0.1 * x[0] * x[0]; }
static double InitV(const std::vector<double>& x) { return -0.2 * x[0] * x[2] + 0.05 * x[1]; }
static double InitW(const std::vector<double>& x) { return 0.15 * x[0] * x[1] - 0.1 * x[2] * x[2]; }
static double InitP(const std::vector<double>& x) { return 1.0 + 0.5 * x[0] - 0.25 * x[1] + 0.125 * x[2]; }

int main(int argc, char** argv) {
  if (argc < 5) {
    std::cerr << "usage: " << argv[0] << " <nx> <ny> <nz> <levels>\n";
    return 1;
  }
  const unsigned nx = std::atoi(argv[1]), ny = std::atoi(argv[2]), nz = std::atoi(argv[3]), nlevels = std::atoi(argv[4]);

  FemusInit init(argc, argv, MPI_COMM_WORLD);
  Files files;
  files.CheckIODirectories(true);

  MultiLevelMesh ml_msh;
  ml_msh.GenerateCoarseBoxMesh(nx, ny, nz, 0., 1., 0., 1., 0., 1., HEX27, "seventh");
  ml_msh.RefineMesh(nlevels, nlevels, NULL);
  ml_msh.PrintInfo();

  MultiLevelSolution ml_sol(&ml_msh);
  ml_sol.AddSolution("U", LAGRANGE, SECOND);
  ml_sol.AddSolution("V", LAGRANGE, SECOND);
  ml_sol.AddSolution("W", LAGRANGE, SECOND);
  ml_sol.AddSolution("P", LAGRANGE, FIRST);
  ml_sol.Initialize("U", InitU);
  ml_sol.Initialize("V", InitV);
  ml_sol.Initialize("W", InitW);
  ml_sol.Initialize("P", InitP);
  const bool ns = argc > 5 && std::string(argv[5]) == "ns";
  MultiLevelProblem ml_prob(&ml_sol);
  const bool fix = argc > 6 && std::string(argv[6]) == "fix";      // MultiLevelSolution::FixSolutionAtOnePoint("P"), as the enclosed-flow tutorials do
  if (ns) {
    if (fix) {
      ml_sol.AttachSetBoundaryConditionFunction(SetBoundaryConditionEnclosed);
      ml_sol.FixSolutionAtOnePoint("P");
    } else {
      ml_sol.AttachSetBoundaryConditionFunction(SetBoundaryConditionNS);
    }
    ml_sol.GenerateBdc("All", "Steady", &ml_prob);
  } else {
    ml_sol.AttachSetBoundaryConditionFunction(SetBoundaryCondition);      // SteadyStokes's own
    ml_sol.GenerateBdc("U");
    ml_sol.GenerateBdc("V");
    ml_sol.GenerateBdc("W");
    ml_sol.GenerateBdc("P");
  }

  Parameter parameter(1., 1.);
  Fluid fluid(parameter, 0.25, 1., "Newtonian");       // IReynolds = viscosity / (density Uref Lref) = 0.25
  std::cout << "Fluid properties: " << std::endl << fluid << std::endl;
  ml_prob.parameters.set<Fluid>("Fluid") = fluid;

  LinearImplicitSystem* sys = nullptr;
  if (ns) {
    NonLinearImplicitSystem& nls = ml_prob.add_system<NonLinearImplicitSystem>("NS");
    // "newton <n>" after "ns": n Newton iterations (on ONE level the linear solves are the host backend's exact LU, so the
    // "Nonlinear Eps_l2norm" lines the reference prints are the Newton updates themselves); otherwise one, for the dump
    unsigned newton = 1;
    for (int k = 6; k + 1 < argc; k++)
      if (std::string(argv[k]) == "newton") newton = std::atoi(argv[k + 1]);
    nls.SetMaxNumberOfNonLinearIterations(newton);
    nls.SetNonLinearConvergenceTolerance(1.e-30);
    sys = &nls;
  } else {
    sys = &ml_prob.add_system<LinearImplicitSystem>("Stokes");
  }
  LinearImplicitSystem& system = *sys;
  system.AddSolutionToSystemPDE("U");
  system.AddSolutionToSystemPDE("V");
  system.AddSolutionToSystemPDE("W");
  system.AddSolutionToSystemPDE("P");
  if (ns) system.SetAssembleFunction(femus::AssembleNavierStokes_AD);
  else system.SetAssembleFunction(AssembleMatrixResSteadyStokes);
  system.SetMaxNumberOfLinearIterations(1);
  system.SetAbsoluteLinearConvergenceTolerance(1.e-30);
  system.SetMgType(V_CYCLE);
  system.SetNumberPreSmoothingStep(1);
  system.SetNumberPostSmoothingStep(1);
  system.init();
  system.SetSolverFineGrids(RICHARDSON);
  system.SetPreconditionerFineGrids(JACOBI_PRECOND);
  system.SetTolerances(1.e-12, 1.e-20, 1.e+50, 1);
  system.ClearVariablesToBeSolved();
  system.AddVariableToBeSolved("All");
  system.SetDirichletBCsHandling(PENALTY);
  // the fields the callback will read: the initial functions, overwritten on Dirichlet nodes by the boundary values of
  // SetBoundaryCondition (GenerateBdc) -- written next to the backend's dump, in solution-dof numbering per variable
  if (const char* dir = std::getenv("FEMUS_REF_DUMP")) {
    const unsigned top = ml_msh.GetNumberOfLevels() - 1;
    const char* names[4] = {"U", "V", "W", "P"};
    for (int k = 0; k < 4; k++) {
      const NumericVector& v = *ml_sol.GetSolutionLevel(top)->_Sol[ml_sol.GetIndex(names[k])];
      std::vector<double> a(v.size());
      for (unsigned i = 0; i < a.size(); i++) a[i] = v(i);
      std::ofstream f(std::string(dir) + "/L" + std::to_string(top) + "_SOL_" + names[k] + ".f8", std::ios::binary);
      f.write(reinterpret_cast<const char*>(a.data()), (std::streamsize)(a.size() * sizeof(double)));
    }
  }
  system.MGsolve();      // assembly on the finest level, Galerkin chain, level setup (the dump happens there), one cycle

  std::cout << "IReynolds " << std::setprecision(17) << (ns ? 1. : ml_prob.parameters.get<Fluid>("Fluid").get_IReynolds_number()) << std::endl;
  ml_prob.clear();
  return 0;
}
