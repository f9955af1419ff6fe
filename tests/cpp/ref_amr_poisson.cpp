// A FEMuS application for the AMR row of the path, written against the reference's public API and
// compiled together with the reference's own sources (oracle/ref_build, femus_b200/ref_build.py): selectively refined
// 3-D meshes -> hanging-node constraint matrices (Mesh::GetAMRRestrictionAndAMRSolidMark, Mesh.cpp:1354-...;
// LinearImplicitSystem::BuildAmrProlongatorMatrix, LinearImplicitSystem.cpp:912-1028), KK <- Pamr^T KKamr Pamr on the
// non-homogeneous levels (:329-342), prolongators multiplied by the constraint matrices (:255-258), V-cycle or
// F-cycle (:296-312).  Everything above the algebra backend is the REFERENCE's code; the element loop and the
// boundary function are those of applications/001_Poisson/main.cpp, compiled in place (its main() is renamed away),
// the refinement criterion is the shrinking circle of applications/MGAMR/ex5/ex5.cpp:49-70 around the box centre.
//
//   ref_amr_poisson <nx> <uniform levels> <selective levels> <V|F> <jacobi|sor> [cycles] [device]
//
// "device" (B200 build only): the element loop itself runs on the GPU -- femus::AssemblePoissonB200
// (femus_b200/host/RefAssemble.hpp, the replacement callback of INTEGRATION.md section 5) is registered instead of
// 001_Poisson's callback; everything else stays the reference's code.
//
// prints the reference's own "Linear Res L2norm" lines (LinearImplicitSystem.cpp:426).  Test infrastructure: the host
// backend run generates tests/golden/ref_amr_*.npz, the B200 backend run is compared with it on the GPU.
#define main femus_001_poisson_main_unused
#include "applications/001_Poisson/main.cpp"
#undef main
#ifdef B2_REF_DEVICE_ASSEMBLY
#include "RefAssemble.hpp"
#endif

static bool RefineInsideShrinkingCircle(const std::vector<double>& x, const int& /*elemgroupnumber*/, const int& level) {
  const double radius = 0.25 / level;
  return x[0] * x[0] + x[1] * x[1] < radius * radius;
}

int main(int argc, char** argv) {
  if (argc < 6) {
    std::cerr << "usage: " << argv[0] << " <nx> <uniform levels> <selective levels> <V|F> <jacobi|sor> [cycles]\n";
    return 1;
  }
  const unsigned nx = std::atoi(argv[1]), nUniform = std::atoi(argv[2]), nSelective = std::atoi(argv[3]);
  const bool fcycle = argv[4][0] == 'F';
  const bool sor = std::string(argv[5]) == "sor";
  const unsigned cycles = argc > 6 ? std::atoi(argv[6]) : 6;
  const bool device = argc > 7 && std::string(argv[7]) == "device";

  FemusInit init(argc, argv, MPI_COMM_WORLD);
  Files files;
  files.CheckIODirectories(true);

  MultiLevelMesh ml_msh;
  ml_msh.GenerateCoarseBoxMesh(nx, nx, nx, -0.5, 0.5, -0.5, 0.5, -0.5, 0.5, HEX27, "seventh");
  ml_msh.RefineMesh(nUniform + nSelective, nUniform, RefineInsideShrinkingCircle);
  ml_msh.PrintInfo();

  MultiLevelSolution ml_sol(&ml_msh);
  ml_sol.AddSolution("Sol", LAGRANGE, SECOND);
  ml_sol.Initialize("All");
  ml_sol.AttachSetBoundaryConditionFunction(SetBoundaryCondition);      // 001_Poisson's: Dirichlet 0, flux 0.2 on face 3
  ml_sol.GenerateBdc("Sol");

  fpsource.SetExpression("1.");
  fpsource.SetIndependentVariables("x,y,z,t");
  fpsource.Parse();

  MultiLevelProblem ml_prob(&ml_sol);
  LinearImplicitSystem& system = ml_prob.add_system<LinearImplicitSystem>("Poisson");
  system.AddSolutionToSystemPDE("Sol");
  system.SetAssembleFunction(AssemblePoissonMatrixandRhs);
  if (device) {
#ifdef B2_REF_DEVICE_ASSEMBLY
    system.SetAssembleFunction(femus::AssemblePoissonB200);      // nu = 1, f = 1: what fpsource and main.cpp:391-413 say in 3-D
    std::cout << " element loop: femus::AssemblePoissonB200 (device)" << std::endl;
#else
    std::cerr << "this build has no device assembly" << std::endl;
    return 1;
#endif
  }
  system.SetMaxNumberOfLinearIterations(cycles);
  system.SetAbsoluteLinearConvergenceTolerance(1.e-30);
  system.SetMgType(fcycle ? F_CYCLE : V_CYCLE);
  system.SetNumberPreSmoothingStep(1);
  system.SetNumberPostSmoothingStep(1);
  system.init();
  system.SetSolverFineGrids(RICHARDSON);
  system.SetPreconditionerFineGrids(sor ? SOR_PRECOND : JACOBI_PRECOND);
  system.SetTolerances(1.e-12, 1.e-20, 1.e+50, 4);
  system.ClearVariablesToBeSolved();
  system.AddVariableToBeSolved("All");
  system.SetDirichletBCsHandling(PENALTY);
  system.MGsolve();

  // the solution itself: a checksum of every level's Sol (the hanging-node values are interpolated ones)
  for (unsigned l = 0; l < ml_msh.GetNumberOfLevels(); l++) {
    const NumericVector& s = *ml_sol.GetSolutionLevel(l)->_Sol[ml_sol.GetIndex("Sol")];
    std::cout << "AMR level " << l << " dofs " << s.size() << " Sol l2 " << std::scientific << std::setprecision(12) << s.l2_norm() << " linf "
              << s.linfty_norm() << std::endl;
  }
  ml_prob.clear();
  return 0;
}
