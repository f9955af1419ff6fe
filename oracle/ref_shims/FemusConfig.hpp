// TEST INFRASTRUCTURE (oracle/): the header FEMuS's CMake generates from src/00_utils/FemusConfig.hpp.in, written by hand
// for the single-process oracle build of the reference's OWN mesh / solution / system sources (oracle/ref_build).
// HAVE_PETSC is defined so that LSOLVER = PETSC_SOLVERS and the reference's three factories (NumericVector::build,
// SparseMatrix::build, LinearEquationSolver::build) take their PETSc branch -- but PETSc itself is absent: the include
// guards of the reference's Petsc* class headers are pre-defined here (their content is skipped) and the classes of
// those NAMES are provided by the backend header the build pre-includes (-include): a plain host backend for the
// oracle (oracle/ref_build/HostBackend.hpp), or femus_b200's B200 adapters for the drop-in run on the GPU.
#ifndef __femus_FemusConfig_hpp__
#define __femus_FemusConfig_hpp__
#include <climits>
#include <cstdlib>
#include <iostream>
#include <cmath>
using std::isnan;      // the reference calls isnan(double) unqualified (petsc.h drags <math.h> in)
using std::isinf;
#define FEMTTU_VERSION_MAJOR 1
#define FEMTTU_VERSION_MINOR 0
#define HAVE_MPI
#define HAVE_PETSC
#define HAVE_JSONCPP
#define HAVE_ADEPT
#define HAVE_B64
#define HAVE_METIS
#define HAVE_FPARSER
#define HAVE_HDF5
#undef LSOLVER
#define LSOLVER PETSC_SOLVERS
#define FEMTTU_DETECTED_PETSC_VERSION_MAJOR 3
#define FEMTTU_DETECTED_PETSC_VERSION_MINOR 20
#define FEMTTU_DETECTED_PETSC_VERSION_SUBMINOR 2

// the reference's PETSc-backed class headers: skipped (see above)
#define __femus_algebra_PetscVector_hpp__
#define __femus_algebra_PetscMatrix_hpp__
#define __femus_algebra_PetscPreconditioner_hpp__
#define __femus_algebra_PetscMacro_hpp__
#define __femus_algebra_LinearEquationSolverPetsc_hpp__
#define __femus_algebra_LinearEquationSolverPetscAsm_hpp__
#define __femus_algebra_LinearEquationSolverPetscFieldSplit_hpp__
#define __femus_enums_FieldSplitTree_hpp__

// opaque PETSc handle types that the reference's solver-independent headers mention
// (LinearEquationSolver.hpp:132 `virtual KSP* GetKSP()`, LinearImplicitSystem.cpp:1065 `PetscInt`)
typedef struct femus_b200_opaque_KSP* KSP;
typedef struct femus_b200_opaque_PC* PC;
typedef struct femus_b200_opaque_IS* IS;
typedef struct femus_b200_opaque_Mat* Mat;
typedef struct femus_b200_opaque_Vec* Vec;
typedef int PetscInt;
typedef int PetscErrorCode;
typedef double PetscScalar;
typedef double PetscReal;
#define PETSC_COMM_WORLD MPI_COMM_WORLD
#define CHKERRABORT(comm, ierr) do { if (ierr) abort(); } while (0)
static inline int PetscInitialize(int*, char***, const char*, const char*) { return 0; }
static inline int PetscFinalize() { return 0; }
typedef void* PetscViewer;
#define PETSCVIEWERASCII "ascii"
static inline int PetscViewerCreate(int, PetscViewer*) { return 0; }
static inline int PetscViewerSetType(PetscViewer, const char*) { return 0; }
static inline int PetscViewerFileSetName(PetscViewer, const char*) { return 0; }
static inline int PetscLogView(PetscViewer) { return 0; }
namespace femus {
class FieldSplitTree;      // FieldSplitTree.hpp is PETSc code; the base solver interface only passes pointers to it
}
#endif
