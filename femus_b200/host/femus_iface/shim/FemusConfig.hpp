// What FEMuS's CMake would generate from src/00_utils/FemusConfig.hpp.in with the B200 backend
// selected and no PETSc: used only to compile the adapters against the reference's own headers
// (tests/test_adapters.py, -DB2_WITH_FEMUS_HEADERS).  INTEGRATION.md describes the real change.
#ifndef __femus_FemusConfig_hpp__
#define __femus_FemusConfig_hpp__
#define HAVE_B200
#undef LSOLVER
#define LSOLVER PETSC_SOLVERS      /* becomes B200_SOLVERS once the enum value exists */
#endif
