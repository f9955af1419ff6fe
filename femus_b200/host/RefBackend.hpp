// The femus_b200 backend INSIDE a FEMuS build: what the reference's three factories return when LSOLVER selects it
// (NumericVector::build, SparseMatrix::build, LinearEquationSolver::build -- NumericVector.cpp:35-56,
// SparseMatrix.cpp:42-63, LinearEquationSolver.cpp:40-74).  With this header pre-included (-include) into those three
// translation units, the reference's UNMODIFIED sources and applications (applications/001_Poisson/main.cpp) run on
// libfemus_b200.so: every vector / matrix / solver object the application touches is a device object behind the
// reference's own abstract interfaces.  tests/ref_apps_build.py builds exactly that (single rank, the reference's MPI /
// PETSc headers replaced by the single-process shims of the test build -- build plumbing, no arithmetic).
//
// LinearEquationSolverB200Ref derives from the reference's LinearEquationSolver (LinearEquationSolver.hpp:55) and
// implements its pure virtuals on b2_mg_*: MGInit / MGSetLevel / MGSolve / Solve as LinearEquationSolverPetsc.cpp:185-353
// does on PCMG.  Level smoothers: Richardson(scale) around Jacobi (JACOBI_PRECOND) or around one SOR iteration of the
// whole level (SOR_PRECOND, what 001_Poisson sets, main.cpp:239-242: the one-block SSOR sweep of b2_schwarz, rows of a
// dependency level in parallel).  Level 0: exact solve (one-block dense inverse) up to 4096 rows, Jacobi-PCG to 1e-14
// beyond (the reference: PREONLY + LU through MUMPS).
#pragma once
#ifndef B2_WITH_FEMUS_HEADERS
#define B2_WITH_FEMUS_HEADERS
#endif
#include <fstream>
#include <string>
#include <type_traits>
#include "B200Matrix.hpp"
#include "LinearEquationSolver.hpp"
#include "Mesh.hpp"
#include "Solution.hpp"

namespace femus {

class LinearEquationSolverB200Ref : public LinearEquationSolver {
 public:
  LinearEquationSolverB200Ref(const unsigned& igrid, Solution* other_solution)
      : LinearEquationSolver(igrid, other_solution), _level(igrid), _richardsonScaleFactor(0.5), _mg(nullptr), _levelMax(0), _sweep(nullptr),
        _coarse(nullptr), _bdcIndexIsInitialized(false) {}
  ~LinearEquationSolverB200Ref() {
    this->MGClear();
    if (_sweep) b2_schwarz_destroy(_sweep);
    if (_coarse) b2_schwarz_destroy(_coarse);
  }

  void SetTolerances(const double&, const double&, const double&, const unsigned&, const unsigned&) override {}
  void SetRichardsonScaleFactor(const double& richardsonScaleFactor) override { _richardsonScaleFactor = richardsonScaleFactor; }

  // rows that are Dirichlet (Bdc < 1.5) or whose variable is not solved (LinearEquationSolverPetsc.cpp:53-90)
  void BuildBdcIndex(const std::vector<unsigned>& variable_to_be_solved) {
    _bdcIndexIsInitialized = true;
    _bdcIndex.clear();
    std::vector<bool> included(_SolPdeIndex.size(), false);
    for (unsigned v : variable_to_be_solved) included[v] = true;
    for (unsigned k = 0; k < _SolPdeIndex.size(); k++) {
      const unsigned indexSol = _SolPdeIndex[k], soltype = _SolType[indexSol];
      const unsigned i0 = GetMeshFromLinEq()->_dofOffset[soltype][processor_id()], i1 = GetMeshFromLinEq()->_dofOffset[soltype][processor_id() + 1];
      for (unsigned i = i0; i < i1; i++)
        if (!included[k] || (*(*_Bdc)[indexSol])(i) < 1.5) _bdcIndex.push_back((int32_t)(KKoffset[k][processor_id()] + (i - i0)));
    }
    std::sort(_bdcIndex.begin(), _bdcIndex.end());
  }

  void MGInit(const MgSmootherType& mg_smoother_type, const unsigned& levelMax, const SolverType&) override {
    if (mg_smoother_type != MULTIPLICATIVE) { std::fprintf(stderr, "femus_b200: MGInit: only the multiplicative V-cycle is implemented\n"); std::abort(); }
    this->MGClear();
    _levelMax = levelMax;
    B2_ABORT_IF(b2_mg_create(B200Context::get(), (int)levelMax, &_mg), "b2_mg_create");
    B2_ABORT_IF(b2_mg_set_coarse(_mg, 1.e-14, 10000), "b2_mg_set_coarse");
  }
  void MGClear() override {
    if (_mg) b2_mg_destroy(_mg);
    _mg = nullptr;
  }
  void MGSetLevel(LinearEquationSolver* LinSolver, const unsigned& levelMax, const std::vector<unsigned>& variable_to_be_solved, SparseMatrix* PP,
                  SparseMatrix*, const unsigned& npre, const unsigned& npost) override {
    LinearEquationSolverB200Ref* top = static_cast<LinearEquationSolverB200Ref*>(LinSolver);
    if (!top->_mg || levelMax + 1 != top->_levelMax) { std::fprintf(stderr, "femus_b200: MGSetLevel: MGInit was not called on the finest solver\n"); std::abort(); }
    if (!_bdcIndexIsInitialized) this->BuildBdcIndex(variable_to_be_solved);
    if (const char* dir = std::getenv("FEMUS_REF_DUMP")) this->dump_level(dir, PP);      // before the penalty
    this->configure_level(top->_mg, (int)_level, PP, npre, npost);
  }
  // one multiplicative V-cycle as outer PREONLY (:294-353): ZerosBoundaryResiduals; EPSC = V(RES); RESC = KK EPSC; RES -= RESC; EPS += EPSC
  void MGSolve(const bool) override {
    if (!_mg) { std::fprintf(stderr, "femus_b200: MGSolve on a level that does not own the multigrid\n"); std::abort(); }
    B200Vector &RES = static_cast<B200Vector&>(*_RES), &EPS = static_cast<B200Vector&>(*_EPS);
    RES.close();
    EPS.close();
    const char* dir = _dumped ? nullptr : std::getenv("FEMUS_REF_DUMP");
    if (dir) dump_vector(dir, "RES", RES);
    B2_ABORT_IF(b2_mg_solve(_mg, RES.handle(), EPS.handle()), "b2_mg_solve");
    RES.touched();
    EPS.touched();
    if (dir) {
      dump_vector(dir, "RES_after", RES);
      dump_vector(dir, "EPS_after", EPS);
    }
    _dumped = true;
  }
  // one level, no multigrid: penalty + the solve of that level (a one-level hierarchy: its level 0 is solved)
  void Solve(const std::vector<unsigned>& variable_to_be_solved, const bool&) override {
    if (!_bdcIndexIsInitialized) this->BuildBdcIndex(variable_to_be_solved);
    this->MGClear();
    _levelMax = 1;
    B2_ABORT_IF(b2_mg_create(B200Context::get(), 1, &_mg), "b2_mg_create");
    B2_ABORT_IF(b2_mg_set_coarse(_mg, 1.e-14, 100000), "b2_mg_set_coarse");
    this->configure_level(_mg, 0, nullptr, 0, 0);
    this->MGSolve(false);
  }

 private:
  // ---- diagnostics (tools/ref_app_diag.py): level operators, prolongators and vectors read back from the device objects
  template <class T>
  void dump_array(const char* dir, const std::string& name, const std::vector<T>& a) const {
    const char* suffix = sizeof(T) == 8 ? (std::is_floating_point<T>::value ? "f8" : "i8") : "i4";
    std::ofstream f(std::string(dir) + "/L" + std::to_string(_level) + "_" + name + "." + suffix, std::ios::binary);
    f.write(reinterpret_cast<const char*>(a.data()), (std::streamsize)(a.size() * sizeof(T)));
  }
  void dump_vector(const char* dir, const char* name, const NumericVector& v) const {
    std::vector<double> a((size_t)v.size());
    for (int i = 0; i < v.size(); i++) a[(size_t)i] = v(i);
    dump_array(dir, name, a);
  }
  void dump_csr(const char* dir, const std::string& name, const B200Matrix& A) const {
    dump_array(dir, name + "_rowptr", A.host_rowptr());
    dump_array(dir, name + "_col", A.host_col());
    dump_array(dir, name + "_val", A.host_val());
    dump_array(dir, name + "_shape", std::vector<int>{A.m(), A.n()});
  }
  void dump_level(const char* dir, SparseMatrix* PP) const {
    dump_csr(dir, "KK", B200Matrix::cast(*_KK));
    if (PP) dump_csr(dir, "PP", B200Matrix::cast(*PP));
    dump_array(dir, "bdcIndex", _bdcIndex);
  }
  bool _dumped = false;

  void configure_level(b2_mg* mg, const int level, SparseMatrix* PP, const unsigned npre, const unsigned npost) {
    B200Matrix& KK = static_cast<B200Matrix&>(*_KK);
    KK.close();
    b2_csr* P = nullptr;
    if (level > 0) {
      const B200Matrix& Pm = B200Matrix::cast(*PP);
      Pm.close();
      P = Pm.handle();
    }
    const int64_t n = KK.m();
    if (level > 0) {
      if (this->_levelSolverType != RICHARDSON) { std::fprintf(stderr, "femus_b200: level solver %d: Richardson is what this adapter drives\n", (int)this->_levelSolverType); std::abort(); }
      const PreconditionerType pc = this->preconditioner_type();
      if (pc == JACOBI_PRECOND) {
        B2_ABORT_IF(b2_mg_set_smoother(mg, level, 0, 0., 0.), "b2_mg_set_smoother");
      } else if (pc == SOR_PRECOND) {          // PCSOR on the level = the SSOR sweep of ONE block holding every dof
        if (_sweep && _sweepGen != KK.generation()) { b2_schwarz_destroy(_sweep); _sweep = nullptr; }      // the level matrix was rebuilt (F-cycle)
        if (!_sweep) {
          _sweepGen = KK.generation();
          const int64_t bp[2] = {0, n}, gp[2] = {0, 1};
          std::vector<int32_t> all((size_t)n);
          for (int64_t i = 0; i < n; i++) all[(size_t)i] = (int32_t)i;
          const int32_t gb[1] = {0};
          B2_ABORT_IF(b2_schwarz_create(B200Context::get(), KK.handle(), 1, bp, all.data(), 1, gp, gb, &_sweep), "b2_schwarz_create");
          B2_ABORT_IF(b2_schwarz_set_subsolver(_sweep, 1), "b2_schwarz_set_subsolver");
          B2_ABORT_IF(b2_schwarz_set_row_levels(_sweep, 1), "b2_schwarz_set_row_levels");
        }
        B2_ABORT_IF(b2_mg_set_level_schwarz(mg, level, _sweep), "b2_mg_set_level_schwarz");
      } else {
        std::fprintf(stderr, "femus_b200: preconditioner %d on the levels: Jacobi and SOR are what this adapter drives\n", (int)pc);
        std::abort();
      }
    } else if (n <= 4096) {                    // exact coarse solve, like the reference's LU
      if (_coarse && _coarseGen != KK.generation()) { b2_schwarz_destroy(_coarse); _coarse = nullptr; }
      if (!_coarse) {
        _coarseGen = KK.generation();
        const int64_t bp[2] = {0, n}, gp[2] = {0, 1};
        std::vector<int32_t> all((size_t)n);
        for (int64_t i = 0; i < n; i++) all[(size_t)i] = (int32_t)i;
        const int32_t gb[1] = {0};
        B2_ABORT_IF(b2_schwarz_create(B200Context::get(), KK.handle(), 1, bp, all.data(), 1, gp, gb, &_coarse), "b2_schwarz_create");
      }
      B2_ABORT_IF(b2_mg_set_coarse_schwarz(mg, _coarse), "b2_mg_set_coarse_schwarz");
    }
    // SetPenalty (:428-436) happens inside: Dirichlet rows -> identity, pattern kept
    B2_ABORT_IF(b2_mg_set_level(mg, level, KK.handle(), P, _bdcIndex.data(), (int64_t)_bdcIndex.size(), (int)npre, (int)npost, _richardsonScaleFactor),
                "b2_mg_set_level");
    KK.touched();
  }

  unsigned _level;
  double _richardsonScaleFactor;
  b2_mg* _mg;
  unsigned _levelMax;
  b2_schwarz *_sweep, *_coarse;
  uint64_t _sweepGen = 0, _coarseGen = 0;      // generation of the level matrix the two block objects were created on
  std::vector<int32_t> _bdcIndex;
  bool _bdcIndexIsInitialized;
};

// the names the reference's factories instantiate
using PetscVector = B200Vector;
using PetscMatrix = B200Matrix;
using LinearEquationSolverPetsc = LinearEquationSolverB200Ref;
using LinearEquationSolverPetscAsm = LinearEquationSolverB200Ref;
using LinearEquationSolverPetscFieldSplit = LinearEquationSolverB200Ref;

}  // namespace femus
